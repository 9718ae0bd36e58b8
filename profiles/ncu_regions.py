import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
kern=None; hdr=None; cur=None; curfile=None; data={}
for r in rows:
    if len(r)>=2 and r[0]=="Function Name": kern=r[1]; data.setdefault(kern,{}); continue
    if len(r)>=2 and r[0]=="File Path": curfile=r[1]; continue
    if len(r)>10 and r[0]=="Line No":
        hdr=r; iI=hdr.index("Instructions Executed"); iT=hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r)<=iI or kern is None: continue
    if r[0]!="":
        try: cur=(curfile.split("/")[-1], int(r[0]))
        except ValueError: pass
        continue
    try: data[kern][(r[2],cur)]=(int(r[iI]),int(r[iT]))
    except ValueError: pass
src=open('/root/repo/nptranscript_b200/csrc/kernels.cuh').read().splitlines()
def find(s): return next(i+1 for i,l in enumerate(src) if s in l)
L=dict(add=find("auto add = [&]"), depth=find("auto depth_run"), loop=find("for (;;) {"), newr=find("(A0) read transitions"), endr=find("---- end of read: CigarCluster.addEnd"),
       op=find("(Rc) CIGAR tile of the lanes"), sync2=find("if (__all_sync(0xffffffffu, done)) break;"), refill=find("---- (R) the refill point"),
       win=find("uint32_t mis = 0;  // mismatches of this step"), red=find("if (MODE == 1) {\n") if False else find("// error positions: mismatches of this step"), end=find("if (MODE == 0 && slow_windows)"))
regions=[("add lambda",L['add'],L['depth']-1),("depth_run",L['depth'],L['loop']-1),("A: loop/new read",L['loop'],L['endr']-1),("A: end read",L['endr'],L['op']-1),("A: op",L['op'],L['sync2']-2),
 ("sync/all",L['sync2']-1,L['sync2']),("R: refill",L['refill'],L['win']-2),("B: window",L['win']-1,L['red']-2),("REDs",L['red']-1,L['end']-1)]
n=float(sys.argv[2])
for k,d in data.items():
    if 'walk_kernel<(int)2' in k: continue
    tot=sum(v[0] for v in d.values())
    print(k[:50], f"{tot/n:.0f} warp-inst/read")
    agg={}
    for (addr,(f,l)),(i,t) in d.items():
        name='other:'+f
        if f=='kernels.cuh':
            name='kernels.cuh:other'
            for nm,a,b in regions:
                if a<=l<=b: name=nm
        a=agg.setdefault(name,[0,0]); a[0]+=i; a[1]+=t
    for nm,a in sorted(agg.items(), key=lambda x:-x[1][0]):
        print(f"   {nm:40s} {a[0]/n:8.0f}/read {100*a[0]/tot:5.1f}%  lanes {a[1]/max(a[0],1):4.1f}")
