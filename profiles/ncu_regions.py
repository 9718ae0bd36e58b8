"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export of walk_kernel by code region.
usage: python profiles/ncu_regions.py <both.csv> <reads_in_launch>"""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
nreads = float(sys.argv[2])
hdr = None
seq = []
cur = None; curfile = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        curfile = r[1]; continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r; iInst = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples"); iThr = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) <= iInst:
        continue
    if r[0] != "":
        try: cur = (curfile.split("/")[-1], int(r[0]), r[1].strip())
        except ValueError: pass
        continue
    try: seq.append((cur, r[3].strip(), int(r[iInst]), int(r[iThr]), int(r[iSamp])))
    except ValueError: pass
tot_i = sum(s[2] for s in seq); tot_t = sum(s[3] for s in seq); tot_s = sum(s[4] for s in seq)
print(f"sass {len(seq)}  warp-inst/read {tot_i/nreads:.0f}  thread-inst/32/read {tot_t/32/nreads:.0f}  samples {tot_s}")
byline = {}
for cur, sass, i, t, s in seq:
    a = byline.setdefault(cur, [0, 0, 0]); a[0] += i; a[1] += t; a[2] += s
print("--- top source lines by issued warp instructions")
for k, a in sorted(byline.items(), key=lambda x: -x[1][0])[:40]:
    print(f"{100*a[0]/tot_i:5.1f}% inst {a[0]/nreads:6.0f}/read thr/inst {a[1]/max(a[0],1):5.1f} samp {100*a[2]/tot_s:5.1f}%  {k[0]}:{k[1]} {k[2][:90]}")
print("--- top source lines by stall samples")
for k, a in sorted(byline.items(), key=lambda x: -x[1][2])[:25]:
    print(f"{100*a[2]/tot_s:5.1f}% samp {100*a[0]/tot_i:5.1f}% inst  {k[0]}:{k[1]} {k[2][:100]}")
