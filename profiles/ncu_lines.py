"""Per-kernel source-line ranking from `ncu --page source --csv --print-source cuda,sass`.
usage: python profiles/ncu_lines.py <both.csv> <reads_per_launch> [kernel substring] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); n = float(sys.argv[2]); want = sys.argv[3] if len(sys.argv) > 3 else ""; top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
kern = None; data = {}; hdr = None; cur = None; curfile = None
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name": kern = r[1]; data.setdefault(kern, {}); continue
    if len(r) >= 2 and r[0] == "File Path": curfile = r[1]; continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iT = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) <= iI or kern is None: continue
    if r[0] != "":
        try: cur = (curfile.split("/")[-1], int(r[0]), r[1].strip())
        except ValueError: pass
        continue
    try: data[kern][(r[2], cur)] = (int(r[iI]), int(r[iT]), int(r[iS]))   # keyed by SASS address: de-duplicates repeated sections
    except ValueError: pass
for k, d in data.items():
    if want not in k: continue
    ti = sum(v[0] for v in d.values()); tt = sum(v[1] for v in d.values()); ts = sum(v[2] for v in d.values())
    print(f"===== {k[:70]}  warp-inst/read {ti/n:.0f}  thread-inst/read {tt/n:.0f}  avg lanes {tt/max(ti,1):.1f}  static sass {len(d)}")
    by = {}
    for (addr, cur), v in d.items():
        a = by.setdefault(cur, [0, 0, 0]); a[0] += v[0]; a[1] += v[1]; a[2] += v[2]
    for c, a in sorted(by.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{100*a[0]/ti:5.1f}% inst {a[0]/n:6.0f}/read lanes {a[1]/max(a[0],1):5.1f} stall-samples {100*a[2]/max(ts,1):5.1f}%  {c[0]}:{c[1]} {c[2][:95]}")
