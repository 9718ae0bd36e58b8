"""Host-side timeline of the end-to-end path: how long each C-ABI call takes, and the raw pinned H2D/D2H bandwidth of the box.
usage: python profiles/e2e_trace.py [reads_per_chunk] [chunks]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np
import torch
from nptranscript_b200 import abi, engine, synth, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for name, a, b in (("h2d", x, y), ("d2h", y, x)):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"pinned {name}: {3 * x.numel() / dt / 1e9:.1f} GB/s")
w = workloads.SARS2; g = synth.genome_codes(w); L = engine.lib()
pinned = []
def palloc(nbytes):
    p = L.npt_alloc_pinned(nbytes); pinned.append(p)
    return (ctypes.c_char * nbytes).from_address(p)
bs = [synth.generate(w, 1, n, first_index=i * n, alloc=palloc) for i in range(chunks)]
eng = engine.Engine(engine.make_config(bin=100, break_thresh=1000, record_depth=1, max_coverage_clusters=2048))
for it in range(3):
    T = {}
    def tick(k, t0): torch.cuda.synchronize() if k.endswith("+sync") else None; T[k] = T.get(k, 0) + time.perf_counter() - t0
    t_all = time.perf_counter()
    t = time.perf_counter(); eng.begin_chrom(0, g, w.annotation); tick("begin_chrom", t)
    for b in bs:
        t = time.perf_counter(); eng.submit(b); tick("submit", t)
    t = time.perf_counter(); eng.end_chrom(); tick("end_chrom", t)
    t = time.perf_counter(); r = eng.fetch_raw(abi.NPT_FETCH_ALL); tick("fetch", t)
    t = time.perf_counter(); eng.release(); tick("release", t)
    tot = time.perf_counter() - t_all
    print(f"iter {it}: total {tot*1e3:.1f} ms = {chunks*n/tot/1e6:.1f} M reads/s  " + "  ".join(f"{k} {v*1e3:.1f}" for k, v in T.items()), eng.kernel_ms())
