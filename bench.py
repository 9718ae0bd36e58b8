#!/usr/bin/env python
"""bench.py — reads/s of the read→transcript-cluster hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          # our arm (N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N ...          # CPU arm: the oracle port on the host cores

One step = one pass of the hot path (npt_begin_chrom → npt_submit → npt_end_chrom) over one chromosome's batch of
synthetic SARS-CoV-2-shape direct-RNA reads (BASELINE.json configs[1]: coverage + break points on).
  value   reads/s with the batch already resident in HBM (device timeline, CUDA events inside libnpt on its compute
          stream, max over ranks)
  e2e     reads/s through the C ABI with HOST (pinned) batches: H2D copies, kernels, end-of-chromosome finalisation and
          D2H of every result (per-read rows, tables, coverage rows) inside the timed region
  roofline  algorithmic bytes of the walk kernel / its CUDA-event duration vs the measured HBM copy peak
  cpu_baseline  the oracle (C++ restatement, NOT the JVM) timed on this box's host cores on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads/sec clustered"
UNIT = "reads/s"
CFG_KW = dict(bin=100, break_thresh=1000, record_depth=1, calc_breaks=1, track_break_sums=1, num_sources=1,
              max_coverage_clusters=2048)
SEED = 20200302
CHUNK = 1 << 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons of one GPU during the timed region.  In-process NVML (nvidia_ml_py) when available: an
    `nvidia-smi -lms` loop per rank serialises in the driver and stalled the other ranks' CUDA calls (profiles/NOTES.md);
    nvidia-smi stays as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.p, self.t = gpu, [], None, None
        self.sm, self.mx, self.reasons, self.stop_flag, self.nvml = [], [], set(), False, None

    def _nvml_loop(self):
        nv, h = self.nvml
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for name, bit in names:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        if self.gpu is None:
            return
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = [float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))]
            self.nvml = (nv, h)
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def algorithmic_bytes(batch_bytes_in, n_reads, total_breaks):
    """every input byte once + per-read output 40 + 4*nB."""
    return batch_bytes_in + 40 * n_reads + 4 * total_breaks


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def run_reference(args, rank, world):
    if rank != 0:
        return
    from nptranscript_b200 import synth, workloads
    from oracle import oracle
    w = workloads.SARS2
    g = synth.genome_codes(w)
    threads = os.cpu_count() or 1
    sample = args.ref_reads
    b = synth.generate(w, SEED, sample)
    cfg = oracle.make_config(**CFG_KW)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o = oracle.OracleRun(cfg, g, w.annotation, 0, keep_reads=True)
        o.submit(b, nthreads=threads)
        o.close()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    v = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"{w.name}: {sample} reads/step (bounded sample of the 10M-read config), coverage+breakpoints on",
                   "genome_len": w.genome_len, "flags": CFG_KW},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} reads, oracle.cpp (C++ restatement of the Java path, not the JVM), per-read stage on {threads} threads, sequential commit"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm (GPU)
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from nptranscript_b200 import abi, engine, synth, workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    w = workloads.SARS2
    g = synth.genome_codes(w)
    n_reads = args.reads
    cfg_kw = dict(CFG_KW, device=local_rank)
    cfg = engine.make_config(**cfg_kw)
    L = engine.lib()

    # ---- synthetic reads for this rank (weak scaling: every rank has its own n_reads, a contiguous range of the global order)
    first = rank * n_reads
    t0 = time.perf_counter()
    nchunks = (n_reads + CHUNK - 1) // CHUNK
    keep_host = min(nchunks, args.e2e_chunks)  # pinned host chunks kept for the e2e leg
    host_chunks, dev_parts = [], []
    pinned = []

    def pinned_alloc(nbytes):
        import ctypes
        p = L.npt_alloc_pinned(nbytes)
        if not p:
            raise MemoryError("npt_alloc_pinned")
        pinned.append(p)
        return (ctypes.c_char * nbytes).from_address(p)

    cig_base = seq_base = 0
    parts = {k: [] for k in ("pos", "flag", "src", "cigar", "seq4")}
    coffs, soffs = [], []
    in_bytes = 0
    for c in range(nchunks):
        m = min(CHUNK, n_reads - c * CHUNK)
        b = synth.generate(w, SEED, m, first_index=first + c * CHUNK, alloc=pinned_alloc if c < keep_host else None)
        in_bytes += b.nbytes()
        if c < keep_host:
            host_chunks.append(b)
        parts["pos"].append(torch.from_numpy(b.pos).cuda())
        parts["flag"].append(torch.from_numpy(b.flag.view(np.int16)).cuda())
        parts["src"].append(torch.from_numpy(b.src.view(np.int16)).cuda())
        parts["cigar"].append(torch.from_numpy(b.cigar.view(np.int32)).cuda())
        sq = torch.from_numpy(b.seq4).cuda()
        pad = (-sq.numel()) % 4  # keep every chunk's base address 4-byte aligned inside the concatenation
        if pad:
            sq = torch.cat([sq, torch.zeros(pad, dtype=torch.uint8, device="cuda")])
        parts["seq4"].append(sq)
        last = c == nchunks - 1
        coffs.append((b.cigar_off.astype(np.int64) + cig_base)[: None if last else -1])
        soffs.append((b.seq_off.astype(np.int64) + seq_base)[: None if last else -1])
        cig_base += int(b.cigar_off[-1])
        seq_base += int(b.seq_off[-1]) + pad
        if c >= keep_host:
            del b
    if cig_base >= 2 ** 32:
        raise SystemExit("batch too large for uint32 cigar offsets; lower --reads")
    d = {k: torch.cat(v) for k, v in parts.items()}
    d["cigar_off"] = torch.from_numpy(np.concatenate(coffs).astype(np.uint32).view(np.int32)).cuda()
    d["seq_off"] = torch.from_numpy(np.concatenate(soffs)).cuda()
    d["seq4"] = torch.cat([d["seq4"], torch.zeros(64, dtype=torch.uint8, device="cuda")])
    del coffs, soffs
    del parts
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0

    eng = engine.Engine(cfg)
    if world > 1:
        # the data path's collectives run inside libnpt on its own NCCL communicator (csrc/comm.inl); torch.distributed carries
        # the 128-byte id, the barriers and the max-over-ranks of the timings
        ident = [engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            eng.comm_init(ident[0], rank, world)
        finally:
            os.dup2(saved, 1)
            os.close(saved)

    def device_step():
        eng.set_read_index_base(first)
        eng.begin_chrom(0, g, w.annotation)
        eng.submit_device(n_reads, d["pos"].data_ptr(), d["flag"].data_ptr(), d["src"].data_ptr(), d["cigar_off"].data_ptr(),
                          d["cigar"].data_ptr(), d["seq_off"].data_ptr(), d["seq4"].data_ptr())
        kx_ms = red_ms = 0.0
        if world > 1:
            eng.end_chrom_all()   # key exchange + numbering + all-reduce of the counters: all inside the library's step span
            kx_ms, red_ms = eng.exchange_ms()
        else:
            eng.end_chrom()
        t = eng.kernel_ms()
        t["merge_ms"] = 0.0
        t["kx_ms"], t["reduce_ms"] = kx_ms, red_ms
        return t

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- HBM-resident leg
    for _ in range(args.warmup):
        if world > 1:
            dist.barrier()
        device_step()
        eng.release()
    sampler = ClockSampler(local_rank if rank == 0 else None)  # one nvidia-smi loop per job: its queries serialise in the driver
    barrier()
    sampler.start()
    step_ms, walk_ms, fin_ms, launches = [], [], [], 0
    stages = []
    kx_list, red_list = [], []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        if world > 1:
            dist.barrier()  # ranks start a step together: a step's span contains the key exchange, which waits for the slowest rank
        t = device_step()
        step_ms.append(t["step_ms"] + t["merge_ms"]); walk_ms.append(t["walk_ms"]); fin_ms.append(t["finalize_ms"])
        stages.append([t["fastwalk_ms"], t["walk0_ms"], t["walk2_ms"], t["cluster_ms"], t["walk1_ms"], t["fastcov_ms"]])
        launches += t["total_launches"]
        kx_list.append(t["kx_ms"]); red_list.append(t["reduce_ms"])
        last = eng.fetch_raw(abi.NPT_FETCH_READS) if _ == args.steps - 1 else None
        if last is not None:
            total_breaks = int(np.ctypeslib.as_array(last.break_off, shape=(n_reads + 1,))[-1])
            n_clusters, n_subs, n_slow = last.n_clusters, last.n_subs, last.n_slow_reads
        eng.release()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0) / args.steps
    clocks = sampler.stop()
    ms = float(np.mean(step_ms))
    if world > 1:
        tt = torch.tensor([ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = world * n_reads / (ms / 1e3)

    # ---- end-to-end leg: pinned host batches through npt_submit, everything fetched back
    e2e_reads = sum(b.n for b in host_chunks)
    h2d = sum(b.nbytes() for b in host_chunks)

    # the same chunks in the compact transfer format (npt_packed_batch: 16-bit CIGAR words, 2-bit bases), also pinned
    from nptranscript_b200 import packed as packed_mod
    packed_chunks = [packed_mod.pack(b, alloc=pinned_alloc) for b in host_chunks]
    h2d_packed = sum(pb.nbytes() for pb in packed_chunks)

    def e2e_step(use_packed=True):
        eng.set_read_index_base(first)
        eng.begin_chrom(0, g, w.annotation)
        if use_packed:
            for pb in packed_chunks:
                eng.submit_packed(pb)
        else:
            for b in host_chunks:
                eng.submit(b)
        if world > 1:
            eng.end_chrom_all()
        else:
            eng.end_chrom()
        r = eng.fetch_raw(abi.NPT_FETCH_ALL)
        nb = int(np.ctypeslib.as_array(r.break_off, shape=(e2e_reads + 1,))[-1])
        d2h = e2e_reads * (9 * 4 + 8) + nb * 4 + 4 * r.n_clusters * cfg.num_sources * r.cov_stride * 4
        eng.release()
        return d2h

    def e2e_leg(use_packed):
        e2e_ms, d2h = [], 0
        for it in range(max(1, args.warmup // 2) + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            d2h = e2e_step(use_packed)
            torch.cuda.synchronize()
            dt = 1e3 * (time.perf_counter() - t0)
            if it >= max(1, args.warmup // 2):
                e2e_ms.append(dt)
        t = float(np.mean(e2e_ms))
        if world > 1:
            tt = torch.tensor([t], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t, d2h

    e2e_plain_t, _ = e2e_leg(False)
    e2e_t, d2h = e2e_leg(True)
    e2e_value = world * e2e_reads / (e2e_t / 1e3)

    # ---- untimed self-check of the same end-to-end result: conservation sums that hold whatever the input size
    eng.set_read_index_base(first)
    eng.begin_chrom(0, g, w.annotation)
    for pb in packed_chunks:
        eng.submit_packed(pb)
    eng.end_chrom()
    res = eng.fetch()
    n_ok = int((res["cluster_id"] >= 0).sum())
    checks = {
        "reads_clustered == total_counts == sum(cluster counts) == sum(sub-cluster counts)":
            n_ok == int(res["total_counts"].sum()) == int(res["cl_read_count"].sum()) == int(res["sub_count"].sum()),
        "sum(depth) == sum(per-read emitted positions)": int(res["depth"].sum(dtype=np.int64)) == int(res["maps_sum"].sum(dtype=np.int64)),
        "sum(errors) == sum(per-read error positions)": int(res["errors"].sum(dtype=np.int64)) == int(res["err_sum"].sum(dtype=np.int64)),
        "sum(mapStart) == sum(mapEnd)": int(res["map_start"].sum(dtype=np.int64)) == int(res["map_end"].sum(dtype=np.int64)),
        "sum(sub-cluster break sums) == sum(per-read breaks)": int(res["sub_sums"].sum(dtype=np.int64)) == int(res["breaks"].sum(dtype=np.int64)),
        "cluster ids in first-appearance order": bool((np.diff(res["cl_first_read"]) > 0).all()),
    }
    eng.release()
    bad = [k for k, v in checks.items() if not v]
    if bad and not os.environ.get("NPT_LIBNPT"):  # an experimental library (NPT_LIBNPT) may compute wrong results on purpose
        raise SystemExit("bench self-check failed: " + "; ".join(bad))

    # ---- roofline of the dominant kernel: the one with the largest live CUDA-event share of the step.  The path needs all
    # of them, so the all-kernel pipeline figure is reported beside it ("pipeline").
    peak, peak_kind = peaks()
    alg = algorithmic_bytes(in_bytes, n_reads, total_breaks)
    wk = float(np.mean(walk_ms))
    st = np.mean(np.array(stages), axis=0)
    names = ["fast_walk_kernel (single pass: CIGAR -> 16-position tasks -> base compare -> breaks + coverage entries)",
             "walk_kernel<0> (exact serial walker, reads the fast walk refused)", "walk_kernel<2> (long break lists)",
             "cluster_kernel (tokens, hash tables, counters)", "walk_kernel<1> (coverage of the serial walker's reads)",
             "cov_key_kernel + radix sort + accumulate_kernel (positional popcount of the coverage entries per cluster)"]
    keys6 = ["fastwalk", "walk0", "walk2", "cluster", "walk1", "fastcov"]
    dom = int(np.argmax(st))
    dom_ms = float(st[dom])
    achieved = alg / (dom_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "walk_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj["dram_bytes_per_read"][keys6[dom]] * n_reads
        except Exception:
            traffic = None
    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle
        threads = os.cpu_count() or 1
        sample = min(args.cpu_reads, host_chunks[0].n)
        sb = host_chunks[0].slice(0, sample)
        ocfg = oracle.make_config(**CFG_KW)
        t0 = time.perf_counter()
        o = oracle.OracleRun(ocfg, g, w.annotation, 0, keep_reads=True)
        o.submit(sb, nthreads=threads)
        dt = time.perf_counter() - t0
        # the same sample through the device path, every field compared with the oracle's (untimed): a bench number whose
        # results differ from the reference restatement is not a number
        from tests.common import assert_same
        ora = o.results(coverage=True)
        eng.set_read_index_base(0)
        eng.begin_chrom(0, g, w.annotation)
        eng.submit(sb)
        eng.end_chrom()
        dev = eng.fetch()
        eng.release()
        assert_same(dev, ora)
        o.close()
        cpu = {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port", "parity": f"device == oracle on all fields of these {sample} reads",
               "sample": f"first {sample} reads of the same workload; oracle.cpp (C++ restatement, not the JVM), per-read stage on {threads} threads + sequential commit, {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"{w.name}: {n_reads} reads per GPU per step, 1 source BAM, coverage + break points on (BASELINE.json configs[1])",
                       "genome_len": w.genome_len, "flags": CFG_KW, "mean_read_bases": float(2 * (d['seq4'].numel() - 64) / n_reads),
                       "mean_cigar_ops": float(d["cigar"].numel() / n_reads), "input_bytes_per_read": in_bytes / n_reads,
                       "l2_policy": "inputs (%.1f GB) larger than L2; no flush needed" % (in_bytes / 1e9),
                       "clusters": n_clusters, "sub_clusters": n_subs, "reads_on_serial_path": n_slow,
                       "self_check": {"reads": e2e_reads, "passed": sorted(checks)}, "gen_seconds": gen_s,
                       "merge_breakdown_ms": {"key_exchange_ms": float(np.mean(kx_list)), "reduce_ms": float(np.mean(red_list)),
                                              "note": "CUDA-event spans on rank 0 inside npt_end_chrom_all: key exchange = header all-gather + grouped NCCL broadcasts + import kernels (includes waiting for the slowest rank); reduce = grouped NCCL all-reduces of the counters; both inside the step span"} if world > 1 else None,
                       "timing": "CUDA events inside libnpt on its compute stream (npt_last_step_ms / npt_last_kernel_ms); wall ms/step %.2f" % wall_ms},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": names[dom], "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / ms, "algorithmic_bytes": alg,
                         "algorithmic_bytes_per_read": alg / n_reads, "peak_source": peak_kind,
                         "kernels_ms": dict({k: float(v) for k, v in zip(keys6, st)}, finalize=float(np.mean(fin_ms))),
                         "pipeline": {"ms": wk, "achieved": alg / (wk / 1e3) / 1e9, "frac": alg / (wk / 1e3) / 1e9 / peak,
                                      "note": "all per-batch kernels together: the whole path, not only its dominant kernel"}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_packed, "d2h_bytes_per_step": d2h, "reads_per_step": e2e_reads,
                    "ms_per_step": e2e_t,
                    "note": "pinned host batches in the compact transfer format through npt_submit_packed (1Mi reads each; expanded on the device), "
                            "npt_end_chrom, npt_fetch_results(ALL)",
                    "plain_npt_batch": {"value": world * e2e_reads / (e2e_plain_t / 1e3), "ms_per_step": e2e_plain_t, "h2d_bytes_per_step": h2d,
                                        "note": "the same chunks as expanded npt_batch arrays through npt_submit"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    for p in pinned:
        L.npt_free_pinned(p)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (config 2: default 10 M; other configs: see bench_configs.py)")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="1Mi-read pinned host chunks used by the end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=500_000)
    ap.add_argument("--ref-reads", type=int, default=500_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json configuration 1..5 (2 = the headline one; the others: bench_configs.py)")
    ap.add_argument("--host-scale", type=float, default=0.04, help="config 5: chromosome lengths as a fraction of GRCh38")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.config != 2:
        import bench_configs
        if args.impl == "reference":
            bench_configs.run_reference(args, rank, world)
        else:
            bench_configs.run(args, rank, world, local_rank)
        return
    if not args.reads:
        args.reads = 10_000_000
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
