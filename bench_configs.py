"""bench_configs.py — the BASELINE.json configurations other than the headline one (bench.py --config 1|3|4|5).

Same JSON line as bench.py (value = device-resident reads/s, e2e through the C ABI from pinned host batches in the compact
transfer format, roofline of the dominant kernel, cpu_baseline = the oracle on a bounded sample of the same reads WITH the
device result of that sample compared field by field).  What differs per configuration is the plan:

  1  run_example.sh: VIC01 FASTA + Coordinates.csv (tests/golden copies of the reference's files), 100 k reads, --bin 10
     --breakThresh 1000 (scripts/run_example.sh:17-18); reads shard over the ranks when N > 1
  3  229E (WT_229_reference.fasta.gz + Coordinates.csv), 8 source BAMs x 5 M reads combined (scripts/run_slurm_229E.sh:37):
     every GPU holds 5 M reads; with 8 GPUs GPU k holds BAM k, with fewer the sources alternate inside a GPU's reads
  4  SARS-CoV-2 500 M-read stress run, STRONG scaling: 500 M / N reads per GPU, walked in waves (repeated npt_submit on one
     chromosome) of one HBM-resident 10 M-read batch, ranks taking turns in the global submit order
  5  human-host shape (scripts/opts_host.txt:1-25): 25 chromosomes with a synthetic exon table, spliced long reads, source 0
     = the transcripts themselves (firstIsTranscriptome); chromosomes are independent (a new cluster table per chromosome,
     ViralTranscriptAnalysisCmd2.java:885), so they are dealt round-robin to the GPUs and nothing is exchanged
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _plan(no, rank, world, args):
    from nptranscript_b200 import synth, workloads
    P = dict(no=no, collective=world > 1, scaling="weak", waves=1)
    if no == 1:
        w = workloads.sars2_vic01_real()
        total = args.reads or 100_000
        per = total // world
        P.update(name="configs[0] run_example.sh: VIC01 FASTA + Coordinates.csv, %d reads in all" % total, scaling="strong",
                 cfg_kw=dict(bin=10, break_thresh=1000, record_depth=0, calc_breaks=1, track_break_sums=1, num_sources=1),
                 chroms=[dict(index=0, genome=synth.genome_codes(w), annotation=w.annotation, n=per, first=rank * per,
                              gen=lambda n, first, alloc=None: synth.generate(w, 20200301, n, first_index=first, alloc=alloc))])
    elif no == 3:
        w = workloads.cov229e_real()
        per = args.reads or 5_000_000
        fixed = rank if world == 8 else -1
        P.update(name="configs[2] 229E, 8 source BAMs combined, %d reads per GPU (%s)" % (per, "GPU k holds BAM k" if world == 8 else "sources alternate"),
                 # bin 10 on 40 M reads: millions of sub-clusters and break-point pairs once every rank holds every rank's keys
                 cfg_kw=dict(bin=10, break_thresh=1000, record_depth=1, calc_breaks=1, track_break_sums=1, num_sources=8, max_coverage_clusters=1024,
                             max_subclusters=1 << 23, max_break_pairs=1 << 23),
                 chroms=[dict(index=0, genome=synth.genome_codes(w), annotation=w.annotation, n=per, first=rank * per,
                              gen=lambda n, first, alloc=None: synth.generate(w, 20200310, n, first_index=first, num_sources=8, src_fixed=fixed, alloc=alloc))])
    elif no == 4:
        w = workloads.SARS2
        total = args.reads or 500_000_000
        wave = min(10_000_000, total // world)
        waves = max(1, total // world // wave)
        P.update(name="configs[3] SARS-CoV-2 stress run, %d reads in all = %d waves of %d reads per GPU (one HBM-resident batch walked again per wave)"
                      % (waves * wave * world, waves, wave), scaling="strong", waves=waves,
                 cfg_kw=dict(bin=100, break_thresh=1000, record_depth=1, calc_breaks=1, track_break_sums=1, num_sources=1, max_coverage_clusters=2048),
                 chroms=[dict(index=0, genome=synth.genome_codes(w), annotation=w.annotation, n=wave, first=rank * wave,
                              gen=lambda n, first, alloc=None: synth.generate(w, 20200400, n, first_index=first, alloc=alloc))])
    elif no == 5:
        chs = workloads.human_like(scale=args.host_scale)
        per_chrom = args.reads or (2_000_000 if world >= 8 else 400_000)
        mine = [c for c in chs if c.index % world == rank]
        P.update(name="configs[4] human-host shape (opts_host.txt): 25 synthetic chromosomes (scale %.3g of GRCh38, %d exons), %d spliced reads per chromosome, "
                      "chromosomes dealt round-robin to the GPUs" % (args.host_scale, sum(len(c.exon_start) for c in chs), per_chrom),
                 collective=False,
                 cfg_kw=dict(bin=100, break_thresh=50, coronavirus=0, include_start=0, enforce_strand=0, record_depth=0, calc_breaks=1,
                             first_is_transcriptome=1, track_break_sums=1, num_sources=2, rna=[1, 1]),
                 chroms=[])
        for c in mine:
            g = synth.host_genome_codes(c)
            P["chroms"].append(dict(index=c.index, genome=g, annotation=c.annotation, n=per_chrom, first=0,
                                    gen=(lambda c, g: lambda n, first, alloc=None: synth.generate_host(c, g, 20200500 + c.index, n, first_index=first, num_sources=2,
                                                                                                   transcriptome=True, alloc=alloc))(c, g)))
    else:
        raise SystemExit("--config must be 1, 2, 3, 4 or 5")
    return P


def run_reference(args, rank, world):
    """CPU arm for these configurations: the oracle on a bounded sample of the plan's first chromosome (rank 0 only)."""
    if rank != 0:
        return
    import bench as B
    from oracle import oracle
    P = _plan(args.config, 0, 1, args)
    ch = P["chroms"][0]
    threads = os.cpu_count() or 1
    n = min(args.ref_reads, ch["n"])
    b = ch["gen"](n, ch["first"])
    cfg = oracle.make_config(**P["cfg_kw"])
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o = oracle.OracleRun(cfg, ch["genome"], ch["annotation"], ch["index"], keep_reads=True)
        o.submit(b, nthreads=threads)
        o.close()
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    v = n / (ms / 1e3)
    print(json.dumps({
        "impl": "reference", "metric": B.METRIC, "value": v, "unit": B.UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": P["scaling"], "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": P["name"] + f" — bounded sample: {n} reads of the first chromosome per step", "flags": P["cfg_kw"]},
        "cpu_baseline": {"value": v, "unit": B.UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} reads, oracle.cpp (C++ restatement of the Java path, not the JVM), per-read stage on {threads} threads, sequential commit"},
        "e2e": {"value": v, "unit": B.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)


def run(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import bench as B
    from nptranscript_b200 import abi, engine, packed as packed_mod

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.all_reduce(torch.zeros(1, device="cuda")); torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1); os.close(saved)
    P = _plan(args.config, rank, world, args)
    cfg = engine.make_config(**dict(P["cfg_kw"], device=local_rank))
    L = engine.lib()
    pinned = []

    def pinned_alloc(nbytes):
        import ctypes
        p = L.npt_alloc_pinned(nbytes)
        if not p:
            raise MemoryError("npt_alloc_pinned")
        pinned.append(p)
        return (ctypes.c_char * nbytes).from_address(p)

    # ---- reads of this rank: HBM-resident per chromosome; the first chromosome's first chunks also stay on the host (pinned, packed)
    t0 = time.perf_counter()
    in_bytes = cig_ops = bases = n_rank = 0
    host_chunks, sample = [], None
    for ci, ch in enumerate(P["chroms"]):
        parts = {k: [] for k in ("pos", "flag", "src", "cigar", "seq4")}
        coffs, soffs, cb, sb = [], [], 0, 0
        nchunks = max(1, (ch["n"] + B.CHUNK - 1) // B.CHUNK)
        for c in range(nchunks):
            m = min(B.CHUNK, ch["n"] - c * B.CHUNK)
            if m <= 0:
                break
            b = ch["gen"](m, ch["first"] + c * B.CHUNK)
            in_bytes += b.nbytes(); cig_ops += len(b.cigar); bases += 2 * len(b.seq4)
            if ci == 0 and c < args.e2e_chunks:
                host_chunks.append(packed_mod.pack(b, alloc=pinned_alloc))
            if ci == 0 and c == 0:
                sample = b.slice(0, min(args.cpu_reads, b.n))
            parts["pos"].append(torch.from_numpy(b.pos).cuda()); parts["flag"].append(torch.from_numpy(b.flag.view(np.int16)).cuda())
            parts["src"].append(torch.from_numpy(b.src.view(np.int16)).cuda()); parts["cigar"].append(torch.from_numpy(b.cigar.view(np.int32)).cuda())
            sq = torch.from_numpy(b.seq4).cuda()
            pad = (-sq.numel()) % 4
            if pad:
                sq = torch.cat([sq, torch.zeros(pad, dtype=torch.uint8, device="cuda")])
            parts["seq4"].append(sq)
            last = c == nchunks - 1
            coffs.append((b.cigar_off.astype(np.int64) + cb)[: None if last else -1]); soffs.append((b.seq_off.astype(np.int64) + sb)[: None if last else -1])
            cb += int(b.cigar_off[-1]); sb += int(b.seq_off[-1]) + pad
        d = {k: torch.cat(v) for k, v in parts.items()}
        d["cigar_off"] = torch.from_numpy(np.concatenate(coffs).astype(np.uint32).view(np.int32)).cuda()
        d["seq_off"] = torch.from_numpy(np.concatenate(soffs)).cuda()
        d["seq4"] = torch.cat([d["seq4"], torch.zeros(64, dtype=torch.uint8, device="cuda")])
        ch["dev"] = d
        n_rank += ch["n"] * P["waves"]
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0

    eng = engine.Engine(cfg)
    if world > 1 and P["collective"]:
        ident = [engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        saved = os.dup(1); os.dup2(2, 1)
        try:
            eng.comm_init(ident[0], rank, world)
        finally:
            os.dup2(saved, 1); os.close(saved)

    def device_step(fetch=False):
        tot = dict(step_ms=0.0, walk_ms=0.0, finalize_ms=0.0, total_launches=0, stages=np.zeros(6), kx=0.0, red=0.0, clusters=0, subs=0, slow=0, breaks=0)
        for ch in P["chroms"]:
            d = ch["dev"]
            eng.set_read_index_base(ch["first"])
            eng.begin_chrom(ch["index"], ch["genome"], ch["annotation"])
            for k in range(P["waves"]):
                if P["waves"] > 1:
                    eng.set_next_read_index((k * world + rank) * ch["n"])
                eng.submit_device(ch["n"], d["pos"].data_ptr(), d["flag"].data_ptr(), d["src"].data_ptr(), d["cigar_off"].data_ptr(),
                                  d["cigar"].data_ptr(), d["seq_off"].data_ptr(), d["seq4"].data_ptr())
            if world > 1 and P["collective"]:
                eng.end_chrom_all()
                kx, red = eng.exchange_ms()
                tot["kx"] += kx; tot["red"] += red
            else:
                eng.end_chrom()
            t = eng.kernel_ms()
            for k in ("step_ms", "walk_ms", "finalize_ms", "total_launches"):
                tot[k] += t[k]
            tot["stages"] += np.array([t["fastwalk_ms"], t["walk0_ms"], t["walk2_ms"], t["cluster_ms"], t["walk1_ms"], t["fastcov_ms"]])
            if fetch:
                r = eng.fetch_raw(abi.NPT_FETCH_TABLES)
                tot["clusters"] += r.n_clusters; tot["subs"] += r.n_subs; tot["slow"] += r.n_slow_reads
            eng.release()
        return tot

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        if world > 1:
            dist.barrier()
        device_step()
    sampler = B.ClockSampler(local_rank if rank == 0 else None)
    barrier()
    sampler.start()
    steps = []
    wall0 = time.perf_counter()
    for i in range(args.steps):
        if world > 1:
            dist.barrier()
        steps.append(device_step(fetch=(i == args.steps - 1)))
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0) / args.steps
    clocks = sampler.stop()
    ms = float(np.mean([s["step_ms"] for s in steps]))
    n_all = n_rank
    if world > 1:
        tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        nn = torch.tensor([n_rank], device="cuda", dtype=torch.int64); dist.all_reduce(nn); n_all = int(nn.item())
    value = n_all / (ms / 1e3)
    last = steps[-1]

    # ---- end to end: the first chromosome's pinned packed chunks through npt_submit_packed, everything fetched back
    ch0 = P["chroms"][0]
    e2e_reads = sum(pb.n for pb in host_chunks)
    h2d = sum(pb.nbytes() for pb in host_chunks)

    def e2e_step():
        eng.set_read_index_base(ch0["first"])
        eng.begin_chrom(ch0["index"], ch0["genome"], ch0["annotation"])
        for pb in host_chunks:
            eng.submit_packed(pb)
        if world > 1 and P["collective"]:
            eng.end_chrom_all()
        else:
            eng.end_chrom()
        r = eng.fetch_raw(abi.NPT_FETCH_ALL)
        nb = int(np.ctypeslib.as_array(r.break_off, shape=(e2e_reads + 1,))[-1])
        d2h = e2e_reads * (9 * 4 + 8) + nb * 4 + (4 * r.n_clusters * cfg.num_sources * r.cov_stride * 4 if P["cfg_kw"].get("record_depth") else 0)
        eng.release()
        return d2h

    e2e_ms, d2h = [], 0
    for it in range(1 + args.e2e_steps):
        barrier()
        t1 = time.perf_counter()
        d2h = e2e_step()
        torch.cuda.synchronize()
        if it >= 1:
            e2e_ms.append(1e3 * (time.perf_counter() - t1))
    e2e_t = float(np.mean(e2e_ms))
    e2e_all = e2e_reads
    if world > 1:
        tt = torch.tensor([e2e_t], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_t = float(tt.item())
        nn = torch.tensor([e2e_reads], device="cuda", dtype=torch.int64); dist.all_reduce(nn); e2e_all = int(nn.item())

    # ---- CPU baseline + parity (rank 0, N = 1): the oracle on the first reads of the first chromosome, device result compared
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle
        from tests.common import assert_same
        threads = os.cpu_count() or 1
        okw = {k: v for k, v in P["cfg_kw"].items()}
        t1 = time.perf_counter()
        o = oracle.OracleRun(oracle.make_config(**okw), ch0["genome"], ch0["annotation"], ch0["index"], keep_reads=True)
        o.submit(sample, nthreads=threads)
        dt = time.perf_counter() - t1
        ora = o.results(coverage=bool(P["cfg_kw"].get("record_depth")))
        o.close()
        eng.set_read_index_base(0)
        eng.begin_chrom(ch0["index"], ch0["genome"], ch0["annotation"])
        eng.submit(sample)
        eng.end_chrom()
        dev = eng.fetch()
        eng.release()
        assert_same(dev, ora)
        cpu = {"value": sample.n / dt, "unit": B.UNIT, "cores": threads, "kind": "port", "parity": f"device == oracle on all fields of these {sample.n} reads",
               "sample": f"first {sample.n} reads of the first chromosome; oracle.cpp (C++ restatement, not the JVM), per-read stage on {threads} threads + sequential commit, {dt:.1f} s"}

    if rank == 0:
        peak, peak_kind = B.peaks()
        n_local = n_rank
        alg = (in_bytes * P["waves"]) + 40 * n_local   # every input byte once per walk + 40 B of per-read output (break lists not counted here)
        st = np.mean(np.array([s["stages"] for s in steps]), axis=0)
        keys6 = ["fastwalk", "walk0", "walk2", "cluster", "walk1", "fastcov"]
        names = ["fast_walk_kernel", "walk_kernel<0> (exact serial walker)", "walk_kernel<2>", "cluster_kernel", "walk_kernel<1>", "cov_key_kernel + radix sort + accumulate_kernel"]
        dom = int(np.argmax(st))
        wk = float(np.mean([s["walk_ms"] for s in steps]))
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": P["scaling"], "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": P["name"], "flags": P["cfg_kw"], "reads_per_step_all_gpus": n_all, "chromosomes_on_rank0": len(P["chroms"]),
                       "mean_read_bases": bases / max(1, sum(c["n"] for c in P["chroms"])), "mean_cigar_ops": cig_ops / max(1, sum(c["n"] for c in P["chroms"])),
                       "input_bytes_per_read": in_bytes / max(1, sum(c["n"] for c in P["chroms"])),
                       "l2_policy": "inputs of a step (%.2f GB on rank 0) %s" % (in_bytes / 1e9, "larger than L2; no flush needed" if in_bytes > 2.6e8 else
                                                                                  "smaller than L2: a 512 MB buffer is not flushed between steps (the whole workload is L2-sized by definition)"),
                       "clusters": last["clusters"], "sub_clusters": last["subs"], "reads_on_serial_path": last["slow"], "gen_seconds": gen_s,
                       "merge_breakdown_ms": {"key_exchange_ms": last["kx"], "reduce_ms": last["red"]} if world > 1 and P["collective"] else None,
                       "timing": "CUDA events inside libnpt on its compute stream, summed over this rank's chromosomes, max over ranks; wall ms/step %.2f" % wall_ms},
            "roofline": {"bound": "hbm", "achieved": alg / (float(st[dom]) / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (float(st[dom]) / 1e3) / 1e9 / peak, "traffic": None, "kernel": names[dom], "kernel_ms": float(st[dom]),
                         "kernel_share_of_step": float(st[dom]) / ms, "algorithmic_bytes": alg, "peak_source": peak_kind,
                         "kernels_ms": dict({k: float(v) for k, v in zip(keys6, st)}, finalize=float(np.mean([s["finalize_ms"] for s in steps]))),
                         "pipeline": {"ms": wk, "achieved": alg / (wk / 1e3) / 1e9, "frac": alg / (wk / 1e3) / 1e9 / peak}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_all / (e2e_t / 1e3), "unit": B.UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "reads_per_step": e2e_reads,
                    "ms_per_step": e2e_t, "note": "first chromosome's pinned host chunks (compact transfer format) through npt_submit_packed, end of chromosome, npt_fetch_results(ALL)"},
            "gpu_launches": int(sum(s["total_launches"] for s in steps)),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    for p in pinned:
        L.npt_free_pinned(p)
    if world > 1:
        dist.destroy_process_group()
