"""N>1 path on CPU: two gloo ranks each run the oracle on their shard of the reads, merge tables by key
(nptranscript_b200.multigpu), permute and all-reduce the coverage rows, and rank 0 checks everything against the oracle on
the unsharded input — ids per read, tables, break-point pairs, ORF counters, coverage."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(bin=10, break_thresh=1000, record_depth=1, num_sources=2)
N = 6000


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    from nptranscript_b200 import multigpu, synth, workloads
    from oracle import oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = workloads.SARS2
    g = synth.genome_codes(w)
    per = N // world
    b = synth.generate(w, 77, per, first_index=rank * per, num_sources=2, nthreads=2)
    res = oracle.run(oracle.make_config(**KW), g, w.annotation, [b], nthreads=1)
    gathered = [None] * world
    dist.all_gather_object(gathered, multigpu.local_tables(res, idx_base=rank * per))
    glob, maps = multigpu.merge_tables(gathered)            # vectorised numpy implementation
    cmap, smap = maps[rank]
    cl, sb = multigpu.remap_reads(res, cmap, smap)
    cov = {}
    for k in ("depth", "errors", "map_start", "map_end"):
        t = torch.from_numpy(multigpu.permute_rows(res[k], cmap, glob["n_clusters"]))
        dist.all_reduce(t)
        cov[k] = t.numpy()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), cluster_id=cl, sub_id=sb)
    if rank == 0:
        np.savez(os.path.join(outdir, "global.npz"), **{k: v for k, v in glob.items() if isinstance(v, np.ndarray)}, **cov)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_matches_unsharded_oracle():
    from nptranscript_b200 import synth, workloads
    from oracle import oracle
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, 29611, d), nprocs=world, join=True)
        glob = dict(np.load(os.path.join(d, "global.npz")))
        parts = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]
    w = workloads.SARS2
    g = synth.genome_codes(w)
    full = synth.generate(w, 77, N, num_sources=2, nthreads=2)
    ref = oracle.run(oracle.make_config(**KW), g, w.annotation, [full])
    assert np.array_equal(np.concatenate([p["cluster_id"] for p in parts]), ref["cluster_id"])
    assert np.array_equal(np.concatenate([p["sub_id"] for p in parts]), ref["sub_id"])
    for k in ("cl_first_read", "cl_key_off", "cl_key_tok", "cl_read_count", "cl_sub_off", "sub_first_read", "sub_key_off", "sub_key", "sub_count",
              "sub_break_sum", "sub_sum_off", "sub_sums", "sub_flags", "total_counts", "spliced", "unspliced", "pair_src", "pair_level",
              "pair_start", "pair_end", "pair_count", "depth", "errors", "map_start", "map_end"):
        assert np.array_equal(glob[k], ref[k]), k
