"""Run under torchrun on N GPUs: every rank walks its shard on its GPU, table keys are exchanged over NCCL and inserted on
the device, counters are summed with NCCL all-reduces; rank 0 checks ids, tables and coverage against the oracle on the
unsharded input.   torchrun --nproc-per-node 2 tests/multigpu_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from nptranscript_b200 import abi, engine, multigpu, synth, workloads

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
KW = dict(bin=10, break_thresh=1000, record_depth=1, num_sources=2)
N = 40000
per = N // world
w = workloads.SARS2
g = synth.genome_codes(w)
b = synth.generate(w, 78, per, first_index=rank * per, num_sources=2)
eng = engine.Engine(engine.make_config(device=local, **KW))
eng.set_read_index_base(rank * per)
eng.begin_chrom(0, g, w.annotation)
eng.submit(b)
multigpu.exchange_keys(eng, rank, world)
eng.end_chrom()
multigpu.reduce_aggregates(eng)
glob = eng.fetch(abi.NPT_FETCH_ALL)   # global tables and coverage on every rank; per-read ids are global ids of the local reads
cl, sb = glob["cluster_id"], glob["sub_id"]
cov = np.stack([glob[k] for k in ("depth", "errors", "map_start", "map_end")])
small = np.concatenate([glob["total_counts"].ravel(), glob["spliced"].ravel(), glob["unspliced"].ravel()])
ids = [None] * world
dist.all_gather_object(ids, (cl, sb))
if rank == 0:
    from oracle import oracle
    full = synth.generate(w, 78, N, num_sources=2)
    ref = oracle.run(oracle.make_config(**KW), g, w.annotation, [full])
    assert np.array_equal(np.concatenate([i[0] for i in ids]), ref["cluster_id"])
    assert np.array_equal(np.concatenate([i[1] for i in ids]), ref["sub_id"])
    for k in ("cl_first_read", "cl_key_off", "cl_key_tok", "cl_read_count", "cl_sub_off", "sub_first_read", "sub_key_off", "sub_key", "sub_count",
              "sub_break_sum", "sub_sum_off", "sub_sums", "sub_flags", "pair_src", "pair_level", "pair_start", "pair_end", "pair_count"):
        assert np.array_equal(glob[k], ref[k]), k
    for i, k in enumerate(("depth", "errors", "map_start", "map_end")):
        assert np.array_equal(cov[i], ref[k]), k
    S, no = 2, len(w.annotation.start)
    assert np.array_equal(small[:S], ref["total_counts"]) and np.array_equal(small[S:S + S * no].reshape(S, no), ref["spliced"])
    print(f"multigpu_check ok: world={world} clusters={glob['n_clusters']} subs={glob['n_subs']}")
eng.close()
dist.destroy_process_group()
