"""Multi-GPU checks of the collectives INSIDE libnpt (csrc/comm.inl) against the oracle on the unsharded input.

  python tests/mg_check.py --devices 0,1[,2,...]        one process: npt_mg_* (one context + host thread per device, ncclCommInitAll);
                                                         no torch.distributed anywhere
  torchrun --nproc-per-node N tests/mg_check.py          one process per GPU: npt_comm_init + npt_end_chrom_all; torch.distributed only
                                                         carries the 128-byte NCCL id and gathers the per-read ids for the check
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nptranscript_b200 import abi, engine, synth, workloads  # noqa: E402

KW = dict(bin=10, break_thresh=1000, record_depth=1, num_sources=2)
N = 40000
TABLES = ("cl_first_read", "cl_key_off", "cl_key_tok", "cl_read_count", "cl_sub_off", "sub_first_read", "sub_key_off", "sub_key", "sub_count",
          "sub_break_sum", "sub_sum_off", "sub_sums", "sub_flags", "pair_src", "pair_level", "pair_start", "pair_end", "pair_count",
          "total_counts", "spliced", "unspliced", "depth", "errors", "map_start", "map_end")


def reference(w, g):
    from oracle import oracle
    full = synth.generate(w, 78, N, num_sources=2)
    return full, oracle.run(oracle.make_config(**KW), g, w.annotation, [full])


def inproc(devices):
    from tests.common import assert_same
    w = workloads.SARS2
    g = synth.genome_codes(w)
    full, ref = reference(w, g)
    batches = [full.slice(0, 25000), full.slice(25000, 25001), full.slice(25001, N)]
    dev = engine.run_mg(engine.make_config(**KW), devices, g, w.annotation, batches)
    assert_same(dev, ref)
    print(f"mg_check ok (in-process, devices {devices}): clusters={dev['n_clusters']} subs={dev['n_subs']} timing={dev['_timing']}")


def per_rank():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")   # plumbing only: the data path's collectives are libnpt's own NCCL communicator
    ident = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    per = N // world
    w = workloads.SARS2
    g = synth.genome_codes(w)
    b = synth.generate(w, 78, per, first_index=rank * per, num_sources=2)
    eng = engine.Engine(engine.make_config(device=local, **KW))
    eng.comm_init(ident[0], rank, world)
    eng.set_read_index_base(rank * per)
    eng.begin_chrom(0, g, w.annotation)
    eng.submit(b)
    eng.end_chrom_all()
    glob = eng.fetch(abi.NPT_FETCH_ALL)
    ids = [None] * world
    dist.all_gather_object(ids, (glob["cluster_id"], glob["sub_id"]))
    if rank == 0:
        _, ref = reference(w, g)
        assert np.array_equal(np.concatenate([i[0] for i in ids]), ref["cluster_id"])
        assert np.array_equal(np.concatenate([i[1] for i in ids]), ref["sub_id"])
        for k in TABLES:
            assert np.array_equal(glob[k], ref[k]), k
        print(f"mg_check ok (npt_comm_init + npt_end_chrom_all, world={world}): clusters={glob['n_clusters']} subs={glob['n_subs']} "
              f"exchange/reduce ms={eng.exchange_ms()}")
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="")
    a = ap.parse_args()
    if "RANK" in os.environ:
        per_rank()
    else:
        inproc([int(x) for x in a.devices.split(",")] if a.devices else [0, 0])
