"""Platform ceiling of host-to-device ingest with N processes (one per GPU): a bare loop of pinned cudaMemcpyAsync copies of
the size npt_submit_packed moves per chunk, no library involved.   torchrun --nproc-per-node N profiles/h2d_ceiling.py
Prints per-GPU and aggregate GB/s (max time over ranks), for comparison with the e2e leg of bench.py at the same N."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
CHUNK = 900 << 20          # one 1 Mi-read chunk in the compact format is ~0.9 GB
host = [torch.empty(CHUNK, dtype=torch.uint8).pin_memory() for _ in range(2)]
dev = [torch.empty(CHUNK, dtype=torch.uint8, device="cuda") for _ in range(2)]
s = torch.cuda.Stream()


def run(n):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for i in range(n):
            dev[i & 1].copy_(host[i & 1], non_blocking=True)
    s.synchronize()
    return time.perf_counter() - t0


run(2)
dt = run(8)
t = torch.tensor([dt], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    gb = 8 * CHUNK / 1e9
    print(f"h2d_ceiling: {world} GPU(s), {gb:.1f} GB per GPU in {t.item() * 1e3:.1f} ms -> {gb / t.item():.1f} GB/s per GPU, {world * gb / t.item():.1f} GB/s aggregate")
if world > 1:
    dist.destroy_process_group()
