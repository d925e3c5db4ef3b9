"""Aggregate pinned host<->device bandwidth with one process per GPU running concurrently (torchrun)."""
import os, time, torch
import torch.distributed as dist
rank = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dist.init_process_group("gloo")
n = 64 << 20
h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
ss = [torch.cuda.Stream() for _ in range(4)]
for direction in ("d2h", "h2d", "both"):
    def go():
        for hh, dd, s in zip(h, d, ss):
            with torch.cuda.stream(s):
                if direction in ("h2d", "both"): dd.copy_(hh, non_blocking=True)
                if direction in ("d2h", "both"): hh.copy_(dd, non_blocking=True)
    go(); torch.cuda.synchronize(); dist.barrier()
    t = time.perf_counter()
    for _ in range(20): go()
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t
    gb = n * 4 * 20 * (2 if direction == "both" else 1) / dt / 1e9
    tot = torch.tensor([gb], dtype=torch.float64); dist.all_reduce(tot)
    if rank == 0: print("%d GPUs %s: %.1f GB/s per GPU (rank 0), %.1f GB/s aggregate" % (world, direction, gb, tot.item()), flush=True)
