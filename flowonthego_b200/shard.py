"""Sharding of independent frame pairs over the GPUs of one node.

A frame pair depends on no other pair (the engine is stateless, `initflow` is unused by the CLI), so
pair k of an n-pair stream simply goes to rank k*world/n (contiguous blocks; each rank needs one
extra frame).  There is no collective on the data path; torch.distributed is used only to gather
results (or their summaries) on rank 0 -- NCCL on GPUs, gloo in the CPU tests.
"""
import numpy as np


def partition(n_pairs, world, rank):
    """Contiguous block [k0, k1) of pair indices owned by `rank`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, rem = divmod(n_pairs, world)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def frames_needed(n_pairs, world, rank):
    """Frame indices [f0, f1) a rank must hold: its pairs (k, k+1) share frames."""
    k0, k1 = partition(n_pairs, world, rank)
    return (k0, k1 + 1) if k1 > k0 else (k0, k0)


def run_shard(frames, first_pair, engine_fn):
    """Runs engine_fn(frame_k, frame_k+1) -> flow for every consecutive pair in `frames`.
    Returns {global pair index: flow}."""
    return {first_pair + i: engine_fn(frames[i], frames[i + 1]) for i in range(len(frames) - 1)}


def gather_summaries(local, n_pairs, dist=None, device=None):
    """Gathers one float32 vector per pair (e.g. mean u, mean v, mean |flow|) on rank 0.
    `local` = {pair index: 1-D array of length m}.  Returns an (n_pairs, m) array on rank 0, None elsewhere."""
    import torch
    m = len(next(iter(local.values()))) if local else 0
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = np.full((n_pairs, m), np.nan, np.float32)
        for k, v in local.items():
            out[k] = v
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    ms = torch.tensor([m], dtype=torch.int64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    m = int(ms.item())
    biggest = max(partition(n_pairs, world, r)[1] - partition(n_pairs, world, r)[0] for r in range(world))
    buf = torch.full((biggest, m + 1), float("nan"), dtype=torch.float32, device=device)
    for i, (k, v) in enumerate(sorted(local.items())):
        buf[i, 0] = float(k)
        buf[i, 1:] = torch.as_tensor(np.asarray(v, np.float32), device=device)
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, parts, dst=0)
    if rank != 0:
        return None
    out = np.full((n_pairs, m), np.nan, np.float32)
    for p in parts:
        p = p.cpu().numpy()
        for row in p:
            if not np.isnan(row[0]):
                out[int(row[0])] = row[1:]
    return out
