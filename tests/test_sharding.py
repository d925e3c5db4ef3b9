"""CPU tests of the multi-GPU host logic: pair partitioning and the result gather, with
world_size 2 over gloo (the GPU path uses the same code over NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

from flowonthego_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_every_pair_once():
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                k0, k1 = shard.partition(n, world, r)
                assert 0 <= k0 <= k1 <= n
                seen += list(range(k0, k1))
                f0, f1 = shard.frames_needed(n, world, r)
                assert (f1 - f0) == ((k1 - k0 + 1) if k1 > k0 else 0)
            assert seen == list(range(n))
    sizes = [shard.partition(1024, 8, r)[1] - shard.partition(1024, 8, r)[0] for r in range(8)]
    assert sizes == [128] * 8  # C5: 1024 pairs over 8 GPUs


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import port as oracle
    from tests.synth import texture, warp, affine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_pairs, w, h = 5, 96, 64
    base = texture(w, h, 7)
    frames = [warp(base, affine(w, h, rot_deg=0.2 * k, scale=1.0, shift=(0.6 * k, 0.0))) for k in range(n_pairs + 1)]
    pd = dict(lv_f=2, lv_l=1, maxiter=8, miniter=8, mindprate=0.05, mindrrate=0.95, minimgerr=0.0, patchsz=8, poverl=0.4,
              usefbcon=0, patnorm=1, costfct=0, usetvref=1, tv_alpha=10.0, tv_gamma=10.0, tv_delta=5.0, tv_innerit=1,
              tv_solverit=3, tv_sor=1.6, verbosity=0)
    f0, f1 = shard.frames_needed(n_pairs, world, rank)
    # the CPU oracle stands in for the GPU engine: this test is about the sharding logic only
    local = shard.run_shard(frames[f0:f1], f0, lambda a, b: oracle.run_u8(a, b, pd))
    summ = {k: np.array([v[..., 0].mean(), v[..., 1].mean()], np.float32) for k, v in local.items()}
    out = shard.gather_summaries(summ, n_pairs, dist)
    if rank == 0:
        ref = np.stack([[f[..., 0].mean(), f[..., 1].mean()] for f in
                        (oracle.run_u8(frames[k], frames[k + 1], pd) for k in range(n_pairs))]).astype(np.float32)
        q.put((out.tolist(), ref.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_matches_serial_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, ref = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(np.asarray(out, np.float32), np.asarray(ref, np.float32))
