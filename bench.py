#!/usr/bin/env python
"""bench.py -- headline benchmark of the DIS optical-flow hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.md C5/C3 -- a synthetic 1080p video stream (affine triangle-wave
trajectory of a procedural texture), reference operating point 3 (patchsz 12, overlap 0.75, lv 6->2,
16 Gauss-Newton iterations) + variational refinement.  One step = one pass of the whole hot path
(pyramid -> inverse search -> densify -> refine -> upsample) over a batch of `--batch` consecutive
frame pairs per GPU.  Independent pairs shard across ranks with no data-path collective (weak
scaling: the per-GPU batch is fixed); torch.distributed (NCCL) is used only for the barrier and the
max-over-ranks of the device time.

  value : pairs/s with the frames already resident in HBM, results left in HBM; batched handles
          (dis_submit_u8_device_batch, --pairs-per-launch pairs per kernel launch, --batch-handles in flight)
  e2e   : pairs/s through the reference-facing C-ABI call dis_submit_u8/dis_wait with pinned HOST
          buffers (H2D of both frames and D2H of the full-resolution flow inside the timed region),
          one pair per call on --streams single-pair handles
  extras (N=1 unless noted): 4K pair (C4a), the video front end, engine-level output (all N)
  roofline / cpu_baseline : see DESIGN.md; the CPU leg is the reference's own engine (oracle/_ref,
          compiled verbatim) or, if that was not built, the C restatement (oracle/).

--impl reference times the reference's CPU implementation on the host cores for the same workload.
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

W1080, H1080 = 1920, 1080
METRIC, UNIT = "frame_pairs_per_sec", "pairs/s"
WORKLOAD = "C5/C3: 1080p synthetic stream, operating point 3 (p12 ov0.75 lv6->2 16it) + variational refinement"


# ------------------------------------------------------------------------------------------ data
def c5_matrix(t, w, h):
    """rotation 0.05 deg*t about the centre x scale 1+1e-4 t + shift (0.9t,-0.4t)."""
    from tests.synth import affine
    return affine(w, h, rot_deg=0.05 * t, scale=1.0 + 1e-4 * t, shift=(0.9 * t, -0.4 * t))


def c5_frames(w, h, n_frames, seed=1234):
    from tests.synth import texture, warp
    base = texture(w, h, seed)
    frames = np.empty((n_frames, h, w), np.uint8)
    for k in range(n_frames):
        t = k % 64 if (k // 64) % 2 == 0 else 64 - (k % 64)
        frames[k] = warp(base, c5_matrix(t, w, h))
    return frames


def alg_bytes(W, H, ps, ov, lvf, lvl, tv, tv_innerit=1, tv_solverit=3):
    """Algorithmic bytes per pair -- SURVEY.md Appendix A, verbatim."""
    import math
    sc = 2 ** lvf
    Wp = W + (sc - W % sc) % sc
    Hp = H + (sc - H % sc) % sc
    steps = max(1, int(math.floor(ps * (1 - ov))))
    B = 2 * W * H
    for l in range(lvf + 1):
        Np = ((Wp >> l) + 2 * ps) * ((Hp >> l) + 2 * ps)
        B += 8 * Np
        if l >= lvl:
            B += 8 * Np
    for l in range(lvf, lvl - 1, -1):
        w, h = Wp >> l, Hp >> l
        N = w * h
        Np = (w + 2 * ps) * (h + 2 * ps)
        npch = math.ceil(w / steps) * math.ceil(h / steps)
        B += 16 * Np + (8 * (N // 4) if l < lvf else 0) + 16 * npch
        B += 16 * npch + 8 * Np + 8 * N
        if tv:
            inner = tv_innerit * (l + 1)
            B += N * (28 + 40)
            B += N * inner * (16 + 64 + 32 + tv_solverit * 44 + 24)
            B += 8 * N
    return B


# ------------------------------------------------------------------------------- CPU reference
def _cpu_worker(args):
    """One process: `n` frame pairs through the reference CPU path (pyramids included), single thread."""
    kind, frames, pd, n = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        pass
    t0 = time.perf_counter()
    if kind == "reference":
        from oracle import ref_driver
        for i in range(n):
            ref_driver.run_dense_ref(frames[i % (len(frames) - 1)], frames[i % (len(frames) - 1) + 1], pd)
    else:
        from oracle import port
        for i in range(n):
            port.run_u8(frames[i % (len(frames) - 1)], frames[i % (len(frames) - 1) + 1], pd)
    return time.perf_counter() - t0


def cpu_reference_throughput(frames, pd, pairs_per_proc=2, max_procs=None):
    """pairs/s of the reference CPU implementation using every host core (independent single-thread
    processes over the pair list -- the reference itself is single-threaded, kroeger/CMakeLists.txt:28-33)."""
    import multiprocessing as mp
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdis_ref.so")) else "port"
    if kind == "port":
        from oracle import port
        port.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if max_procs:
        cores = min(cores, max_procs)
    ctx = mp.get_context("spawn")  # cv2 threads make fork unsafe
    jobs = [(kind, frames, pd, pairs_per_proc)] * cores
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        per = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    total = cores * pairs_per_proc
    return dict(value=total / wall, unit=UNIT, cores=cores, kind=kind,
                sample="%d pairs of the workload (%d per process x %d single-thread processes), %.1f s wall, "
                       "%.1f CPU-s; %.0f ms/pair single-thread" % (total, pairs_per_proc, cores, wall, sum(per),
                                                                    1e3 * sum(per) / total)), wall, total


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="frame pairs per step per GPU")
    ap.add_argument("--streams", type=int, default=64, help="engine instances (CUDA streams) per GPU")
    ap.add_argument("--pairs-per-launch", type=int, default=8,
                    help="batched handles for the device-resident arm (dis_create_batch); 1 = one pair per launch")
    ap.add_argument("--batch-handles", type=int, default=32, help="number of batched handles per GPU")
    ap.add_argument("--no-extra", action="store_true", help="skip the 4K (C4a) side measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N, K, Wm, B = args.gpus, args.steps, max(args.warmup, 0), args.batch

    import flowonthego_b200 as F
    pd = F.Params.preset(3, W1080, verbosity=0).to_dict()
    config = {"workload": WORKLOAD, "resolution": [W1080, H1080], "pairs_per_step_per_gpu": B,
              "params": "6 2 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0", "sharding": "independent pairs, dp%d" % N}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        frames = c5_frames(W1080, H1080, 9)
        cores = len(os.sched_getaffinity(0))
        ppp = 1 if cores >= 16 else 2
        # W warm-up + K timed steps, each step a bounded sample (ppp pairs per core)
        for _ in range(min(Wm, 1)):
            cpu_reference_throughput(frames, pd, pairs_per_proc=1)
        walls, tot = [], 0
        cb = None
        for _ in range(min(K, 3)):
            cb, wall, total = cpu_reference_throughput(frames, pd, pairs_per_proc=ppp)
            walls.append(wall)
            tot += total
        v = tot / sum(walls)
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": N,
                          "steps": min(K, 3), "warmup": min(Wm, 1), "ms_per_step": 1e3 * sum(walls) / len(walls),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------------------------------------------------------- product arm (B200)
    cpu_base = None
    frames = c5_frames(W1080, H1080, B + 1)
    if N == 1 and rank == 0:  # before CUDA is initialised (fork-safe)
        cpu_base, _, _ = cpu_reference_throughput(frames[:9], pd, pairs_per_proc=2 if os.cpu_count() < 16 else 1)

    import torch
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    S = max(1, min(args.streams, B))
    p = F.Params.from_dict(pd)
    engines = [F.Engine(p, W1080, H1080, local_rank) for _ in range(S)]  # one pair per launch: e2e legs, profile
    nb = max(1, min(args.pairs_per_launch, 8, B))
    Sb = max(1, args.batch_handles)
    bengines = [F.Engine(p, W1080, H1080, local_rank, batch=nb) for _ in range(Sb)] if nb > 1 else []
    streams = [torch.cuda.ExternalStream(e.stream, device=dev) for e in engines + bengines]
    d_frames = torch.from_numpy(frames).to(dev)
    d_out = torch.empty((S, H1080, W1080, 2), dtype=torch.float32, device=dev)
    fptr = [d_frames[i].data_ptr() for i in range(B + 1)]
    optr = [d_out[i].data_ptr() for i in range(S)]
    main_stream = torch.cuda.current_stream()

    # device-resident arm: nb pairs per launch on batched handles (same kernels, 1/nb of the launches per pair)
    d_out_b = torch.empty((Sb, nb, H1080, W1080, 2), dtype=torch.float32, device=dev) if nb > 1 else None
    chunks = [list(range(c, min(c + nb, B))) for c in range(0, B, nb)]
    turn = [0]

    def step_device():
        if nb == 1:
            for i in range(B):
                engines[i % S].submit_u8_device(fptr[i], fptr[i + 1], W1080, H1080, W1080, optr[i % S])
            return
        for idx in chunks:
            k = turn[0] % Sb
            turn[0] += 1
            bengines[k].submit_u8_device_batch([fptr[i] for i in idx], [fptr[i + 1] for i in idx], W1080, H1080, W1080,
                                               [d_out_b[k, j].data_ptr() for j in range(len(idx))])

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_stream)
        for s in streams:
            s.wait_event(e0)
        for _ in range(steps):
            fn()
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main_stream.wait_event(ev)
        e1.record(main_stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    for _ in range(max(Wm, 1)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, K)
    clocks = sampler.stop() if rank == 0 else None
    value = N * B * K / (ms / 1e3)
    launches_per_call = (bengines[0] if nb > 1 else engines[0]).timings()["launches"]
    launches_per_pair = launches_per_call / nb
    if nb > 1:  # free the batched workspaces before the host-buffer legs allocate theirs
        for e in bengines:
            e.wait()

    # ---- e2e: reference-facing call with host buffers, copies inside the timed region
    h_frames = F.pinned_empty(frames.shape, np.uint8)
    h_frames[...] = frames
    h_out = [F.pinned_empty((H1080, W1080, 2), np.float32) for _ in range(S)]

    def step_host():
        for i in range(B):
            e = engines[i % S]
            if i >= S:
                e.wait()
            e.submit_u8(h_frames[i], h_frames[i + 1], h_out[i % S])
        for e in engines:
            e.wait()

    step_host()
    Ke = max(1, min(K, 4))
    ms_e = timed(lambda: step_host(), Ke)
    e2e_value = N * B * Ke / (ms_e / 1e3)

    # ---- side measurement: host buffers in, the ENGINE's output out (DIS_OPT_LEVEL_OUTPUT: level-lv_l flow as the
    # OFC::OFClass constructor delivers it; the x4 resize + crop of run_dense.cpp:407-414 is left to the caller)
    level_extra = None
    if not args.no_extra:
        try:
            from flowonthego_b200 import api as _api
            wp_, hp_, _, _ = F.padded_size(W1080, H1080, p.lv_f)
            lshape = (hp_ >> p.lv_l, wp_ >> p.lv_l, 2)
            h_lvl = [F.pinned_empty(lshape, np.float32) for _ in range(S)]
            for e in engines:
                e.set_option(_api.OPT_LEVEL_OUTPUT, 1)

            def step_host_level():
                for i in range(B):
                    e = engines[i % S]
                    if i >= S:
                        e.wait()
                    e.submit_u8(h_frames[i], h_frames[i + 1], h_lvl[i % S])
                for e in engines:
                    e.wait()

            step_host_level()
            ms_l = timed(step_host_level, Ke)
            for e in engines:
                e.set_option(_api.OPT_LEVEL_OUTPUT, 0)
            level_extra = {"value": N * B * Ke / (ms_l / 1e3), "unit": UNIT, "h2d_bytes_per_step": 2 * W1080 * H1080 * B,
                           "d2h_bytes_per_step": int(np.prod(lshape)) * 4 * B,
                           "what": "dis_submit_u8 with DIS_OPT_LEVEL_OUTPUT=1: u8 frames in, level-%d flow %dx%d out "
                                   "(the OFClass output); device-timed like e2e" % (p.lv_l, lshape[1], lshape[0])}
        except Exception as ex:
            level_extra = {"error": repr(ex)}

    # ---- side measurement: the same host-buffer workload through the video front end (dis_video_*: every frame
    # uploaded once, pairs pipelined over S handles)
    stream_extra = None
    if rank == 0 and N == 1 and not args.no_extra:
        try:
            import ctypes
            L = F.lib()
            vid = ctypes.c_void_p()
            rc = L.dis_video_create(ctypes.byref(p), 1, W1080, H1080, local_rank, S, ctypes.byref(vid))
            if rc != 0:
                raise RuntimeError(L.dis_last_error(None).decode())
            fp = ctypes.POINTER(ctypes.c_float)

            def step_video():
                L.dis_video_push(vid, h_frames[0].ctypes.data, W1080, None)
                for i in range(B):
                    if L.dis_video_pending(vid) >= S:
                        L.dis_video_pop(vid, None)
                    if L.dis_video_push(vid, h_frames[i + 1].ctypes.data, W1080, h_out[i % S].ctypes.data_as(fp)) != 0:
                        raise RuntimeError(L.dis_last_error(None).decode())
                while L.dis_video_pending(vid) > 0:
                    L.dis_video_pop(vid, None)

            step_video()
            t0 = time.perf_counter()
            for _ in range(Ke):
                step_video()
            dt = time.perf_counter() - t0
            L.dis_video_destroy(vid)
            stream_extra = {"value": B * Ke / dt, "unit": UNIT, "h2d_bytes_per_step": W1080 * H1080 * (B + 1),
                            "d2h_bytes_per_step": 8 * W1080 * H1080 * B, "pairs_in_flight": S,
                            "how": "dis_video_push/pop with pinned host frames and flow buffers, host wall clock"}
        except Exception as ex:
            stream_extra = {"error": repr(ex)}

    # ---- results are only *gathered* over NCCL (no data-path collective): per-pair flow summaries of the
    # last S pairs of every rank go to rank 0
    from flowonthego_b200 import shard
    summ_local = {}
    outs_flat = d_out_b.view(-1, H1080, W1080, 2) if nb > 1 else d_out
    for i in range(S):
        f = outs_flat[i % outs_flat.shape[0]]
        summ_local[rank * S + i] = np.array([float(f[..., 0].mean()), float(f[..., 1].mean())], np.float32)
    summaries = shard.gather_summaries(summ_local, world * S, dist, dev)

    # ---- roofline of the dominant kernel: per-kernel CUDA-event times (separate, un-graphed pass)
    roof, per_kernel = None, {}
    if rank == 0:
        e = engines[0]
        e.enable_kernel_profile(True)
        npairs = 4
        for i in range(npairs):
            e.submit_u8_device(fptr[i], fptr[i + 1], W1080, H1080, W1080, optr[0])
            e.wait()
        for r in e.kernel_profile():
            k = per_kernel.setdefault(r["name"], dict(ms=0.0, launches=0, alg_bytes=0.0))
            k["ms"] += r["ms"] / npairs
            k["launches"] += r["launches"] // npairs
            k["alg_bytes"] += r["alg_bytes"] / npairs
        e.enable_kernel_profile(False)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        top = max(per_kernel.items(), key=lambda kv: kv[1]["ms"])
        ach = top[1]["alg_bytes"] / (top[1]["ms"] * 1e-3) / 1e9
        pair_bytes = alg_bytes(W1080, H1080, 12, 0.75, 6, 2, True)
        tot_ms = sum(k["ms"] for k in per_kernel.values())
        # ncu --set full captures (profiles/): dram bytes read+write per launch of the dominant kernels
        traffic = {}
        try:  # dram bytes per launch from the committed ncu pass of the same pair (profiles/r01_traffic_c3.json)
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_c3.json")))["per_pair"]
            for kname, r in prof.items():
                traffic[kname.split("<")[0]] = (r["dram_read_MB"] + r["dram_write_MB"]) * 1e6 / r["launches"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": top[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(top[0]), "peak_source": peak_src,
                "note": "k_sor_wavefront keeps the reference's lexicographic Gauss-Seidel order, so it is bound by the "
                        "dependency chain (one warp per 32 rows, ~270 cycles per column step), not by HBM: its HBM "
                        "fraction is small by construction and throughput comes from concurrent pairs; the kernel that "
                        "fills the GPU is k_patch_search (issue-bound, ~76% issue-active in profiles/)",
                "kernel_ms_per_pair": top[1]["ms"], "kernel_launches_per_pair": top[1]["launches"],
                "kernel_share_of_pair": top[1]["ms"] / tot_ms,
                "how": "CUDA events around every launch of one un-graphed pass over %d pairs on the engine's stream" % npairs,
                "whole_pair": {"alg_bytes": pair_bytes, "achieved_gbs": pair_bytes * value / N / 1e9,
                               "frac": pair_bytes * value / N / 1e9 / peak},
                "per_kernel_ms": {k: round(v["ms"], 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1]["ms"])}}

    # ---- side measurement: C4a (4K, lv 7->0, 16 it, variational) latency and throughput on this GPU
    extra = None
    if rank == 0 and N == 1 and not args.no_extra:
        try:
            for e in engines:
                e.close()
            engines = []
            from tests.synth import synth_pair
            a4, b4, _ = synth_pair(3840, 2160, seed=2)
            p4 = F.Params.from_argv("7 0 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0".split())
            S4 = 16  # saturates at 4K (tools/streams_sweep_4k.py: 4 -> 6.7, 8 -> 6.1, 16 -> 5.9, 32 -> 5.9 ms/pair)
            eng4 = [F.Engine(p4, 3840, 2160, local_rank) for _ in range(S4)]
            da, db = torch.from_numpy(a4).to(dev), torch.from_numpy(b4).to(dev)
            o4 = torch.empty((S4, 2160, 3840, 2), dtype=torch.float32, device=dev)
            for e in eng4:
                e.submit_u8_device(da.data_ptr(), db.data_ptr(), 3840, 2160, 3840, o4[0].data_ptr())
                e.wait()
            from flowonthego_b200 import api as _api
            eng4[0].set_option(_api.OPT_SOR_GROUP, 16)  # lone pair: the low-latency SOR instantiation
            eng4[0].submit_u8_device(da.data_ptr(), db.data_ptr(), 3840, 2160, 3840, o4[0].data_ptr())
            eng4[0].wait()
            t0 = time.perf_counter()
            for _ in range(5):
                eng4[0].submit_u8_device(da.data_ptr(), db.data_ptr(), 3840, 2160, 3840, o4[0].data_ptr())
                eng4[0].wait()
            lat = (time.perf_counter() - t0) / 5 * 1e3
            eng4[0].set_option(_api.OPT_SOR_GROUP, 0)
            eng4[0].submit_u8_device(da.data_ptr(), db.data_ptr(), 3840, 2160, 3840, o4[0].data_ptr())
            eng4[0].wait()
            t0 = time.perf_counter()
            for i in range(4 * S4):
                eng4[i % S4].submit_u8_device(da.data_ptr(), db.data_ptr(), 3840, 2160, 3840, o4[i % S4].data_ptr())
            for e in eng4:
                e.wait()
            thr = (time.perf_counter() - t0) / (4 * S4) * 1e3
            b4k = alg_bytes(3840, 2160, 12, 0.75, 7, 0, True)
            extra = {"workload": "C4a: 3840x2160 synthetic pair, p12 ov0.75 lv7->0 16it + variational",
                     "latency_ms_per_pair_1stream": lat, "ms_per_pair_%dstreams" % S4: thr,
                     "alg_bytes": b4k, "hbm_frac_at_throughput": b4k / (thr * 1e-3) / 1e9 / (roof["peak"] if roof else 6650.0)}
            for e in eng4:
                e.close()
        except Exception as ex:  # the side measurement must never break the headline line
            extra = {"error": repr(ex)}

    if rank == 0:
        h2d = 2 * W1080 * H1080 * B
        d2h = 8 * W1080 * H1080 * B
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": K, "warmup": Wm,
                "ms_per_step": ms / K, "ms_per_pair": ms / K / B, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config, streams_per_gpu=S, batched_handles=(Sb if nb > 1 else 0), pairs_per_launch=nb,
                               l2="inputs larger than L2: %d MB of frames per step + %d MB of per-engine workspace"
                                  % ((B + 1) * W1080 * H1080 >> 20, 0)),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": Ke, "ms_per_step": ms_e / Ke},
                "gpu_launches": int(launches_per_call) * len(chunks) * K if nb > 1 else int(launches_per_call) * B * K,
                "launches_per_pair": launches_per_pair, "pairs_per_launch": nb,
                "clocks": clocks, "roofline": roof,
                "gathered_summaries": None if summaries is None else int(np.isfinite(summaries).all(axis=1).sum())}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if extra is not None:
            line["extra_c4a_4k"] = extra
        if stream_extra is not None:
            line["extra_video_stream_e2e"] = stream_extra
        if level_extra is not None:
            line["extra_e2e_engine_output"] = level_extra
        print(json.dumps(line))
    for e in engines + bengines:
        e.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
