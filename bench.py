#!/usr/bin/env python
"""bench.py -- headline benchmark of the DIS optical-flow hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c5|c4a|c4b] [--arith exact|fast]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads:
  c5  (default)  C5/C3: 1080p video stream -- triangle-wave affine trajectory of the grey road_HD frame
                 (tests/golden/road_HD_gray.png = cv2.imread(images/road_HD.jpg, GRAYSCALE)),
                 operating point 3 (p12 ov0.75 lv6->2 16it) + variational refinement
  c4a            C4a: 3840x2160 stream built the same way from yosemite_4k (one C3-sized affine step per frame),
                 p12 ov0.75 lv7->0 16it + variational refinement;  c4b: the same with 128 iterations
One step = one pass of the whole hot path (pyramid -> inverse search -> densify -> refine -> upsample) over
`batch` consecutive frame pairs per GPU.  Independent pairs shard across ranks with no data-path collective
(weak scaling: the per-GPU batch is fixed; `--total-pairs P` fixes the job instead -- P / N pairs per GPU and step,
"scaling": "strong"; P = 1024 is the C5 stream of BASELINE.md).

  value : pairs/s with the frames already resident in HBM; full-resolution flow left in HBM; with N > 1 the
          engine's level flows of every step are gathered on rank 0 over NCCL (overlapped with the next step)
  e2e   : pairs/s through the reference-facing C-ABI video front end (dis_video_push / dis_video_pop) with pinned
          HOST buffers: every u8 frame uploaded once, the engine's own output (OFC::OFClass outflow, level lv_l)
          copied back per pair -- all inside the timed region.  `extra_e2e_full_flow` is the same workload
          through dis_submit_u8 with the full-resolution flow (16.6 MB per 1080p pair) copied back.
  roofline / cpu_baseline : see DESIGN.md; the CPU leg is the reference's own engine (oracle/_ref, compiled
          verbatim) or, if that was not built, the C restatement (oracle/).

--impl reference times the reference's CPU implementation on the host cores for the same workload (this arm
never imports the product package).
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

METRIC, UNIT = "frame_pairs_per_sec", "pairs/s"

CONFIGS = {
    "c5": dict(workload="C5/C3: 1080p stream from road_HD (tests/golden/road_HD_gray.png, triangle-wave affine trajectory), "
                        "operating point 3 (p12 ov0.75 lv6->2 16it) + variational refinement",
               w=1920, h=1080, base="road_HD_gray.png", traj=(0.05, 1e-4, (0.9, -0.4), 64),
               argv="6 2 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0", batch=128, streams=128, nb=8, bh=32,
               cpu_pairs_per_core=4),
    "c4a": dict(workload="C4a: 3840x2160 stream from yosemite_4k (tests/golden/yosemite_4k_gray.png, one C3 affine step per "
                         "frame), p12 ov0.75 lv7->0 16it + variational refinement",
                w=3840, h=2160, base="yosemite_4k_gray.png", traj=(0.4, 0.004, (3.5, -2.25), 8),
                argv="7 0 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0", batch=16, streams=16, nb=1, bh=0,
                cpu_pairs_per_core=1),
}
CONFIGS["c4b"] = dict(CONFIGS["c4a"], workload=CONFIGS["c4a"]["workload"].replace("C4a", "C4b").replace("16it", "128it"),
                      argv="7 0 128 128 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0")
PARAM_NAMES = ("lv_f", "lv_l", "maxiter", "miniter", "mindprate", "mindrrate", "minimgerr", "patchsz", "poverl", "usefbcon",
               "patnorm", "costfct", "usetvref", "tv_alpha", "tv_gamma", "tv_delta", "tv_innerit", "tv_solverit", "tv_sor",
               "verbosity")
_INT = {"lv_f", "lv_l", "maxiter", "miniter", "patchsz", "usefbcon", "patnorm", "costfct", "usetvref", "tv_innerit",
        "tv_solverit", "verbosity"}


def params_dict(argv):
    return {n: (int(float(v)) if n in _INT else float(v)) for n, v in zip(PARAM_NAMES, argv.split())}


# ------------------------------------------------------------------------------------------ data
def make_frames(cfg, n_frames):
    """Frames 0..n_frames-1 of the workload's stream (u8, [n, h, w]) from the committed first frame."""
    from tests import synth
    base = synth.load_gray(cfg["base"])
    assert base.shape == (cfg["h"], cfg["w"]), base.shape
    rot, dsc, sh, period = cfg["traj"]
    return np.stack([synth.stream_frame(base, k, rot, dsc, sh, period) for k in range(n_frames)])


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


def cfg_alg_bytes(cfg):
    p = params_dict(cfg["argv"])
    return alg_bytes(cfg["w"], cfg["h"], p["patchsz"], p["poverl"], p["lv_f"], p["lv_l"], bool(p["usetvref"]),
                     p["tv_innerit"], p["tv_solverit"])


# ------------------------------------------------------------------------------- CPU reference
_W = {}


def _cpu_init(kind, frames, pd):
    """Worker start-up (outside every timed region): imports, library load, the frames of the sample."""
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        pass
    if kind == "reference":
        from oracle import ref_driver
        ref_driver.ref_lib()
        _W["run"] = lambda a, b: ref_driver.run_dense_ref(a, b, pd)
    else:
        from oracle import port
        port.build()
        _W["run"] = lambda a, b: port.run_u8(a, b, pd)
    _W["frames"] = frames


def _cpu_run(job):
    """`n` frame pairs starting at pair `first` through the reference CPU path (pyramids, engine, upsample), one thread."""
    first, n = job
    fr, run = _W["frames"], _W["run"]
    t0 = time.perf_counter()
    for i in range(n):
        k = (first + i) % (len(fr) - 1)
        run(fr[k], fr[k + 1])
    return time.perf_counter() - t0


class CpuReference:
    """The reference's CPU implementation on every host core: one single-thread worker process per core (the
    reference itself is single-threaded, kroeger/CMakeLists.txt:28-33), created and warmed BEFORE anything is
    timed; a step hands each worker `pairs_per_core` pairs and is timed from dispatch to the last result."""

    def __init__(self, frames, pd, max_procs=None):
        import multiprocessing as mp
        self.kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdis_ref.so")) else "port"
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        if max_procs:
            self.cores = min(self.cores, max_procs)
        ctx = mp.get_context("spawn")  # cv2 threads make fork unsafe
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init, initargs=(self.kind, frames, pd))
        self.pool.map(_cpu_run, [(i, 1) for i in range(self.cores)], chunksize=1)  # untimed: page in, first-call costs

    def step(self, pairs_per_core):
        """-> (wall seconds, pairs, CPU seconds summed over the workers)"""
        jobs = [(i * pairs_per_core, pairs_per_core) for i in range(self.cores)]
        t0 = time.perf_counter()
        per = self.pool.map(_cpu_run, jobs, chunksize=1)
        return time.perf_counter() - t0, self.cores * pairs_per_core, float(sum(per))

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self, walls, pairs, cpu_s, pairs_per_core):
        single_ms = 1e3 * cpu_s / pairs
        ideal = self.cores / (single_ms * 1e-3)
        value = pairs / sum(walls)
        return dict(value=value, unit=UNIT, cores=self.cores, kind=self.kind,
                    single_thread_ms_per_pair=single_ms, cores_over_single_thread=ideal, agreement=value / ideal,
                    sample="%d timed step(s) of %d pairs of the workload (%d per worker x %d single-thread workers, pool "
                           "created and warmed before the clock starts): %.2f s wall, %.1f CPU-s; %.0f ms/pair per thread "
                           "-> %d cores / that = %.1f pairs/s, measured %.1f" %
                           (len(walls), self.cores * pairs_per_core, pairs_per_core, self.cores, sum(walls), cpu_s,
                            single_ms, self.cores, ideal, value))


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


# ------------------------------------------------------------------------------------ product arm
class Bench:
    """Everything the product arm measures for one workload on this rank's GPU."""

    def __init__(self, cfg, args, rank, local_rank, world):
        import torch
        import flowonthego_b200 as F
        self.torch, self.F = torch, F
        self.cfg, self.rank, self.world = cfg, rank, world
        self.dev = torch.device("cuda", local_rank)
        self.local_rank = local_rank
        self.dist = None
        self.W, self.H = cfg["w"], cfg["h"]
        self.B = (args.total_pairs // self.world if args.total_pairs else 0) or args.batch or cfg["batch"]
        self.S = max(1, min(args.streams or cfg["streams"], self.B))
        self.nb = max(1, min(args.pairs_per_launch if args.pairs_per_launch is not None else cfg["nb"], 8, self.B))
        self.nb_e2e = max(1, min(args.e2e_pairs_per_launch if args.e2e_pairs_per_launch is not None else self.nb, 8))
        self.Sb = max(1, args.batch_handles or cfg["bh"] or 1)
        self.p = F.Params.from_argv(cfg["argv"].split())
        self.arith = args.arith
        self.no_gather = args.no_gather
        self.main_stream = torch.cuda.current_stream()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, streams, tail=None):
        """K calls of fn between two events on the main stream that every engine stream is fenced against;
        barrier + synchronize on both sides; max over ranks.  Returns milliseconds."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.main_stream)
        for s in streams:
            s.wait_event(e0)
        for _ in range(steps):
            fn()
        if tail is not None:
            tail()
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            self.main_stream.wait_event(ev)
        e1.record(self.main_stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def make_engine(self, batch=1):
        e = self.F.Engine(self.p, self.W, self.H, self.local_rank, batch=batch)
        if self.arith == "fast":
            from flowonthego_b200 import api
            e.set_option(api.OPT_ARITH, 1)
        return e

    # ---- value: device-resident frames, nb pairs per launch on batched handles
    def device_arm(self, frames, K, Wm, sample_clocks=True):
        torch, F = self.torch, self.F
        B, nb, W, H = self.B, self.nb, self.W, self.H
        if nb > 1:
            engines = [self.make_engine(nb) for _ in range(self.Sb)]
        else:
            engines = [self.make_engine() for _ in range(self.S)]
        nE = len(engines)
        streams = [torch.cuda.ExternalStream(e.stream, device=self.dev) for e in engines]
        d_frames = torch.from_numpy(frames).to(self.dev)
        d_out = torch.empty((nE, nb, H, W, 2), dtype=torch.float32, device=self.dev)
        fptr = [d_frames[i].data_ptr() for i in range(B + 1)]
        optr = [[d_out[k, j].data_ptr() for j in range(nb)] for k in range(nE)]
        chunks = [list(range(c, min(c + nb, B))) for c in range(0, B, nb)]
        lshape = None
        gather = self.dist is not None and not self.no_gather
        turn = [0]
        NBUF = 4
        if gather:  # level flows of a step -> staging (D2D on the engine's stream) -> NCCL gather on rank 0; NBUF staging
            # buffers, so a step only waits for the gather issued NBUF steps earlier (never for a recent one: a handle
            # that idles until every rank has delivered the previous step costs ~9 % of the pairs/s)
            lshape = engines[0].level_flow_shape()
            stage = [torch.zeros((B,) + lshape, dtype=torch.float32, device=self.dev) for _ in range(NBUF)]
            gl = [[torch.empty_like(stage[0]) for _ in range(self.world)] for _ in range(NBUF)] if self.rank == 0 else [None] * NBUF
            sptr = [t.data_ptr() for t in stage]
            lbytes = int(np.prod(lshape)) * 4
            gstream = torch.cuda.Stream(device=self.dev)
            g_done = [None] * NBUF
            step_no = [0]

        host_t = [0.0, 0]

        def step():
            t_h0 = time.perf_counter()
            step_inner()
            host_t[0] += time.perf_counter() - t_h0
            host_t[1] += 1

        def step_inner():
            buf = None
            if gather:
                buf = step_no[0] % NBUF
                step_no[0] += 1
            used = set()
            for idx in chunks:
                k = turn[0] % nE
                turn[0] += 1
                e = engines[k]
                if gather and g_done[buf] is not None and k not in used:
                    streams[k].wait_event(g_done[buf])  # staging buffer free again (gather of step-2 finished)
                used.add(k)
                if gather:  # the run's last kernel also writes the level flows into the staging buffer of this step
                    e.set_level_export([sptr[buf] + i * lbytes for i in idx])
                if nb > 1:
                    e.submit_u8_device_batch([fptr[i] for i in idx], [fptr[i + 1] for i in idx], W, H, W, optr[k][:len(idx)])
                else:
                    e.submit_u8_device(fptr[idx[0]], fptr[idx[0] + 1], W, H, W, optr[k][0])
            if gather:
                for k in used:
                    ev = torch.cuda.Event()
                    ev.record(streams[k])
                    gstream.wait_event(ev)
                with torch.cuda.stream(gstream):
                    self.dist.gather(stage[buf], gl[buf], dst=0)
                    g_done[buf] = torch.cuda.Event()
                    g_done[buf].record(gstream)

        def tail():
            if gather:
                self.main_stream.wait_stream(gstream)

        for _ in range(max(Wm, 1)):
            step()
        self.barrier()
        sampler = ClockSampler(self.local_rank)
        if self.rank == 0 and sample_clocks:
            sampler.start()
        ms = self.timed(step, K, streams, tail)
        clocks = sampler.stop() if (self.rank == 0 and sample_clocks) else None
        launches_per_call = engines[0].timings()["launches"]
        res = dict(ms=ms, host_submit_ms_per_step=1e3 * host_t[0] / max(host_t[1], 1), value=self.world * B * K / (ms / 1e3), clocks=clocks, launches_per_call=launches_per_call,
                   calls_per_step=len(chunks), n_engines=nE,
                   gather=None if not gather else dict(
                       what="level-%d flows (%dx%d) of every step gathered on rank 0 with torch.distributed.gather over NCCL, "
                            "overlapped with the following steps' compute" % (self.p.lv_l, lshape[1], lshape[0]),
                       bytes_per_step_into_rank0=int(np.prod(lshape)) * 4 * B * (self.world - 1)))
        summ = None
        if gather and self.rank == 0:  # proof that the gathered fields are the computed ones
            last = (step_no[0] - 1) % NBUF
            self.torch.cuda.synchronize()
            summ = [float(gl[last][r].abs().mean()) for r in range(self.world)]
        res["gathered_mean_abs_flow_per_rank"] = summ
        res["mean_abs_flow"] = float(d_out[0, 0].abs().mean())
        for e in engines:
            e.wait()
            e.close()
        del d_out, d_frames
        torch.cuda.empty_cache()
        return res

    # ---- e2e: the video front end with pinned host buffers
    def e2e_arm(self, frames, K):
        import ctypes
        torch, F = self.torch, self.F
        B, W, H, S = self.B, self.W, self.H, self.S
        L = F.lib()
        vid = ctypes.c_void_p()
        nbv = self.nb_e2e if S % max(self.nb_e2e, 1) == 0 else 1  # pairs per launch of the video front end
        rc = L.dis_video_create_batched(ctypes.byref(self.p), 1, W, H, self.local_rank, S, nbv, ctypes.byref(vid))
        if rc != 0:
            raise RuntimeError(L.dis_last_error(None).decode())
        nH = L.dis_video_handles(vid)
        if self.arith == "fast":
            from flowonthego_b200 import api
            for k in range(nH):
                L.dis_set_option(L.dis_video_handle(vid, k), api.OPT_ARITH, 1)
        fw, fh = ctypes.c_int(), ctypes.c_int()
        nfl = L.dis_video_flow_size(vid, ctypes.byref(fw), ctypes.byref(fh))
        h_frames = F.pinned_empty(frames.shape, np.uint8)
        h_frames[...] = frames
        h_lvl = [F.pinned_empty((fh.value, fw.value, 2), np.float32) for _ in range(S)]
        fp = ctypes.POINTER(ctypes.c_float)
        outp = [x.ctypes.data_as(fp) for x in h_lvl]
        inp = [h_frames[i].ctypes.data for i in range(B + 1)]
        streams = [torch.cuda.ExternalStream(L.dis_stream(L.dis_video_handle(vid, k)), device=self.dev) for k in range(nH)]
        push, pop, pending = L.dis_video_push, L.dis_video_pop, L.dis_video_pending
        npush = [0]

        def step():  # frames 0..B of the stream: B pairs; the first frame of a step is pushed only once per stream
            if npush[0] == 0:
                push(vid, inp[0], W, None)
            for i in range(B):
                if pending(vid) >= S:
                    pop(vid, None)
                if push(vid, inp[i + 1], W, outp[i % S]) != 0:
                    raise RuntimeError(L.dis_last_error(None).decode())
            npush[0] += 1

        def drain():
            while pending(vid) > 0:
                pop(vid, None)

        step()
        drain()
        ms = self.timed(step, K, streams, drain)
        L.dis_video_destroy(vid)
        return dict(value=self.world * B * K / (ms / 1e3), unit=UNIT, h2d_bytes_per_step=W * H * B,
                    d2h_bytes_per_step=int(nfl) * 4 * B, steps=K, ms_per_step=ms / K, pairs_in_flight=S,
                    pairs_per_launch=nbv,
                    api="dis_video_create_batched/dis_video_push/dis_video_pop (include/dis_c.h): pinned host u8 frames in, each uploaded once; "
                        "per pair the engine's own output -- OFC::OFClass outflow, level-%d flow %dx%d -- copied to pinned "
                        "host memory (DIS_VIDEO_OUT_LEVEL, the stream default)" % (self.p.lv_l, fw.value, fh.value),
                    mean_abs_level_flow=float(np.abs(h_lvl[0]).mean()))

    # ---- the round-1 e2e definition, kept as a side line: dis_submit_u8, full-resolution flow copied back
    def e2e_full_arm(self, frames, K):
        F = self.F
        B, W, H = self.B, self.W, self.H
        S = min(self.S, 32)  # 16.6 MB of pinned host memory per handle
        engines = [self.make_engine() for _ in range(S)]
        streams = [self.torch.cuda.ExternalStream(e.stream, device=self.dev) for e in engines]
        h_frames = F.pinned_empty(frames.shape, np.uint8)
        h_frames[...] = frames
        h_out = [F.pinned_empty((H, W, 2), np.float32) for _ in range(S)]

        def step():
            for i in range(B):
                e = engines[i % S]
                if i >= S:
                    e.wait()
                e.submit_u8(h_frames[i], h_frames[i + 1], h_out[i % S])
            for e in engines:
                e.wait()

        step()
        ms = self.timed(step, K, streams)
        for e in engines:
            e.close()
        return dict(value=self.world * B * K / (ms / 1e3), unit=UNIT, h2d_bytes_per_step=2 * W * H * B,
                    d2h_bytes_per_step=8 * W * H * B, steps=K,
                    api="dis_submit_u8/dis_wait, one pair per call, both frames uploaded, full-resolution fp32 flow copied back")

    # ---- host <-> device copy rates of this rank while every rank copies (names the e2e limiter)
    def pcie_probe(self):
        torch = self.torch
        n = 256 << 20
        h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d = torch.empty(n, dtype=torch.uint8, device=self.dev)
        out = {}
        for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
            dst.copy_(src, non_blocking=True)
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            e1.record()
            self.barrier()
            out[name + "_gbs_per_rank_all_ranks_active"] = 4 * n / (self.max_over_ranks(e0.elapsed_time(e1)) * 1e-3) / 1e9
        return out

    # ---- per-kernel CUDA-event times of un-graphed runs: roofline of the dominant kernel
    def roofline(self, frames, value_per_gpu, npairs=4):
        torch = self.torch
        W, H = self.W, self.H
        e = self.make_engine()
        d_frames = torch.from_numpy(frames[:npairs + 1]).to(self.dev)
        d_out = torch.empty((H, W, 2), dtype=torch.float32, device=self.dev)
        e.submit_u8_device(d_frames[0].data_ptr(), d_frames[1].data_ptr(), W, H, W, d_out.data_ptr())
        e.wait()
        e.enable_kernel_profile(True)
        for i in range(npairs):
            e.submit_u8_device(d_frames[i].data_ptr(), d_frames[i + 1].data_ptr(), W, H, W, d_out.data_ptr())
            e.wait()
        per_kernel = {}
        for r in e.kernel_profile():
            k = per_kernel.setdefault(r["name"], dict(ms=0.0, launches=0, alg_bytes=0.0))
            k["ms"] += r["ms"] / npairs
            k["launches"] += r["launches"] // npairs
            k["alg_bytes"] += r["alg_bytes"] / npairs
        e.enable_kernel_profile(False)
        e.close()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else \
            "fallback 6.65 TB/s (B200_PROFILING.md)"
        # the SOR stage (sor_coupled) runs as k_sor_wavefront on the large levels and k_sor_small on the others: one
        # logical kernel for the roofline line; per_kernel_ms below keeps them apart
        stage = dict(per_kernel)
        sor = [k for k in stage if k.startswith("k_sor_")]
        if len(sor) > 1:
            merged = dict(ms=sum(stage[k]["ms"] for k in sor), launches=sum(stage[k]["launches"] for k in sor),
                          alg_bytes=sum(stage[k]["alg_bytes"] for k in sor))
            for k in sor:
                del stage[k]
            stage["+".join(sorted(sor))] = merged
        top = max(stage.items(), key=lambda kv: kv[1]["ms"])
        ach = top[1]["alg_bytes"] / (top[1]["ms"] * 1e-3) / 1e9
        pair_bytes = cfg_alg_bytes(self.cfg)
        tot_ms = sum(k["ms"] for k in per_kernel.values())
        traffic = None
        try:  # dram bytes per launch of that kernel from the committed `ncu --set full` pass over the same pair
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_%s.json" % self.cfg["key"])))["per_pair"]
            hit = [r for kname, r in tj.items() if kname.split("<")[0] in top[0].split("+")]
            if hit:
                traffic = sum(r["dram_read_MB"] + r["dram_write_MB"] for r in hit) * 1e6 / sum(r["launches"] for r in hit)
        except Exception:
            pass
        return {"bound": "hbm", "kernel": top[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel_ms_per_pair": top[1]["ms"], "kernel_launches_per_pair": top[1]["launches"],
                "kernel_share_of_pair": top[1]["ms"] / tot_ms,
                "how": "CUDA events around every launch of an un-graphed pass over %d pairs on the engine's stream "
                       "(dis_enable_kernel_profile); achieved = SURVEY 8(d) algorithmic bytes of those launches / their time" % npairs,
                "whole_pair": {"alg_bytes": pair_bytes, "achieved_gbs": pair_bytes * value_per_gpu / 1e9,
                               "frac": pair_bytes * value_per_gpu / 1e9 / peak,
                               "what": "algorithmic bytes per pair x measured pairs/s per GPU (the whole path at throughput)"},
                "per_kernel_ms": {k: round(v["ms"], 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1]["ms"])},
                "per_kernel_hbm_frac": {k: round(v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 4)
                                        for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1]["ms"]) if v["ms"] > 0}}

    # ---- one pair alone on the GPU (graph replay), CUDA events
    def lone_latency(self, frames, reps=10):
        torch = self.torch
        W, H = self.W, self.H
        from flowonthego_b200 import api
        e = self.make_engine()
        e.set_option(api.OPT_SOR_GROUP, 16)
        st = torch.cuda.ExternalStream(e.stream, device=self.dev)
        da, db = torch.from_numpy(frames[0]).to(self.dev), torch.from_numpy(frames[1]).to(self.dev)
        out = torch.empty((H, W, 2), dtype=torch.float32, device=self.dev)
        for _ in range(3):
            e.submit_u8_device(da.data_ptr(), db.data_ptr(), W, H, W, out.data_ptr())
        e.wait()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            e.submit_u8_device(da.data_ptr(), db.data_ptr(), W, H, W, out.data_ptr())
        e1.record(st)
        e.wait()
        ms = e0.elapsed_time(e1) / reps
        e.close()
        return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--arith", default="exact", choices=["exact", "fast"],
                    help="fast = DIS_OPT_ARITH tolerance mode (FMA contraction; NOT the parity claim), see DESIGN.md")
    ap.add_argument("--batch", type=int, default=0, help="frame pairs per step per GPU (default: per config)")
    ap.add_argument("--total-pairs", type=int, default=0,
                    help="strong scaling: frame pairs per step over ALL GPUs (1024 = the C5 stream); default: weak scaling")
    ap.add_argument("--streams", type=int, default=0, help="engine instances (CUDA streams) per GPU for the host-buffer arms")
    ap.add_argument("--pairs-per-launch", type=int, default=None,
                    help="batched handles for the device-resident arm (dis_create_batch); 1 = one pair per launch")
    ap.add_argument("--batch-handles", type=int, default=0, help="number of batched handles per GPU")
    ap.add_argument("--e2e-pairs-per-launch", type=int, default=None,
                    help="pairs per launch of the video front end in the e2e arm (default: as --pairs-per-launch; 1 = dis_video_create)")
    ap.add_argument("--nccl-channels", type=int, default=0, help="NCCL_MAX_NCHANNELS for the result gather (default: by N)")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: leave the level flows on their GPUs (diagnostic)")
    ap.add_argument("--no-extra", action="store_true", help="skip the side measurements (4K, full-flow e2e, latency)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: keep a private copy of it and send everything else that writes to fd 1
    # (NCCL's version banner, library chatter) to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N, K, Wm = args.gpus, args.steps, max(args.warmup, 0)
    cfg = dict(CONFIGS[args.config], key=args.config)
    B = (args.total_pairs // world if args.total_pairs else 0) or args.batch or cfg["batch"]
    scaling = "strong" if args.total_pairs else "weak"
    pd = params_dict(cfg["argv"])
    S = max(1, min(args.streams or cfg["streams"], B))
    nb = max(1, min(args.pairs_per_launch if args.pairs_per_launch is not None else cfg["nb"], 8, B))
    Sb = max(1, args.batch_handles or cfg["bh"] or 1)
    frames_mb = (B + 1) * cfg["w"] * cfg["h"] >> 20
    # one dict for both arms (the reference arm prints it verbatim)
    config = {"workload": cfg["workload"], "config_id": args.config, "resolution": [cfg["w"], cfg["h"]],
              "pairs_per_step_per_gpu": B, "params": cfg["argv"], "arith": args.arith,
              "sharding": "independent pairs, dp%d" % N, "streams_per_gpu": S, "batched_handles": (Sb if nb > 1 else 0),
              "pairs_per_launch": nb,
              "l2": "inputs larger than L2: %d MB of distinct frames per step (+ per-engine workspaces)" % frames_mb}

    # ---------------------------------------------------------------- reference arm (CPU; never imports the product)
    if args.impl == "reference":
        if rank != 0:
            return
        frames = make_frames(cfg, 9 if args.config == "c5" else 3)
        ref = CpuReference(frames, pd)
        ppc = cfg["cpu_pairs_per_core"]
        Kr, Wr = K, Wm
        if args.config != "c5":  # a 4K pair costs ~16 CPU-seconds: bound the run to a few minutes and say so
            Kr, Wr = min(K, 3), min(Wm, 1)
        for _ in range(Wr):
            ref.step(ppc)
        walls, pairs, cpu_s = [], 0, 0.0
        for _ in range(Kr):
            wall, n, cs = ref.step(ppc)
            walls.append(wall)
            pairs += n
            cpu_s += cs
        ref.close()
        cb = ref.describe(walls, pairs, cpu_s, ppc)
        v = cb["value"]
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": N,
                          "steps": Kr, "warmup": Wr, "ms_per_step": 1e3 * sum(walls) / len(walls),
                          "ms_per_pair": 1e3 / v, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              file=json_out, flush=True)
        return

    # ---------------------------------------------------------------- product arm (B200)
    frames = make_frames(cfg, B + 1)
    cpu_base = None
    if N == 1 and rank == 0:  # before CUDA is initialised in this process
        ref = CpuReference(frames[:9], pd)
        ppc = cfg["cpu_pairs_per_core"] * (2 if args.config == "c5" else 1)
        wall, n, cs = ref.step(ppc)
        ref.close()
        cpu_base = ref.describe([wall], n, cs, ppc)

    import torch
    torch.cuda.set_device(local_rank)
    bench = Bench(cfg, args, rank, local_rank, world)
    if world > 1:
        import torch.distributed as dist
        # the only collective is the result gather (<= 1 GB per step into rank 0): a few channels move that easily
        # and leave the SMs to the flow kernels
        os.environ.setdefault("NCCL_MAX_NCHANNELS", str(args.nccl_channels or max(2, min(8, N))))
        dist.init_process_group("nccl", device_id=bench.dev)
        bench.dist = dist

    dev_res = bench.device_arm(frames, K, Wm)
    ms, value = dev_res["ms"], dev_res["value"]
    Ke = max(1, min(K, 6))
    e2e = bench.e2e_arm(frames, Ke)
    pcie = bench.pcie_probe()
    e2e["host_copy_rates"] = pcie
    per_gpu_e2e = e2e["value"] / N
    h2d_need = per_gpu_e2e * cfg["w"] * cfg["h"] / 1e9
    d2h_need = per_gpu_e2e * e2e["d2h_bytes_per_step"] / B / 1e9
    e2e["limiter"] = ("host link: needs %.1f GB/s h2d (%.0f%% of the probed rate) and %.1f GB/s d2h (%.0f%%) per rank; "
                      "device-resident rate is %.0f pairs/s per GPU" %
                      (h2d_need, 100 * h2d_need / pcie["h2d_gbs_per_rank_all_ranks_active"], d2h_need,
                       100 * d2h_need / pcie["d2h_gbs_per_rank_all_ranks_active"], value / N))

    extras = {}
    roof = None
    if rank == 0:
        roof = bench.roofline(frames, value / N)
    if not args.no_extra:
        try:
            extras["extra_e2e_full_flow"] = bench.e2e_full_arm(frames, max(1, min(K, 3)))
        except Exception as ex:
            extras["extra_e2e_full_flow"] = {"error": repr(ex)}
    if rank == 0 and N == 1 and not args.no_extra:
        try:
            extras["latency_ms_one_pair_alone"] = bench.lone_latency(frames)
        except Exception as ex:
            extras["latency_ms_one_pair_alone"] = {"error": repr(ex)}
        if args.config == "c5":  # the north-star 4K configuration as a side line (first-class: --config c4a)
            try:
                cfg4 = dict(CONFIGS["c4a"], key="c4a")
                a4 = argparse.Namespace(**vars(args))
                a4.batch, a4.streams, a4.pairs_per_launch, a4.batch_handles = 0, 0, None, 0
                b4 = Bench(cfg4, a4, rank, local_rank, world)
                fr4 = make_frames(cfg4, b4.B + 1)
                r4 = b4.device_arm(fr4, 3, 3, sample_clocks=False)
                ab4 = cfg_alg_bytes(cfg4)
                extras["extra_c4a_4k"] = {"workload": cfg4["workload"], "pairs_per_sec": r4["value"],
                                          "ms_per_pair": 1e3 / r4["value"], "steps": 3, "pairs_per_step": b4.B,
                                          "timing": "CUDA events, same code path as the headline (bench.py --config c4a)",
                                          "latency_ms_one_pair_alone": b4.lone_latency(fr4, reps=5),
                                          "alg_bytes": ab4, "hbm_frac_at_throughput": ab4 * r4["value"] / 1e9 / roof["peak"]}
            except Exception as ex:  # a side measurement must never break the headline line
                extras["extra_c4a_4k"] = {"error": repr(ex)}

    if rank == 0:
        calls = dev_res["calls_per_step"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": K, "warmup": Wm,
                "ms_per_step": ms / K, "ms_per_pair": ms / K / B, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "e2e": e2e, "gpu_launches": int(dev_res["launches_per_call"]) * calls * K,
                "launches_per_pair": dev_res["launches_per_call"] / nb, "pairs_per_launch": nb,
                "clocks": dev_res["clocks"], "roofline": roof, "gather": dev_res["gather"],
                "gathered_mean_abs_flow_per_rank": dev_res["gathered_mean_abs_flow_per_rank"],
                "mean_abs_flow": dev_res["mean_abs_flow"]}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        line.update(extras)
        print(json.dumps(line), file=json_out, flush=True)
    if bench.dist is not None:
        bench.dist.barrier()
        bench.dist.destroy_process_group()


if __name__ == "__main__":
    main()
