"""Python host-side mirror of the reference interface, on top of the C-ABI (include/dis_c.h).

* ``Params``        -- struct dis_params: the 20 CLI parameters of kroeger/run_dense.cpp:271-291
* ``Engine``        -- one dis_handle (one CUDA stream + workspace + CUDA graph)
* ``OFClass(...)``  -- same positional arguments as ``OFC::OFClass::OFClass`` (kroeger/oflow.h:84-111)
* ``run_dense(...)``-- the three call variants of the reference CLI (kroeger/README.md:48-88)

Everything computes in libdis_b200.so (hand-written sm_100a kernels).  There is no CPU path:
if the library is missing or no B200 is visible, calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("DIS_B200_LIB") or os.path.join(_HERE, "libdis_b200.so")  # override: experiments only
_LIB = None

PARAM_NAMES = ("lv_f", "lv_l", "maxiter", "miniter", "mindprate", "mindrrate", "minimgerr", "patchsz",
               "poverl", "usefbcon", "patnorm", "costfct", "usetvref", "tv_alpha", "tv_gamma", "tv_delta",
               "tv_innerit", "tv_solverit", "tv_sor", "verbosity")
_FLOAT_PARAMS = {"mindprate", "mindrrate", "minimgerr", "poverl", "tv_alpha", "tv_gamma", "tv_delta", "tv_sor"}

# enum dis_tap
TAP_IMG_A, TAP_IMG_A_DX, TAP_IMG_A_DY, TAP_IMG_B, TAP_PATCH_FLOW, TAP_FLOW_DENSE, TAP_FLOW_REFINED, \
    TAP_IMG_B_DX, TAP_IMG_B_DY = range(9)


class DisError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dis error %d: %s" % (code, msg))
        self.code = code


class Params(ctypes.Structure):
    """struct dis_params (include/dis_c.h)."""
    _fields_ = [(n, ctypes.c_float if n in _FLOAT_PARAMS else ctypes.c_int32) for n in PARAM_NAMES]

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for n in PARAM_NAMES:
            setattr(p, n, float(d[n]) if n in _FLOAT_PARAMS else int(d[n]))
        return p

    @classmethod
    def preset(cls, preset, width_org, verbosity=0):
        """Operating points 1..4, kroeger/run_dense.cpp:225-267."""
        p = cls()
        _check(lib().dis_params_preset(ctypes.byref(p), int(preset), int(width_org)), None)
        p.verbosity = verbosity
        return p

    @classmethod
    def from_argv(cls, args):
        """20 explicit parameters in CLI order (strings), kroeger/run_dense.cpp:271-291."""
        args = [str(a).encode() for a in args]
        arr = (ctypes.c_char_p * len(args))(*args)
        p = cls()
        _check(lib().dis_params_from_argv(ctypes.byref(p), len(args), arr), None)
        return p

    def to_dict(self):
        return {n: getattr(self, n) for n in PARAM_NAMES}

    def copy(self, **kw):
        d = self.to_dict()
        d.update(kw)
        return Params.from_dict(d)


class Timings(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("total_ms", "h2d_ms", "pyramid_ms", "search_ms", "densify_ms",
                                               "varref_ms", "finish_ms", "d2h_ms")] + [("launches", ctypes.c_int32)]


class KernelTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 32), ("level", ctypes.c_int32), ("launches", ctypes.c_int32),
                ("ms", ctypes.c_float), ("alg_bytes", ctypes.c_double)]


def lib():
    """Loads libdis_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(_LIB_PATH + " not built: run `python -c 'import __graft_entry__ as g; g.build()'`"
                              " or `make -C flowonthego_b200/csrc`")
        L = ctypes.CDLL(_LIB_PATH)
        vp, ip, fp = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float)
        fpp = ctypes.POINTER(fp)
        pp = ctypes.POINTER(Params)
        L.dis_version.restype = ctypes.c_char_p
        L.dis_last_error.restype = ctypes.c_char_p
        L.dis_last_error.argtypes = [vp]
        L.dis_params_preset.argtypes = [pp, ip, ip]
        L.dis_params_from_argv.argtypes = [pp, ip, ctypes.POINTER(ctypes.c_char_p)]
        L.dis_params_validate.argtypes = [pp, ctypes.c_char_p, ctypes.c_size_t]
        L.dis_padded_size.argtypes = [ip, ip, ip] + [ctypes.POINTER(ip)] * 4
        L.dis_auto_first_scale.argtypes = [ip, ip, ip]
        L.dis_create.argtypes = [pp, ip, ip, ip, ctypes.POINTER(vp)]
        L.dis_create_c.argtypes = [pp, ip, ip, ip, ip, ctypes.POINTER(vp)]
        L.dis_create_batch.argtypes = [pp, ip, ip, ip, ip, ip, ctypes.POINTER(vp)]
        L.dis_batch_size.argtypes = [vp]
        L.dis_submit_u8_device_batch.argtypes = [vp, ip, ctypes.POINTER(vp), ctypes.POINTER(vp), ip, ip, ip, ctypes.POINTER(vp)]
        L.dis_destroy.argtypes = [vp]
        L.dis_set_params.argtypes = [vp, pp]
        L.dis_set_option.argtypes = [vp, ip, ip]
        L.dis_run_pyramids.argtypes = [vp] + [fpp] * 6 + [ip, ip, ip, fp, fp]
        L.dis_run_u8.argtypes = [vp, vp, vp, ip, ip, ip, fp]
        L.dis_submit_u8.argtypes = [vp, vp, vp, ip, ip, ip, fp]
        L.dis_submit_u8_device.argtypes = [vp, vp, vp, ip, ip, ip, vp]
        L.dis_wait.argtypes = [vp]
        L.dis_fetch_level_flow.argtypes = [vp, fp, ctypes.c_size_t]
        L.dis_stream.restype = vp
        L.dis_stream.argtypes = [vp]
        L.dis_get_timings.argtypes = [vp, ctypes.POINTER(Timings)]
        L.dis_enable_stage_timing.argtypes = [vp, ip]
        L.dis_enable_taps.argtypes = [vp, ip]
        L.dis_fetch_tap.argtypes = [vp, ip, ip, fp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.dis_enable_kernel_profile.argtypes = [vp, ip]
        L.dis_get_kernel_profile.argtypes = [vp, ctypes.POINTER(KernelTime), ip, ctypes.POINTER(ip)]
        L.dis_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
        L.dis_host_free.argtypes = [vp]
        L.dis_read_image_gray.argtypes = [ctypes.c_char_p, vp, ctypes.c_size_t, ctypes.POINTER(ip), ctypes.POINTER(ip)]
        L.dis_read_image_bgr.argtypes = L.dis_read_image_gray.argtypes
        L.dis_video_create.argtypes = [pp, ip, ip, ip, ip, ip, ctypes.POINTER(vp)]
        L.dis_video_create_batched.argtypes = [pp, ip, ip, ip, ip, ip, ip, ctypes.POINTER(vp)]
        L.dis_video_handles.argtypes = [vp]
        L.dis_video_destroy.argtypes = [vp]
        L.dis_video_destroy.restype = None
        L.dis_video_push.argtypes = [vp, vp, ip, fp]
        L.dis_video_pop.argtypes = [vp, ctypes.POINTER(fp)]
        L.dis_video_pending.argtypes = [vp]
        L.dis_video_set_output.argtypes = [vp, ip]
        L.dis_video_set_reuse.argtypes = [vp, ip]
        L.dis_video_reuse.argtypes = [vp]
        L.dis_video_flow_size.argtypes = [vp, ctypes.POINTER(ip), ctypes.POINTER(ip)]
        L.dis_video_flow_size.restype = ctypes.c_size_t
        L.dis_video_handle.argtypes = [vp, ip]
        L.dis_video_handle.restype = vp
        L.dis_level_flow_size.argtypes = [vp, ctypes.POINTER(ip), ctypes.POINTER(ip)]
        L.dis_copy_level_flow_device.argtypes = [vp, ip, vp]
        L.dis_set_level_export.argtypes = [vp, ip, ctypes.POINTER(vp)]
        L.dis_level_flow_ptr.argtypes = [vp, ip]
        L.dis_level_flow_ptr.restype = vp
        L.dis_flow_to_color.argtypes = [fp, ip, ip, ctypes.c_float, ip, vp, fp]
        L.dis_flow_epe.argtypes = [fp, fp, ip, ip, ip, ip, ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]
        L.dis_write_png_bgr.argtypes = [ctypes.c_char_p, vp, ip, ip]
        L.dis_write_flo.argtypes = [ctypes.c_char_p, fp, ip, ip]
        L.dis_read_flo.argtypes = [ctypes.c_char_p, fp, ctypes.c_size_t, ctypes.POINTER(ip), ctypes.POINTER(ip)]
        _LIB = L
    return _LIB


def _check(rc, handle):
    if rc != 0:
        msg = lib().dis_last_error(handle)
        raise DisError(rc, msg.decode() if msg else "")


_fp = ctypes.POINTER(ctypes.c_float)


def _as_fp(a):
    return a.ctypes.data_as(_fp)


def padded_size(w, h, lv_f):
    """(w_pad, h_pad, left, top) -- kroeger/run_dense.cpp:298-311."""
    v = [ctypes.c_int() for _ in range(4)]
    _check(lib().dis_padded_size(w, h, lv_f, *[ctypes.byref(x) for x in v]), None)
    return tuple(x.value for x in v)


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory (dis_host_alloc)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = ctypes.c_void_p()
    if lib().dis_host_alloc(ctypes.byref(p), max(n, 1)) != 0:
        raise MemoryError("dis_host_alloc(%d)" % n)
    buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


OPT_SOR_GROUP, OPT_USE_GRAPH, OPT_LEVEL_OUTPUT, OPT_ARITH, OPT_SOR_SMALL = 1, 2, 3, 4, 5


class Engine:
    """One engine instance = one dis_handle (stream, workspace, CUDA graph)."""

    def __init__(self, params, max_w, max_h, device=0, channels=1, batch=1):
        """channels=1: grey (the reference's run_OF_INT build); channels=3: interleaved BGR
        (run_OF_RGB, SELECTCHANNEL=3).  batch > 1: every kernel launch serves that many pairs
        (dis_create_batch; use submit_u8_device_batch)."""
        self._h = ctypes.c_void_p()
        self.params = params if isinstance(params, Params) else Params.from_dict(params)
        self.channels = int(channels)
        self.batch = int(batch)
        _check(lib().dis_create_batch(ctypes.byref(self.params), self.channels, int(max_w), int(max_h), int(device),
                                      self.batch, ctypes.byref(self._h)), None)
        self._keep = None

    def close(self):
        if self._h:
            lib().dis_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_params(self, params):
        self.params = params if isinstance(params, Params) else Params.from_dict(params)
        _check(lib().dis_set_params(self._h, ctypes.byref(self.params)), self._h)

    def set_option(self, option, value):
        """OPT_SOR_GROUP (8 | 16), OPT_SOR_SMALL (-1 auto | 0 ... 9 row blocks), OPT_USE_GRAPH (0 | 1), OPT_LEVEL_OUTPUT: results never change.
        OPT_ARITH (0 exact | 1 tolerance mode, FMA contraction; maxiter <= 32 only) does change them."""
        _check(lib().dis_set_option(self._h, int(option), int(value)), self._h)
        if int(option) == OPT_LEVEL_OUTPUT:
            self._level_output = bool(value)

    # ---- whole run_dense data path ----------------------------------------------------------
    def run_u8(self, a, b, out=None):
        """Grey u8 frames (h, w) [channels=3: BGR (h, w, 3)] -> full-resolution flow (h, w, 2) float32.
        Synchronous."""
        self.submit_u8(a, b, out)
        return self.wait()

    def submit_u8(self, a, b, out=None):
        def rows_ok(x):  # u8, pixels (and channels) contiguous within a row
            if x.dtype != np.uint8 or x.strides[-1] != 1:
                return False
            return x.ndim == 2 or x.strides[1] == x.shape[2]
        a = a if rows_ok(a) else np.ascontiguousarray(a, np.uint8)
        b = b if rows_ok(b) else np.ascontiguousarray(b, np.uint8)
        want = (3, 3) if self.channels == 3 else (2, None)
        if a.ndim != want[0] or (a.ndim == 3 and a.shape[2] != want[1]) or a.shape != b.shape or \
                a.strides[0] != b.strides[0]:
            raise ValueError("need two %s u8 images of identical shape and pitch" %
                             ("BGR (h, w, 3)" if self.channels == 3 else "grey (h, w)"))
        h, w = a.shape[:2]
        if getattr(self, "_level_output", False):  # OPT_LEVEL_OUTPUT: raw engine output at level lv_l
            wp, hp, _, _ = padded_size(w, h, self.params.lv_f)
            shape = (hp >> self.params.lv_l, wp >> self.params.lv_l, 2)
        else:
            shape = (h, w, 2)
        if out is None:
            out = np.empty(shape, np.float32)
        elif out.shape != shape:
            raise ValueError("flow buffer %s, expected %s" % (out.shape, shape))
        self._keep = (a, b, out)
        _check(lib().dis_submit_u8(self._h, a.ctypes.data, b.ctypes.data, w, h, a.strides[0], _as_fp(out)), self._h)

    def submit_u8_device(self, d_a, d_b, w, h, pitch, d_flow):
        """Device pointers (ints) in, device pointer out; asynchronous on the engine's stream."""
        _check(lib().dis_submit_u8_device(self._h, d_a, d_b, w, h, pitch, d_flow), self._h)

    def submit_u8_device_batch(self, d_a, d_b, w, h, pitch, d_flow):
        """Up to `batch` pairs per call: sequences of device pointers (ints); asynchronous."""
        k = len(d_a)
        arr = ctypes.c_void_p * k
        _check(lib().dis_submit_u8_device_batch(self._h, k, arr(*d_a), arr(*d_b), w, h, pitch, arr(*d_flow)), self._h)

    def wait(self):
        _check(lib().dis_wait(self._h), self._h)
        keep, self._keep = self._keep, None
        return keep[2] if keep else None

    def level_flow_ptr(self, pair=0):
        """Device address of the level-lv_l flow of `pair` inside the workspace (dis_level_flow_ptr)."""
        return lib().dis_level_flow_ptr(self._h, int(pair))

    def level_flow_shape(self):
        w, h = ctypes.c_int(), ctypes.c_int()
        _check(lib().dis_level_flow_size(self._h, ctypes.byref(w), ctypes.byref(h)), self._h)
        return (h.value, w.value, 2)

    def copy_level_flow_device(self, pair, d_dst):
        """Enqueue a device-to-device copy of the level flow of `pair` to device pointer d_dst on the engine's stream."""
        _check(lib().dis_copy_level_flow_device(self._h, int(pair), d_dst), self._h)

    def set_level_export(self, d_level):
        """The following device submits also write pair b's level-lv_l flow to device pointer d_level[b]
        (dis_set_level_export; [] turns it off)."""
        arr = (ctypes.c_void_p * max(len(d_level), 1))(*d_level)
        _check(lib().dis_set_level_export(self._h, len(d_level), arr), self._h)

    def level_flow(self, w, h):
        """Raw engine output (level lv_l, padded size) of the last run_u8."""
        wp, hp, _, _ = padded_size(w, h, self.params.lv_f)
        sc = 1 << self.params.lv_l
        out = np.empty((hp // sc, wp // sc, 2), np.float32)
        _check(lib().dis_fetch_level_flow(self._h, _as_fp(out), out.size), self._h)
        return out

    # ---- the reference engine boundary ------------------------------------------------------
    def run_pyramids(self, pyr_a, pyr_b, width, height, initflow=None):
        """pyr_a/pyr_b = (I, Ix, Iy): lists over levels 0..lv_f of padded float32 arrays (entries below
        lv_l may be None).  Returns flow (height/2^lv_l, width/2^lv_l, 2)."""
        p = self.params
        ptrs, keep = [], []
        for lst in (*pyr_a, *pyr_b):
            arr = (_fp * (p.lv_f + 1))()
            for l in range(p.lv_f + 1):
                x = lst[l] if lst is not None and l < len(lst) else None
                if x is not None:
                    x = np.ascontiguousarray(x, np.float32)
                    keep.append(x)
                    arr[l] = _as_fp(x)
            ptrs.append(arr)
        sc = 1 << p.lv_l
        out = np.empty((height // sc, width // sc, 2), np.float32)
        init = None
        if initflow is not None:
            initflow = np.ascontiguousarray(initflow, np.float32)
            init = _as_fp(initflow)
        _check(lib().dis_run_pyramids(self._h, *ptrs, p.patchsz, width, height, init, _as_fp(out)), self._h)
        return out

    # ---- introspection ------------------------------------------------------------------------
    def enable_taps(self, on=True):
        """True/1: all pyramid levels built level by level; 2: the product's pyramid path (levels >= lv_l only)."""
        _check(lib().dis_enable_taps(self._h, int(on)), self._h)

    def enable_stage_timing(self, on=True):
        _check(lib().dis_enable_stage_timing(self._h, int(on)), self._h)

    def tap(self, which, level):
        n = ctypes.c_size_t()
        _check(lib().dis_fetch_tap(self._h, which, level, None, 0, ctypes.byref(n)), self._h)
        out = np.empty(n.value, np.float32)
        _check(lib().dis_fetch_tap(self._h, which, level, _as_fp(out), out.size, ctypes.byref(n)), self._h)
        return out

    def enable_kernel_profile(self, on=True):
        _check(lib().dis_enable_kernel_profile(self._h, int(on)), self._h)

    def kernel_profile(self):
        """[{name, level, launches, ms, alg_bytes}] accumulated since enable_kernel_profile(True)."""
        n = ctypes.c_int()
        _check(lib().dis_get_kernel_profile(self._h, None, 0, ctypes.byref(n)), self._h)
        arr = (KernelTime * max(n.value, 1))()
        _check(lib().dis_get_kernel_profile(self._h, arr, n.value, ctypes.byref(n)), self._h)
        return [dict(name=arr[i].name.decode(), level=arr[i].level, launches=arr[i].launches, ms=arr[i].ms,
                     alg_bytes=arr[i].alg_bytes) for i in range(n.value)]

    def timings(self):
        t = Timings()
        _check(lib().dis_get_timings(self._h, ctypes.byref(t)), self._h)
        return {n: getattr(t, n) for n, _ in Timings._fields_}

    @property
    def stream(self):
        return lib().dis_stream(self._h)


class FlowStream:
    """Video-stream front end (dis_video_*): push consecutive frames, get one flow per consecutive pair,
    `depth` pairs in flight on the GPU, every frame uploaded once.  output="level" (the default for streams):
    the engine's own output as OFC::OFClass delivers it, (h_pad/2^lv_l, w_pad/2^lv_l, 2) -- equal to
    Engine.level_flow() of that pair bit for bit; output="full": the full-resolution (h, w, 2) field, equal to
    Engine.run_u8 on that pair bit for bit."""

    def __init__(self, params, w, h, depth=8, device=0, channels=1, output="level", reuse=None, pairs_per_launch=1):
        self._v = ctypes.c_void_p()
        self.params = params if isinstance(params, Params) else Params.from_dict(params)
        self.w, self.h, self.depth, self.channels = int(w), int(h), int(depth), int(channels)
        # pairs_per_launch > 1 (dividing depth): batched handles, one launch chain per that many pushes
        _check(lib().dis_video_create_batched(ctypes.byref(self.params), self.channels, self.w, self.h, int(device),
                                              self.depth, int(pairs_per_launch), ctypes.byref(self._v)), None)
        shape = (self.h, self.w) if self.channels == 1 else (self.h, self.w, self.channels)
        self._frames = [pinned_empty(shape, np.uint8) for _ in range(self.depth + 1)]
        if output not in ("level", "full"):
            raise ValueError("output must be 'level' or 'full'")
        if reuse is not None:  # pyramid reuse between consecutive pairs (default: on for 2 <= depth <= 8)
            _check(lib().dis_video_set_reuse(self._v, int(bool(reuse))), None)
        self.reuse = bool(lib().dis_video_reuse(self._v))
        _check(lib().dis_video_set_output(self._v, 0 if output == "level" else 1), None)
        fw, fh = ctypes.c_int(), ctypes.c_int()
        lib().dis_video_flow_size(self._v, ctypes.byref(fw), ctypes.byref(fh))
        self.flow_shape = (fh.value, fw.value, 2)
        self._flows = [pinned_empty(self.flow_shape, np.float32) for _ in range(self.depth)]
        self._n = 0

    def handle(self, k):
        """k-th engine handle (a dis_handle*, e.g. for lib().dis_stream)."""
        return lib().dis_video_handle(self._v, int(k))

    def close(self):
        if self._v:
            lib().dis_video_destroy(self._v)
            self._v = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def pending(self):
        return lib().dis_video_pending(self._v)

    def push(self, frame):
        """Enqueue the next frame.  Requires pending < depth (pop first otherwise)."""
        stage = self._frames[self._n % (self.depth + 1)]
        if frame.shape != stage.shape:
            raise ValueError("frame shape %s, expected %s" % (frame.shape, stage.shape))
        if self._n and self.pending >= self.depth:
            raise DisError(2, "%d pairs in flight, pop one first" % self.depth)
        stage[...] = frame
        out = _as_fp(self._flows[(self._n - 1) % self.depth]) if self._n else None
        _check(lib().dis_video_push(self._v, stage.ctypes.data, stage.strides[0], out), None)
        self._n += 1

    def pop(self):
        """Wait for the oldest pair in flight; returns its flow (a view of a pinned buffer that is reused
        `depth` pairs later -- copy it to keep it)."""
        p = _fp()
        _check(lib().dis_video_pop(self._v, ctypes.byref(p)), None)
        addr = ctypes.cast(p, ctypes.c_void_p).value
        for f in self._flows:
            if f.ctypes.data == addr:
                return f
        raise DisError(1, "dis_video_pop returned an unknown buffer")

    def flows(self, frames):
        """Generator: flow of every consecutive pair of the iterable `frames` (copies)."""
        for fr in frames:
            if self.pending >= self.depth:
                yield self.pop().copy()
            self.push(fr)
        while self.pending > 0:
            yield self.pop().copy()


def OFClass(im_ao, im_ao_dx, im_ao_dy, im_bo, im_bo_dx, im_bo_dy, imgpadding, outflow, initflow, width, height,
            sc_f, sc_l, max_iter, min_iter, dp_thresh, dr_thresh, res_thresh, p_samp_s, patove, usefbcon, costfct,
            noc, patnorm, usetvref, tv_alpha, tv_gamma, tv_delta, tv_innerit, tv_solverit, tv_sor, verbosity,
            device=0):
    """Same positional arguments as ``OFC::OFClass::OFClass`` (kroeger/oflow.h:84-111); like the
    reference, all work happens in this call and the result is written into ``outflow``."""
    if noc not in (1, 3):
        raise DisError(2, "noc must be 1 (grey, SELECTCHANNEL=1) or 3 (RGB, SELECTCHANNEL=3)")
    if imgpadding != p_samp_s:
        raise DisError(2, "imgpadding must equal the patch size (kroeger/run_dense.cpp:393)")
    p = Params.from_dict(dict(lv_f=sc_f, lv_l=sc_l, maxiter=max_iter, miniter=min_iter, mindprate=dp_thresh,
                              mindrrate=dr_thresh, minimgerr=res_thresh, patchsz=p_samp_s, poverl=patove,
                              usefbcon=int(usefbcon), patnorm=patnorm, costfct=costfct, usetvref=int(usetvref),
                              tv_alpha=tv_alpha, tv_gamma=tv_gamma, tv_delta=tv_delta, tv_innerit=tv_innerit,
                              tv_solverit=tv_solverit, tv_sor=tv_sor, verbosity=verbosity))
    with Engine(p, width, height, device, channels=noc) as e:
        fl = e.run_pyramids((im_ao, im_ao_dx, im_ao_dy), (im_bo, im_bo_dx, im_bo_dy), width, height, initflow)
    np.asarray(outflow).reshape(fl.shape)[...] = fl
    return outflow


def write_flo(path, flow):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w = flow.shape[:2]
    _check(lib().dis_write_flo(os.fsencode(path), _as_fp(flow), w, h), None)


def read_flo(path):
    w, h = ctypes.c_int(), ctypes.c_int()
    _check(lib().dis_read_flo(os.fsencode(path), None, 0, ctypes.byref(w), ctypes.byref(h)), None)
    out = np.empty((h.value, w.value, 2), np.float32)
    _check(lib().dis_read_flo(os.fsencode(path), _as_fp(out), out.size, ctypes.byref(w), ctypes.byref(h)), None)
    return out


def read_image_gray(path):
    """PNG / PGM / PPM -> grey u8 (h, w), decoded natively with OpenCV's grey conversion."""
    w, h = ctypes.c_int(), ctypes.c_int()
    _check(lib().dis_read_image_gray(os.fsencode(path), None, 0, ctypes.byref(w), ctypes.byref(h)), None)
    out = np.empty((h.value, w.value), np.uint8)
    _check(lib().dis_read_image_gray(os.fsencode(path), out.ctypes.data, out.size, ctypes.byref(w), ctypes.byref(h)), None)
    return out


def read_image_bgr(path):
    """PNG / PGM / PPM -> BGR u8 (h, w, 3) like cv2.imread(.., IMREAD_COLOR) (kroeger/run_dense.cpp:203-206)."""
    w, h = ctypes.c_int(), ctypes.c_int()
    _check(lib().dis_read_image_bgr(os.fsencode(path), None, 0, ctypes.byref(w), ctypes.byref(h)), None)
    out = np.empty((h.value, w.value, 3), np.uint8)
    _check(lib().dis_read_image_bgr(os.fsencode(path), out.ctypes.data, out.size, ctypes.byref(w), ctypes.byref(h)), None)
    return out


def flow_to_color(flow, maxmotion=-1.0, device=0, want_stats=False):
    """Middlebury colour coding on the GPU (flow_code/C/color_flow.cpp MotionToColor): (h, w, 2) float32 ->
    (h, w, 3) u8 in B,G,R order [and dict(maxrad, minu, maxu, minv, maxv)]."""
    flow = np.ascontiguousarray(flow, np.float32)
    h, w = flow.shape[:2]
    out = np.empty((h, w, 3), np.uint8)
    st = np.zeros(5, np.float32)
    _check(lib().dis_flow_to_color(_as_fp(flow), w, h, float(maxmotion), int(device), out.ctypes.data, _as_fp(st)), None)
    if want_stats:
        return out, dict(zip(("maxrad", "minu", "maxu", "minv", "maxv"), map(float, st)))
    return out


def flow_epe(flow_a, flow_b, margin=0, device=0):
    """(mean, max, count) of the endpoint error |a - b|_2 on the GPU, border margin excluded."""
    a = np.ascontiguousarray(flow_a, np.float32)
    b = np.ascontiguousarray(flow_b, np.float32)
    if a.shape != b.shape or a.ndim != 3 or a.shape[2] != 2:
        raise ValueError("need two (h, w, 2) flow fields of identical shape")
    mean, mx, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
    _check(lib().dis_flow_epe(_as_fp(a), _as_fp(b), a.shape[1], a.shape[0], int(margin), int(device),
                              ctypes.byref(mean), ctypes.byref(mx), ctypes.byref(cnt)), None)
    return mean.value, mx.value, cnt.value


def write_png_bgr(path, bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    _check(lib().dis_write_png_bgr(os.fsencode(path), bgr.ctypes.data, bgr.shape[1], bgr.shape[0]), None)


def run_dense(img1, img2, outfile, *args, device=0, channels=1):
    """``run_dense img1 img2 out [X | 20 params]`` -- the reference CLI variants (kroeger/README.md:48-88).
    img1/img2 are file names (PNG/PGM/PPM, decoded natively to the same grey values as the
    reference's cv::imread(.., GRAYSCALE), kroeger/run_dense.cpp:208-209) or decoded grey u8 arrays;
    channels=3 is the reference's colour binary run_OF_RGB (BGR decode, kroeger/run_dense.cpp:203-206)."""
    def load(x):
        if isinstance(x, np.ndarray):
            return x
        return read_image_bgr(x) if channels == 3 else read_image_gray(x)
    a, b = load(img1), load(img2)
    if len(args) <= 1:
        p = Params.preset(int(args[0]) if args else 2, a.shape[1], verbosity=2)
    else:
        p = Params.from_argv(args)
    with Engine(p, a.shape[1], a.shape[0], device, channels=channels) as e:
        flow = e.run_u8(a, b)
    if outfile is not None:
        write_flo(outfile, flow)
    return flow
