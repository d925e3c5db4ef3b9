"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libdis_oracle.so (dis_oracle.c), the
plain-C restatement of the reference hot path.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs may import this module; the product never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PARAM_FIELDS = [("lv_f", ctypes.c_int32), ("lv_l", ctypes.c_int32), ("maxiter", ctypes.c_int32),
                ("miniter", ctypes.c_int32), ("mindprate", ctypes.c_float), ("mindrrate", ctypes.c_float),
                ("minimgerr", ctypes.c_float), ("patchsz", ctypes.c_int32), ("poverl", ctypes.c_float),
                ("usefbcon", ctypes.c_int32), ("patnorm", ctypes.c_int32), ("costfct", ctypes.c_int32),
                ("usetvref", ctypes.c_int32), ("tv_alpha", ctypes.c_float), ("tv_gamma", ctypes.c_float),
                ("tv_delta", ctypes.c_float), ("tv_innerit", ctypes.c_int32), ("tv_solverit", ctypes.c_int32),
                ("tv_sor", ctypes.c_float), ("verbosity", ctypes.c_int32)]


class DisParams(ctypes.Structure):
    """struct dis_params of include/dis_c.h."""
    _fields_ = PARAM_FIELDS

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for name, ct in PARAM_FIELDS:
            v = d[name]
            setattr(p, name, int(v) if ct is ctypes.c_int32 else float(v))
        return p

    def to_dict(self):
        return {name: getattr(self, name) for name, _ in PARAM_FIELDS}


def build(force=False):
    so = os.path.join(_HERE, "libdis_oracle.so")
    src = os.path.join(_HERE, "dis_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_free.argtypes = [ctypes.c_void_p]
        _LIB.oracle_free.restype = None
    return _LIB


_fp = ctypes.POINTER(ctypes.c_float)
_fpp = ctypes.POINTER(_fp)


def build_pyramid(img_u8, lv_f, pad, grads=True):
    """P1 restatement. Returns (I, Ix, Iy): lists of padded float32 arrays for levels 0..lv_f
    (interleaved (H, W, 3) arrays for a BGR (h, w, 3) input)."""
    L = lib()
    img_u8 = np.ascontiguousarray(img_u8, np.uint8)
    h, w = img_u8.shape[:2]
    noc = 1 if img_u8.ndim == 2 else img_u8.shape[2]
    wp, hp, left, top = padded_size(w, h, lv_f)
    n = lv_f + 1
    I, Ix, Iy = (_fp * n)(), (_fp * n)(), (_fp * n)()
    L.oracle_build_pyramid_c(img_u8.ctypes.data_as(ctypes.c_void_p), w, h, img_u8.strides[0], noc, lv_f, pad,
                             I, Ix if grads else None, Iy if grads else None)
    out = []
    for arr in (I, Ix, Iy):
        lst = []
        for l in range(n):
            if not arr[l]:
                lst.append(None)
                continue
            shape = ((hp >> l) + 2 * pad, (wp >> l) + 2 * pad) + ((noc,) if noc > 1 else ())
            a = np.ctypeslib.as_array(arr[l], shape=shape).copy()
            L.oracle_free(ctypes.cast(arr[l], ctypes.c_void_p))
            lst.append(a)
        out.append(lst)
    return tuple(out)


def padded_size(w, h, lv_f):
    L = lib()
    v = [ctypes.c_int() for _ in range(4)]
    L.oracle_padded_size(w, h, lv_f, *[ctypes.byref(x) for x in v])
    return tuple(x.value for x in v)


def _ptrs(arrs):
    a = (_fp * len(arrs))()
    for i, x in enumerate(arrs):
        a[i] = x.ctypes.data_as(_fp) if x is not None else None
    return a


def run_engine(pyr_a, pyr_b, w_pad, h_pad, params, initflow=None, taps=False, noc=1):
    """E1 restatement (= OFC::OFClass ctor). pyr_* = (I, Ix, Iy) lists. Returns flow at level lv_l;
    with taps=True also ({level: patch_flow}, {level: dense_flow_before_refinement})."""
    L = lib()
    q = DisParams.from_dict(params)
    sc = 2 ** q.lv_l
    flow = np.zeros((h_pad // sc, w_pad // sc, 2), np.float32)
    n = q.lv_f + 1
    tp, td = ((_fp * n)(), (_fp * n)()) if taps else (None, None)
    ptrs = [_ptrs(x) for x in (*pyr_a, *pyr_b)]
    L.oracle_engine(*ptrs, q.patchsz, flow.ctypes.data_as(_fp),
                    initflow.ctypes.data_as(_fp) if initflow is not None else None, w_pad, h_pad,
                    ctypes.byref(q), tp, td, noc)
    if not taps:
        return flow
    steps = max(1, int(np.floor(q.patchsz * (1 - q.poverl))))
    pf, dn = {}, {}
    for l in range(q.lv_l, q.lv_f + 1):
        wl, hl = w_pad >> l, h_pad >> l
        nop = int(np.ceil(np.float32(wl) / np.float32(steps))) * int(np.ceil(np.float32(hl) / np.float32(steps)))
        pf[l] = np.ctypeslib.as_array(tp[l], shape=(nop, 2)).copy()
        dn[l] = np.ctypeslib.as_array(td[l], shape=(hl, wl, 2)).copy()
        L.oracle_free(ctypes.cast(tp[l], ctypes.c_void_p))
        L.oracle_free(ctypes.cast(td[l], ctypes.c_void_p))
    return flow, pf, dn


def run_u8(a_u8, b_u8, params, want_level=False):
    """Whole run_dense data path restated (P1 + E1 + O1). Returns full-res flow (h, w, 2)
    [and the raw level-lv_l engine output].  Grey (h, w) or interleaved BGR (h, w, 3) u8 input."""
    L = lib()
    a_u8 = np.ascontiguousarray(a_u8, np.uint8)
    b_u8 = np.ascontiguousarray(b_u8, np.uint8)
    h, w = a_u8.shape[:2]
    noc = 1 if a_u8.ndim == 2 else a_u8.shape[2]
    q = DisParams.from_dict(params)
    wp, hp, _, _ = padded_size(w, h, q.lv_f)
    flow = np.zeros((h, w, 2), np.float32)
    lvl = np.zeros((hp >> q.lv_l, wp >> q.lv_l, 2), np.float32)
    L.oracle_run_u8c(a_u8.ctypes.data_as(ctypes.c_void_p), b_u8.ctypes.data_as(ctypes.c_void_p), w, h,
                     a_u8.strides[0], noc, ctypes.byref(q), flow.ctypes.data_as(_fp), lvl.ctypes.data_as(_fp))
    return (flow, lvl) if want_level else flow


def finish(level_flow, lv_l, left, top, w_org, h_org):
    """O1 restatement: x2^lv_l, bilinear upsample, crop."""
    L = lib()
    level_flow = np.ascontiguousarray(level_flow, np.float32)
    hl, wl = level_flow.shape[:2]
    out = np.zeros((h_org, w_org, 2), np.float32)
    L.oracle_finish(level_flow.ctypes.data_as(_fp), wl, hl, lv_l, left, top, w_org, h_org,
                    out.ctypes.data_as(_fp))
    return out
