"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Python restatement of the reference's CLI driver ``kroeger/run_dense.cpp:main`` (lines
185-431) around the *verbatim-compiled* reference engine ``oracle/_ref/libdis_ref.so``
(see oracle/Makefile).  The OpenCV C++ SDK the reference links against is absent from this
image; ``cv2`` (4.13) provides the same ``resize`` / ``Sobel`` / ``copyMakeBorder`` calls.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.
"""
import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

# kroeger/run_dense.cpp:271-291 -- order of the 20 explicit CLI parameters
PARAM_NAMES = ("lv_f", "lv_l", "maxiter", "miniter", "mindprate", "mindrrate", "minimgerr",
               "patchsz", "poverl", "usefbcon", "patnorm", "costfct", "usetvref", "tv_alpha",
               "tv_gamma", "tv_delta", "tv_innerit", "tv_solverit", "tv_sor", "verbosity")
_INT_PARAMS = {"lv_f", "lv_l", "maxiter", "miniter", "patchsz", "usefbcon", "patnorm", "costfct",
               "usetvref", "tv_innerit", "tv_solverit", "verbosity"}


def ref_lib(rgb=False):
    """Load oracle/_ref/libdis_ref[_rgb].so (built by ``make -C oracle ref``)."""
    key = bool(rgb)
    if key not in _LIBS:
        path = os.path.join(_HERE, "_ref", "libdis_ref_rgb.so" if rgb else "libdis_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " missing: run `make -C oracle ref` where /root/reference exists")
        lib = ctypes.CDLL(path)
        pp = ctypes.POINTER(ctypes.POINTER(ctypes.c_float))
        fp = ctypes.POINTER(ctypes.c_float)
        lib.dis_ref_ofclass.restype = None
        lib.dis_ref_ofclass.argtypes = [pp] * 6 + [ctypes.c_int, fp, fp] + [ctypes.c_int] * 6 + \
            [ctypes.c_float] * 3 + [ctypes.c_int, ctypes.c_float] + [ctypes.c_int] * 5 + \
            [ctypes.c_float] * 3 + [ctypes.c_int] * 2 + [ctypes.c_float, ctypes.c_int]
        _LIBS[key] = lib
    return _LIBS[key]


def auto_first_scale(width, fratio, patchsz):
    """kroeger/run_dense.cpp:180-183 (float arithmetic)."""
    v = np.float32(2.0) * np.float32(width) / (np.float32(fratio) * np.float32(patchsz))
    return max(0, int(math.floor(math.log2(float(v)))))


def preset_params(width_org, preset=2):
    """Operating points of kroeger/run_dense.cpp:225-267 (verbosity forced to 0)."""
    p = dict(mindprate=0.05, mindrrate=0.95, minimgerr=0.0, usefbcon=0, patnorm=1, costfct=0,
             tv_alpha=10.0, tv_gamma=10.0, tv_delta=5.0, tv_innerit=1, tv_solverit=3, tv_sor=1.6,
             verbosity=0)
    fratio = 5
    if preset == 1:
        p.update(patchsz=8, poverl=0.3, maxiter=16, miniter=16, usetvref=0)
        back = 2
    elif preset == 3:
        p.update(patchsz=12, poverl=0.75, maxiter=16, miniter=16, usetvref=1)
        back = 4
    elif preset == 4:
        p.update(patchsz=12, poverl=0.75, maxiter=128, miniter=128, usetvref=1)
        back = 5
    else:
        p.update(patchsz=8, poverl=0.4, maxiter=12, miniter=12, usetvref=1)
        back = 2
    p["lv_f"] = auto_first_scale(width_org, fratio, p["patchsz"])
    p["lv_l"] = max(p["lv_f"] - back, 0)
    return p


def parse_params(argv20):
    """20 explicit parameters in CLI order (strings or numbers) -> dict."""
    if len(argv20) != 20:
        raise ValueError("need 20 parameters")
    out = {}
    for name, v in zip(PARAM_NAMES, argv20):
        out[name] = int(float(v)) if name in _INT_PARAMS else float(v)
    return out


def pad_geometry(w, h, lv_f):
    """kroeger/run_dense.cpp:298-311 -> (padw, padh, left, top)."""
    sc = 2 ** lv_f
    padw = (sc - w % sc) % sc
    padh = (sc - h % sc) % sc
    return padw, padh, padw // 2, padh // 2


def build_pyramids_cv2(img_u8, params):
    """run_dense.cpp:298-311 (divisibility pad), :326-327 (convertTo) and ConstructImgPyramide
    (:130-178) with cv2.  Returns (I, Ix, Iy) lists of padded float32 arrays, levels 0..lv_f."""
    import cv2
    lv_f, ps = params["lv_f"], params["patchsz"]
    h, w = img_u8.shape[:2]
    padw, padh, left, top = pad_geometry(w, h, lv_f)
    if padw or padh:
        img_u8 = cv2.copyMakeBorder(img_u8, top, padh - top, left, padw - left, cv2.BORDER_REPLICATE)
    cur = img_u8.astype(np.float32)
    I, Ix, Iy = [], [], []
    for lv in range(lv_f + 1):
        if lv > 0:
            cur = cv2.resize(cur, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR)
        dx = cv2.Sobel(cur, cv2.CV_32F, 1, 0, ksize=1, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
        dy = cv2.Sobel(cur, cv2.CV_32F, 0, 1, ksize=1, scale=1, delta=0, borderType=cv2.BORDER_DEFAULT)
        I.append(np.ascontiguousarray(cv2.copyMakeBorder(cur, ps, ps, ps, ps, cv2.BORDER_REPLICATE)))
        Ix.append(np.ascontiguousarray(cv2.copyMakeBorder(dx, ps, ps, ps, ps, cv2.BORDER_CONSTANT, value=0)))
        Iy.append(np.ascontiguousarray(cv2.copyMakeBorder(dy, ps, ps, ps, ps, cv2.BORDER_CONSTANT, value=0)))
    return I, Ix, Iy


def _ptr_array(arrs):
    fp = ctypes.POINTER(ctypes.c_float)
    a = (fp * len(arrs))()
    for i, x in enumerate(arrs):
        a[i] = x.ctypes.data_as(fp)
    return a


def run_engine(pyr_a, pyr_b, w_pad, h_pad, params, initflow=None, rgb=False):
    """Call the reference OFC::OFClass constructor (the drop-in boundary, kroeger/oflow.h:84-111)
    exactly as run_dense.cpp:391-400 does.  Returns flow (h/2^lv_l, w/2^lv_l, 2) float32."""
    lib = ref_lib(rgb)
    p = params
    sc = 2 ** p["lv_l"]
    flow = np.zeros((h_pad // sc, w_pad // sc, 2), np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    ptrs = [_ptr_array(x) for x in (*pyr_a, *pyr_b)]
    init_p = initflow.ctypes.data_as(fp) if initflow is not None else None
    lib.dis_ref_ofclass(*ptrs, p["patchsz"], flow.ctypes.data_as(fp), init_p, w_pad, h_pad,
                        p["lv_f"], p["lv_l"], p["maxiter"], p["miniter"], p["mindprate"],
                        p["mindrrate"], p["minimgerr"], p["patchsz"], p["poverl"],
                        int(p["usefbcon"]), p["costfct"], 3 if rgb else 1, p["patnorm"],
                        int(p["usetvref"]), p["tv_alpha"], p["tv_gamma"], p["tv_delta"],
                        p["tv_innerit"], p["tv_solverit"], p["tv_sor"], p["verbosity"])
    return flow


def finish_flow_cv2(flow, params, w_org, h_org):
    """run_dense.cpp:407-414: scale by 2^lv_l, cv::resize INTER_LINEAR, crop the padding."""
    import cv2
    sc = 2 ** params["lv_l"]
    if params["lv_l"] != 0:
        flow = flow * np.float32(sc)
        flow = cv2.resize(flow, None, fx=sc, fy=sc, interpolation=cv2.INTER_LINEAR)
    padw, padh, left, top = pad_geometry(w_org, h_org, params["lv_f"])
    return np.ascontiguousarray(flow[top:top + h_org, left:left + w_org])


def run_dense_ref(img_a_u8, img_b_u8, params, full_res=True, rgb=False):
    """The whole reference pipeline on two decoded images.  ``full_res=False`` returns the raw
    engine output at level lv_l (padded size), which is what the C-ABI boundary compares."""
    h, w = img_a_u8.shape[:2]
    padw, padh, _, _ = pad_geometry(w, h, params["lv_f"])
    pa = build_pyramids_cv2(img_a_u8, params)
    pb = build_pyramids_cv2(img_b_u8, params)
    flow = run_engine(pa, pb, w + padw, h + padh, params, rgb=rgb)
    if not full_res:
        return flow
    return finish_flow_cv2(flow, params, w, h)


def read_flo(path):
    """Middlebury .flo (flow_code/C/flowIO.cpp:5-25)."""
    with open(path, "rb") as f:
        tag = f.read(4)
        if tag != b"PIEH":
            raise ValueError("bad .flo tag")
        w, h = np.frombuffer(f.read(8), np.int32)
        return np.frombuffer(f.read(), np.float32).reshape(h, w, 2).copy()


def write_flo(path, flow):
    """kroeger/run_dense.cpp:16-57."""
    h, w = flow.shape[:2]
    with open(path, "wb") as f:
        f.write(b"PIEH")
        f.write(np.array([w, h], np.int32).tobytes())
        f.write(np.ascontiguousarray(flow, np.float32).tobytes())
