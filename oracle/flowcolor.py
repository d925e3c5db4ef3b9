"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's Middlebury colour coding and a ctypes
binding of its verbatim build (oracle/_ref/libcolor_ref.so).  Only tests/ may import this module.

Follows flow_code/C/colorcode.cpp:26-77 (makecolorwheel, computeColor) and the loop of
flow_code/C/color_flow.cpp:19-71 (MotionToColor), including C's float/double promotions.  Pinned by the
reference's own known-answer pair kroeger/flows/alley_0001.flo -> alley_0001.png (tests/golden).

One libm dependency: computeColor's `atan2(-fy, -fx)` on floats is atan2f (C++ overload), whose last-bit
rounding depends on the host libm (glibc 2.39's is not correctly rounded).  This restatement and the CUDA
kernel use the correctly rounded float arctangent (double atan2 rounded to float); against the verbatim
build on this image that changes one grey level on ~2 pixels per million and nothing on the golden pair.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
f32, f64 = np.float32, np.float64


def colorwheel():
    """colorcode.cpp:26-51 -> (55, 3) int array of (r, g, b)."""
    RY, YG, GC, CB, BM, MR = 15, 6, 4, 11, 13, 6
    w = []
    w += [(255, 255 * i // RY, 0) for i in range(RY)]
    w += [(255 - 255 * i // YG, 255, 0) for i in range(YG)]
    w += [(0, 255, 255 * i // GC) for i in range(GC)]
    w += [(0, 255 - 255 * i // CB, 255) for i in range(CB)]
    w += [(255 * i // BM, 0, 255) for i in range(BM)]
    w += [(255, 0, 255 - 255 * i // MR) for i in range(MR)]
    return np.array(w, np.int32)


def unknown_flow(fl):
    """flow_code/C/flowIO.cpp:35-39."""
    u, v = fl[..., 0], fl[..., 1]
    with np.errstate(invalid="ignore"):
        return (np.abs(u.astype(f64)) > 1e9) | (np.abs(v.astype(f64)) > 1e9) | np.isnan(u) | np.isnan(v)


def motion_to_color(flow, maxmotion=-1.0):
    """MotionToColor: (h, w, 2) float32 -> ((h, w, 3) u8 in B,G,R byte order, maxrad used)."""
    flow = np.ascontiguousarray(flow, f32)
    W = colorwheel()
    nc = len(W)
    unk = unknown_flow(flow)
    fx = np.where(unk, f32(0), flow[..., 0]).astype(f32)
    fy = np.where(unk, f32(0), flow[..., 1]).astype(f32)
    rad = np.sqrt((fx * fx + fy * fy).astype(f64)).astype(f32)
    maxrad = f32(rad[~unk].max()) if (~unk).any() else f32(-1)
    if maxmotion > 0:
        maxrad = f32(maxmotion)
    if maxrad == 0:
        maxrad = f32(1)
    fx = (fx / maxrad).astype(f32)
    fy = (fy / maxrad).astype(f32)
    rad = np.sqrt((fx * fx + fy * fy).astype(f64)).astype(f32)
    # C++ overload resolution: atan2(float, float) is atan2f (flow_code/C/colorcode.cpp:59 is compiled as C++),
    # restated as the correctly rounded float of the double result; the division by M_PI is in double
    at = np.arctan2((-fy).astype(f64), (-fx).astype(f64)).astype(f32)
    a = (at.astype(f64) / np.pi).astype(f32)
    fk = ((a.astype(f64) + 1.0) / 2.0 * (nc - 1)).astype(f32)
    k0 = fk.astype(np.int32)
    k1 = (k0 + 1) % nc
    f = (fk - k0.astype(f32)).astype(f32)
    out = np.zeros(flow.shape[:2] + (3,), np.uint8)
    for b in range(3):
        col0 = (W[k0, b] / 255.0).astype(f32)
        col1 = (W[k1, b] / 255.0).astype(f32)
        col = ((f32(1) - f) * col0 + f * col1).astype(f32)
        col = np.where(rad <= 1, (f32(1) - rad * (f32(1) - col)).astype(f32), (col.astype(f64) * .75).astype(f32))
        out[..., 2 - b] = (255.0 * col.astype(f64)).astype(np.int32).astype(np.uint8)
    out[unk] = 0
    return out, float(maxrad)


def epe(flow_a, flow_b, margin=0):
    """mean, max, count of |a - b|_2 over known pixels at least `margin` from the border (float64)."""
    a = np.asarray(flow_a, f32)
    b = np.asarray(flow_b, f32)
    h, w = a.shape[:2]
    sl = (slice(margin, h - margin), slice(margin, w - margin))
    a, b = a[sl], b[sl]
    ok = ~(unknown_flow(a) | unknown_flow(b))
    d = a[ok].astype(f64) - b[ok].astype(f64)
    e = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    return (float(e.mean()) if e.size else 0.0), (float(e.max()) if e.size else 0.0), int(e.size)


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libcolor_ref.so"))


def motion_to_color_ref(flow, maxmotion=-1.0):
    """The verbatim computeColor of the reference (oracle/_ref/libcolor_ref.so)."""
    L = ctypes.CDLL(os.path.join(_HERE, "_ref", "libcolor_ref.so"))
    L.color_ref_image.restype = ctypes.c_float
    L.color_ref_image.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    flow = np.ascontiguousarray(flow, f32)
    h, w = flow.shape[:2]
    out = np.zeros((h, w, 3), np.uint8)
    mr = L.color_ref_image(flow.ctypes.data, w, h, float(maxmotion), out.ctypes.data)
    return out, float(mr)
