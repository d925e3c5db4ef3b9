"""Synthetic frame pairs with known ground-truth flow: a procedurally
textured grey image and its affine warp.  No reference data is read at run time."""
import numpy as np


def texture(w, h, seed=0):
    """Band-limited multi-octave noise, u8.  Deterministic for a given (w, h, seed)."""
    import cv2
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float32)
    amp = 1.0
    for octave in range(7):
        sw, sh = max(2, w >> (7 - octave)), max(2, h >> (7 - octave))
        n = rng.standard_normal((sh, sw)).astype(np.float32)
        img += amp * cv2.resize(n, (w, h), interpolation=cv2.INTER_CUBIC)
        amp *= 0.6
    img -= img.min()
    img *= 255.0 / max(float(img.max()), 1e-6)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def affine(w, h, rot_deg=0.4, scale=1.004, shift=(3.5, -2.25)):
    """2x3 matrix: rotation about the image centre x scale + shift (C3/C4 of BASELINE.md)."""
    c, s = np.cos(np.deg2rad(rot_deg)) * scale, np.sin(np.deg2rad(rot_deg)) * scale
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    return np.array([[c, -s, cx - c * cx + s * cy + shift[0]], [s, c, cy - s * cx - c * cy + shift[1]]], np.float64)


def warp(img, M):
    """Second frame b with b(M x) = a(x): content moves by gt(x) = M x - x."""
    import cv2
    return cv2.warpAffine(img, M.astype(np.float32), (img.shape[1], img.shape[0]), flags=cv2.INTER_LINEAR,
                          borderMode=cv2.BORDER_REPLICATE)


def gt_flow(w, h, M):
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    u = M[0, 0] * xs + M[0, 1] * ys + M[0, 2] - xs
    v = M[1, 0] * xs + M[1, 1] * ys + M[1, 2] - ys
    return np.stack([u, v], -1).astype(np.float32)


def synth_pair(w, h, seed=0, **kw):
    a = texture(w, h, seed)
    M = affine(w, h, **kw)
    return a, warp(a, M), gt_flow(w, h, M)


def synth_pair_bgr(w, h, seed=0, **kw):
    """Colour pair: three differently textured channels moved by the same affine flow."""
    M = affine(w, h, **kw)
    a = np.stack([texture(w, h, seed * 3 + k + 100) for k in range(3)], -1)
    b = np.stack([warp(np.ascontiguousarray(a[..., k]), M) for k in range(3)], -1)
    return np.ascontiguousarray(a), np.ascontiguousarray(b), gt_flow(w, h, M)
