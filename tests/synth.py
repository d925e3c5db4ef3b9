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


def c5_matrix(t, w, h):
    """rotation 0.05 deg*t about the centre x scale 1+1e-4 t + shift (0.9t,-0.4t)."""
    return affine(w, h, rot_deg=0.05 * t, scale=1.0 + 1e-4 * t, shift=(0.9 * t, -0.4 * t))


def c5_frame(base, k):
    """Frame k of the C5 stream: triangle-wave affine trajectory (period 128 frames) of `base`."""
    h, w = base.shape[:2]
    t = k % 64 if (k // 64) % 2 == 0 else 64 - (k % 64)
    return warp(base, c5_matrix(t, w, h))


def c5_frames(base, n_frames):
    return np.stack([c5_frame(base, k) for k in range(n_frames)])


def load_gray(name):
    """Committed first frames (tests/golden/*.png, written by tests/golden/make_golden_full.py)."""
    import cv2
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    im = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    if im is None:
        raise FileNotFoundError(path)
    return im


def stream_frame(base, k, rot_deg, dscale, shift, period=64):
    """Frame k of a triangle-wave affine trajectory: step t of the trajectory applies rotation rot_deg*t,
    scale 1 + dscale*t and shift*t to `base` (c5_frame is stream_frame(base, k, 0.05, 1e-4, (0.9, -0.4)))."""
    h, w = base.shape[:2]
    t = k % period if (k // period) % 2 == 0 else period - (k % period)
    return warp(base, affine(w, h, rot_deg=rot_deg * t, scale=1.0 + dscale * t, shift=(shift[0] * t, shift[1] * t)))
