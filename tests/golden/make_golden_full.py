"""Generates the FULL-SIZE fixtures of the BASELINE.json configs (run in the build container, where
/root/reference exists; the GPU box only reads what this script wrote).

  road_HD_gray.png, yosemite_4k_gray.png
      images/road_HD.jpg and images/yosemite_4k.jpg decoded ONCE with cv2.imread(IMREAD_GRAYSCALE) -- the
      reference's own decode (kroeger/run_dense.cpp:208-209) -- stored losslessly.  They are the first
      frames of C3 / C4 / C5; second frames are derived with tests/synth.warp
      (cv2.warpAffine, fixed-point INTER_LINEAR on u8) and pinned by the sha256 recorded below.
  full_digests.npz
      for C2, C3, C4a, C4b and four pairs of the C5 stream: the raw level-lv_l flow of the
      verbatim-compiled reference engine (oracle/_ref/libdis_ref.so) driven by oracle/ref_driver.py,
      reduced to  sha256(flow bytes)  +  flow[::16, ::16]  +  sha256 of both input frames.

Usage:  python tests/golden/make_golden_full.py [c2 c3 c4a c4b c5]     (default: all; c4b takes minutes)
"""
import hashlib
import os
import sys
import time

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_driver as rd  # noqa: E402
from tests import synth  # noqa: E402

REF = os.environ.get("DIS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "full_digests.npz")

ARGV = {
    "c2": "5 0 128 128 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0",
    "c3": "6 2 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0",
    "c4a": "7 0 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0",
    "c4b": "7 0 128 128 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0",
}
ARGV["c5"] = ARGV["c3"]
C5_PAIRS = (0, 1, 37, 63)  # pair k = frames k -> k+1 of the triangle-wave stream (tests/synth.c5_frame)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs(cfg):
    """-> list of (key, a_u8, b_u8)."""
    g = lambda n: cv2.imread(os.path.join(HERE, n), cv2.IMREAD_GRAYSCALE)
    if cfg == "c2":
        return [("c2", g("alley_0001_gray.png"), g("alley_0002_gray.png"))]
    if cfg == "c3":
        a = g("road_HD_gray.png")
        return [("c3", a, synth.warp(a, synth.affine(a.shape[1], a.shape[0])))]
    if cfg in ("c4a", "c4b"):
        a = g("yosemite_4k_gray.png")
        return [(cfg, a, synth.warp(a, synth.affine(a.shape[1], a.shape[0])))]
    if cfg == "c5":
        base = g("road_HD_gray.png")
        return [("c5_%d" % k, synth.c5_frame(base, k), synth.c5_frame(base, k + 1)) for k in C5_PAIRS]
    raise KeyError(cfg)


def main():
    for n in ("road_HD", "yosemite_4k"):
        dst = os.path.join(HERE, n + "_gray.png")
        if not os.path.exists(dst):
            im = cv2.imread(os.path.join(REF, "images", n + ".jpg"), cv2.IMREAD_GRAYSCALE)
            cv2.imwrite(dst, im, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    todo = [a for a in sys.argv[1:] if a in ARGV] or list(ARGV)
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for cfg in todo:
        p = rd.parse_params(ARGV[cfg].split())
        for key, a, b in inputs(cfg):
            t0 = time.perf_counter()
            fl = rd.run_dense_ref(a, b, p, full_res=False)
            out[key + "_sha"] = np.array(sha(fl))
            out[key + "_sub"] = np.ascontiguousarray(fl[::16, ::16])
            out[key + "_shape"] = np.array(fl.shape, np.int32)
            out[key + "_in_sha"] = np.array([sha(a), sha(b)])
            out[key + "_argv"] = np.array(ARGV[cfg])
            print(key, fl.shape, out[key + "_sha"], "%.1f s" % (time.perf_counter() - t0), flush=True)
            np.savez_compressed(OUT, **out)


if __name__ == "__main__":
    main()
