"""Development probe (run under gpurun): stage-by-stage comparison of libdis_b200.so with the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flowonthego_b200 as F  # noqa: E402
from flowonthego_b200 import api  # noqa: E402
from oracle import port  # noqa: E402
from tests.synth import synth_pair  # noqa: E402


def ne(x, y):
    return int((np.ascontiguousarray(x).view(np.uint32) != np.ascontiguousarray(y).view(np.uint32)).sum())


def stats(name, g, o):
    d = np.abs(g.astype(np.float64) - o.astype(np.float64))
    print("  %-28s bits!= %8d / %8d   max %.3e  mean %.3e" % (name, ne(g, o), g.size, d.max(), d.mean()), flush=True)


def probe(a, b, p, label):
    print("==", label, a.shape, {k: v for k, v in p.to_dict().items() if k in ("lv_f", "lv_l", "patchsz", "maxiter", "usetvref", "costfct", "poverl")}, flush=True)
    h, w = a.shape
    wp, hp, left, top = F.padded_size(w, h, p.lv_f)
    t = time.time()
    pa = port.build_pyramid(a, p.lv_f, p.patchsz)
    pb = port.build_pyramid(b, p.lv_f, p.patchsz)
    fo, pf, dn = port.run_engine(pa, pb, wp, hp, p.to_dict(), taps=True)
    full_o = port.finish(fo, p.lv_l, left, top, w, h)
    print("  oracle %.2fs" % (time.time() - t), flush=True)
    with F.Engine(p, w, h, 0) as e:
        e.enable_taps(True)
        t = time.time()
        full_g = e.run_u8(a, b)
        print("  gpu (taps) %.3fs" % (time.time() - t), flush=True)
        for l in range(p.lv_l, p.lv_f + 1):
            stats("I_a   L%d" % l, e.tap(api.TAP_IMG_A, l), pa[0][l].ravel())
            stats("Ix_a  L%d" % l, e.tap(api.TAP_IMG_A_DX, l), pa[1][l].ravel())
            stats("Iy_a  L%d" % l, e.tap(api.TAP_IMG_A_DY, l), pa[2][l].ravel())
            stats("I_b   L%d" % l, e.tap(api.TAP_IMG_B, l), pb[0][l].ravel())
        for l in range(p.lv_f, p.lv_l - 1, -1):
            stats("patch flow L%d" % l, e.tap(api.TAP_PATCH_FLOW, l), pf[l].ravel())
            stats("dense flow L%d" % l, e.tap(api.TAP_FLOW_DENSE, l), dn[l].ravel())
        stats("engine out", e.level_flow(w, h), fo)
        stats("full-res", full_g, full_o)
        e.enable_taps(False)
        t = time.time()
        full_g2 = e.run_u8(a, b)
        t1 = time.time() - t
        t = time.time()
        for _ in range(5):
            full_g2 = e.run_u8(a, b)
        t2 = (time.time() - t) / 5
        stats("full-res (graph)", full_g2, full_o)
        print("  graph first %.4fs, steady %.4fs/pair, launches %d" % (t1, t2, e.timings()["launches"]), flush=True)
    return full_g


if __name__ == "__main__":
    which = sys.argv[1:] or ["small"]
    if "small" in which:
        a, b, _ = synth_pair(256, 192, seed=3)
        probe(a, b, F.Params.preset(2, 256, verbosity=0).copy(lv_f=3, lv_l=1, usetvref=0), "small notv")
        probe(a, b, F.Params.preset(2, 256, verbosity=0).copy(lv_f=3, lv_l=1), "small tv")
        probe(a, b, F.Params.preset(3, 256, verbosity=0).copy(lv_f=3, lv_l=0), "small p12")
    if "alley" in which:
        import cv2
        a = cv2.imread(os.path.join(ROOT, "tests/golden/alley_0001_gray.png"), cv2.IMREAD_UNCHANGED)
        b = cv2.imread(os.path.join(ROOT, "tests/golden/alley_0002_gray.png"), cv2.IMREAD_UNCHANGED)
        g = np.load(os.path.join(ROOT, "tests/golden/alley_0001_flo.npz"))["flow"]
        f = probe(a, b, F.Params.preset(2, 1024, verbosity=0), "alley preset2 (golden)")
        stats("vs golden .flo", f, g)
        probe(a, b, F.Params.preset(3, 1024, verbosity=0), "alley preset3")
    if "hd" in which:
        a, b, _ = synth_pair(1920, 1080, seed=1)
        probe(a, b, F.Params.preset(3, 1920, verbosity=0), "1080p preset3")
