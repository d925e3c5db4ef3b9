"""k_sor_small (one CTA, all sweeps interleaved) against k_sor_wavefront on single levels: kernel-profile time of the
SOR launch and bit-equality of the flow.  usage: sor_small_rate.py [w h]..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flowonthego_b200 as F
from flowonthego_b200 import api
from tests.synth import synth_pair

def run(w, h, small, grp=0):
    a, b, _ = synth_pair(w, h, seed=1)
    p = F.Params.preset(2, 1024, verbosity=0).copy(lv_f=0, lv_l=0, tv_solverit=3, tv_innerit=1, patchsz=8)
    with F.Engine(p, w, h) as e:
        e.set_option(api.OPT_SOR_SMALL, small)
        e.set_option(api.OPT_SOR_GROUP, grp)
        for _ in range(2): flow = e.run_u8(a, b)
        e.enable_kernel_profile(True)
        n = 5
        for _ in range(n): e.run_u8(a, b)
        ms = sum(r["ms"] for r in e.kernel_profile() if r["name"].startswith("k_sor")) / n
    return ms, flow.copy()

args = [int(x) for x in sys.argv[1:]]
shapes = list(zip(args[::2], args[1::2])) or [(30, 17), (60, 34), (120, 68), (240, 136), (480, 272), (128, 55), (64, 27)]
for (w, h) in shapes:
    K = (h + 31) // 32
    m0, f0 = run(w, h, 0)
    m16, f16 = run(w, h, 0, 16)
    m1, f1 = run(w, h, 9) if K <= 9 else (float("nan"), f0)
    steps = (w + 1) // 2 + h - 1 + 4
    print("%4dx%-3d K=%d: wavefront %7.1f us (group 16, one CTA per item: %7.1f), small %7.1f us = %.0f cycles per step; flows equal: %s" %
          (w, h, K, m0 * 1e3, m16 * 1e3, m1 * 1e3, m1 * 1e-3 * 1.965e9 / steps, bool(np.array_equal(f0, f1) and np.array_equal(f0, f16))), flush=True)
