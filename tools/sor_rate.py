"""Natural rates of k_sor_wavefront: one item alone (step time), two row blocks (hand-off lag), two sweeps
(inter-sweep lag).  Engine on a W x H image at a single level, kernel profile -> SOR launch time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flowonthego_b200 as F
from flowonthego_b200 import api
from tests.synth import synth_pair
def sor_ms(w, h, T, grp):
    a, b, _ = synth_pair(w, h, seed=1)
    p = F.Params.preset(2, 1024, verbosity=0).copy(lv_f=0, lv_l=0, tv_solverit=T, tv_innerit=1, patchsz=8)
    with F.Engine(p, w, h) as e:
        e.set_option(api.OPT_SOR_GROUP, grp)
        for _ in range(2): e.run_u8(a, b)
        e.enable_kernel_profile(True)
        n = 5
        for _ in range(n): e.run_u8(a, b)
        ms = sum(r["ms"] for r in e.kernel_profile() if r["name"] == "k_sor_wavefront") / n
    return ms
for grp in [int(x) for x in sys.argv[1:]] or (8, 16):
    base = None
    for (w, h, T) in ((30, 17, 3), (60, 34, 3), (120, 68, 3), (240, 136, 3),  # the coarse levels of a 1080p pair
                   (480, 32, 1), (960, 32, 1), (480, 64, 1), (480, 32, 2), (480, 32, 3), (480, 96, 1), (480, 288, 1), (480, 288, 3)):
        ms = sor_ms(w, h, T, grp)
        steps = w + 31
        print("G=%2d  %4dx%-3d T=%d  K=%d: %7.1f us  = %.0f cycles per step of one item" % (grp, w, h, T, (h + 31) // 32, ms * 1e3, ms * 1e-3 * 1.965e9 / steps), flush=True)
