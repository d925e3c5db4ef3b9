"""Per-kernel device times of one 1080p pair in colour mode (channels=3) next to the grey engine."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flowonthego_b200 as F
from tests.synth import synth_pair, synth_pair_bgr
w, h = 1920, 1080
p = F.Params.preset(3, 1920, verbosity=0)
for ch in (1, 3):
    a, b, gt = (synth_pair if ch == 1 else synth_pair_bgr)(w, h, seed=1)
    with F.Engine(p, w, h, channels=ch) as e:
        for _ in range(3):
            fl = e.run_u8(a, b)
        e.enable_kernel_profile(True)
        for _ in range(5):
            e.run_u8(a, b)
        agg = collections.defaultdict(float)
        for k in e.kernel_profile():
            agg[k["name"]] += k["ms"] / 5
        e.enable_kernel_profile(False)
        print("channels", ch, "epe", float(np.abs(fl - gt)[32:-32, 32:-32].mean()),
              {k: round(v, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1])}, "sum", round(sum(agg.values()), 3))
