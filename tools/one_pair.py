import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from tests.synth import synth_pair
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
if which == "c4":
    w, h = 3840, 2160
    p = F.Params.from_argv("7 0 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0".split())
else:
    w, h = 1920, 1080
    p = F.Params.preset(3, 1920, verbosity=0)
a, b, _ = synth_pair(w, h, seed=1)
with F.Engine(p, w, h) as e:
    e.enable_kernel_profile(True)  # un-graphed
    e.run_u8(a, b)
