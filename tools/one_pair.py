"""One un-graphed pair for ncu: `c3` (road_HD + C3 warp, 1080p, operating point 3) or `c4` (yosemite_4k, C4a)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from tests import synth
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
if which == "c4":
    a = synth.load_gray("yosemite_4k_gray.png")
    p = F.Params.from_argv("7 0 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0".split())
else:
    a = synth.load_gray("road_HD_gray.png")
    p = F.Params.preset(3, 1920, verbosity=0)
h, w = a.shape
b = synth.warp(a, synth.affine(w, h))
with F.Engine(p, w, h) as e:
    e.enable_kernel_profile(True)  # un-graphed
    e.run_u8(a, b)
