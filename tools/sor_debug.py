import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from tests.synth import synth_pair
a, b, _ = synth_pair(1920, 1080, seed=1)
p = F.Params.preset(3, 1920, verbosity=0)
with F.Engine(p, 1920, 1080) as e:
    e.enable_kernel_profile(True)  # un-graphed
    e.run_u8(a, b)
