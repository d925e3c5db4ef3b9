"""One lone SOR item (960x32 image, 1 sweep, 1 row block) for ncu source-level sampling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from flowonthego_b200 import api
from tests.synth import synth_pair
grp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w, h = 960, 32
a, b, _ = synth_pair(w, h, seed=1)
p = F.Params.preset(2, 1024, verbosity=0).copy(lv_f=0, lv_l=0, tv_solverit=1, tv_innerit=1, patchsz=8)
with F.Engine(p, w, h) as e:
    e.set_option(api.OPT_SOR_GROUP, grp)
    e.set_option(api.OPT_USE_GRAPH, 0)
    for _ in range(3):
        e.run_u8(a, b)
