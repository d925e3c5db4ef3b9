"""One SOR launch for ncu source-level sampling: W x H level, T sweeps, SOR variant (DIS_OPT_SOR_GROUP value),
row blocks up to which the one-CTA k_sor_small runs (DIS_OPT_SOR_SMALL, default 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from flowonthego_b200 import api
from tests.synth import synth_pair
w, h, T, grp = [int(x) for x in (sys.argv[1:5] + ["960", "32", "1", "16"][len(sys.argv[1:5]):])]
small = int(sys.argv[5]) if len(sys.argv) > 5 else 0
a, b, _ = synth_pair(w, h, seed=1)
p = F.Params.preset(2, 1024, verbosity=0).copy(lv_f=0, lv_l=0, tv_solverit=T, tv_innerit=1, patchsz=8)
with F.Engine(p, w, h) as e:
    e.set_option(api.OPT_SOR_GROUP, grp)
    e.set_option(api.OPT_USE_GRAPH, 0)
    e.set_option(api.OPT_SOR_SMALL, small)
    for _ in range(3):
        e.run_u8(a, b)
