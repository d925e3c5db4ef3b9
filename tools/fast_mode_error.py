"""Error of the tolerance mode (DIS_OPT_ARITH = 1) against the exact engine on C3 and C4a."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flowonthego_b200 as F
from flowonthego_b200 import api
from tests.test_full_size import load_case
for key in ("c3", "c4a"):
    a, b, p, dig = load_case(key)
    with F.Engine(F.Params.from_dict(p), a.shape[1], a.shape[0]) as e:
        exact = e.run_u8(a, b).copy()
        e.set_option(api.OPT_ARITH, 1)
        fast = e.run_u8(a, b).copy()
    m = p["patchsz"] << p["lv_l"]
    d = np.abs(fast.astype(np.float64) - exact)[m:-m, m:-m]
    print(key, "mean %.3g max %.3g  frac>1e-2 %.3g  frac>1e-3 %.3g  p99.9 %.3g" % (d.mean(), d.max(), (d > 1e-2).mean(), (d > 1e-3).mean(), np.quantile(d, 0.999)))
