import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flowonthego_b200 as F
from oracle import port
from tests.synth import synth_pair
for (w, h, kw) in ((250, 190, dict(lv_f=3, lv_l=0, patchsz=12, poverl=0.75, maxiter=6, miniter=6)),
                   (97, 61, dict(lv_f=1, lv_l=0, patchsz=8, maxiter=4, miniter=4)),
                   (160, 120, dict(lv_f=2, lv_l=1, usefbcon=1, maxiter=4, miniter=4))):
    a, b, _ = synth_pair(w, h, seed=w, shift=(5.5, -3.25), rot_deg=1.0)
    p = F.Params.preset(2, w, verbosity=0).copy(**kw)
    with F.Engine(p, w, h) as e:
        f = e.run_u8(a, b)
    r = port.run_u8(a, b, p.to_dict())
    print(w, h, "bits differ:", int((f.view(np.uint32) != r.view(np.uint32)).sum()), flush=True)
