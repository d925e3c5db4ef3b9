"""One batched call (8 pairs per launch, 1080p preset 3) for ncu: per-kernel time with 8x the work per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from flowonthego_b200 import api
from tests.synth import synth_pair
import torch
w, h, nb = 1920, 1080, 8
p = F.Params.preset(3, 1920, verbosity=0)
a, b, _ = synth_pair(w, h, seed=1)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
do = torch.empty((nb, h, w, 2), dtype=torch.float32, device="cuda")
with F.Engine(p, w, h, batch=nb) as e:
    e.set_option(api.OPT_USE_GRAPH, 0)
    for _ in range(2):
        e.submit_u8_device_batch([da.data_ptr()] * nb, [db.data_ptr()] * nb, w, h, w, [do[i].data_ptr() for i in range(nb)])
        e.wait()
