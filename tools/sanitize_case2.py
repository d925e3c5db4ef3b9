"""compute-sanitizer targets added later in round 1: colour engine, batched handle, level export, tolerance mode, video front end (with pyramid reuse),
colour coding / EPE, level-2 upsampling kernel.  Small sizes; every result is checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import flowonthego_b200 as F
from oracle import port, flowcolor
from tests.synth import synth_pair, synth_pair_bgr
def bd(x, y): return int((np.ascontiguousarray(x).view(np.uint32) != np.ascontiguousarray(y).view(np.uint32)).sum())
# colour, lv_l = 2 (block-mean pyramid + 4x4 upsampling kernel)
a, b, _ = synth_pair_bgr(163, 123, seed=3)
p = F.Params.preset(2, 256, verbosity=0).copy(lv_f=2, lv_l=2, maxiter=4, miniter=4, patchsz=8)
with F.Engine(p, 163, 123, channels=3) as e:
    print("rgb lv2:", bd(e.run_u8(a, b), port.run_u8(a, b, p.to_dict())), flush=True)
# batched handle (3 pairs per launch) with forward-backward merging
w, h, nb = 130, 98, 3
p = F.Params.preset(2, 256, verbosity=0).copy(lv_f=2, lv_l=1, maxiter=4, miniter=4, usefbcon=1)
pairs = [synth_pair(w, h, seed=10 + k)[:2] for k in range(nb)]
da = [torch.from_numpy(x[0]).cuda() for x in pairs]; db = [torch.from_numpy(x[1]).cuda() for x in pairs]
out = torch.zeros((nb, h, w, 2), dtype=torch.float32, device="cuda")
with F.Engine(p, w, h, batch=nb) as e:
    e.submit_u8_device_batch([x.data_ptr() for x in da], [x.data_ptr() for x in db], w, h, w, [out[k].data_ptr() for k in range(nb)])
    e.wait()
    print("batched:", [bd(out[k].cpu().numpy(), port.run_u8(pairs[k][0], pairs[k][1], p.to_dict())) for k in range(nb)], flush=True)
# level-flow export folded into the finish kernel + tolerance-mode kernels (DIS_OPT_ARITH)
with F.Engine(p, w, h, batch=nb) as e:
    stage = torch.zeros((nb,) + e.level_flow_shape(), dtype=torch.float32, device="cuda")
    e.set_level_export([stage[k].data_ptr() for k in range(nb)])
    e.submit_u8_device_batch([x.data_ptr() for x in da], [x.data_ptr() for x in db], w, h, w, [out[k].data_ptr() for k in range(nb)])
    e.wait()
    print("level export:", [bd(stage[k].cpu().numpy(), port.run_u8(pairs[k][0], pairs[k][1], p.to_dict(), want_level=True)[1]) for k in range(nb)], flush=True)
with F.Engine(p, w, h) as e:
    from flowonthego_b200 import api
    e.set_option(api.OPT_ARITH, 1)
    d = np.abs(e.run_u8(*pairs[0]) - port.run_u8(pairs[0][0], pairs[0][1], p.to_dict()))
    print("fast mode: max |d| %.2g" % d.max(), flush=True)
# video front end
frames = [pairs[0][0], pairs[0][1], pairs[1][1], pairs[2][1]]
with F.FlowStream(p, w, h, depth=2, output="full") as s:  # depth 2: pyramids reused between consecutive pairs
    assert s.reuse
    fl = list(s.flows(frames))
print("video:", [bd(fl[k], port.run_u8(frames[k], frames[k + 1], p.to_dict())) for k in range(3)], flush=True)
# colour coding + EPE
f = fl[0]
print("color:", int((F.flow_to_color(f) != flowcolor.motion_to_color(f)[0]).sum()), "epe:", F.flow_epe(f, fl[1], 4)[2], flush=True)
