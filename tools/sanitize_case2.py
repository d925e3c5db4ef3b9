"""compute-sanitizer targets added later in round 1: colour engine, batched handle, group, video front end,
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
# group of 2
with F.EngineGroup(p, w, h, 2) as g:
    out.zero_()
    g.submit_u8_device([da[0].data_ptr(), da[1].data_ptr()], [db[0].data_ptr(), db[1].data_ptr()], w, h, w, [out[0].data_ptr(), out[1].data_ptr()])
    g.wait()
    print("group:", [bd(out[k].cpu().numpy(), port.run_u8(pairs[k][0], pairs[k][1], p.to_dict())) for k in range(2)], flush=True)
# video front end
frames = [pairs[0][0], pairs[0][1], pairs[1][1], pairs[2][1]]
with F.FlowStream(p, w, h, depth=2) as s:
    fl = list(s.flows(frames))
print("video:", [bd(fl[k], port.run_u8(frames[k], frames[k + 1], p.to_dict())) for k in range(3)], flush=True)
# colour coding + EPE
f = fl[0]
print("color:", int((F.flow_to_color(f) != flowcolor.motion_to_color(f)[0]).sum()), "epe:", F.flow_epe(f, fl[1], 4)[2], flush=True)
