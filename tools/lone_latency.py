"""Latency of one 1080p (or WxH) pair alone on the GPU, graph replay, by execution options:
DIS_OPT_SOR_GROUP in (0, 16) x DIS_OPT_SOR_SMALL in 0..5.  usage: lone_latency.py [w h]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flowonthego_b200 as F
from flowonthego_b200 import api
from tests.synth import synth_pair
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
a, b, _ = synth_pair(w, h, seed=1)
p = F.Params.from_argv("6 2 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0".split()) if w == 1920 else F.Params.preset(3, w, verbosity=0)
dev = torch.device("cuda:0")
da, db = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
out = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
ref = None
for grp in (0, 16):
    for small in (0, 1, 2, 3, 5, -1):
        with F.Engine(p, w, h) as e:
            e.set_option(api.OPT_SOR_GROUP, grp)
            e.set_option(api.OPT_SOR_SMALL, small)
            st = torch.cuda.ExternalStream(e.stream, device=dev)
            for _ in range(3):
                e.submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, out.data_ptr())
            e.wait()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20):
                e.submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, out.data_ptr())
            e1.record(st)
            e.wait()
            o = out.cpu().numpy()
            ref = o if ref is None else ref
            print("group %2d small %d: %.3f ms  (same flow: %s)" % (grp, small, e0.elapsed_time(e1) / 20, bool(np.array_equal(o, ref))), flush=True)
