import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from tests.synth import synth_pair
import torch
w, h = 1920, 1080
p = F.Params.preset(3, 1920, verbosity=0)
a, b, _ = synth_pair(w, h, seed=1)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
for S in (16, 32, 64, 96, 128, 192, 256):
    engs = [F.Engine(p, w, h) for _ in range(S)]
    do = torch.empty((S, h, w, 2), dtype=torch.float32, device="cuda")
    for i, e in enumerate(engs):
        e.submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i].data_ptr()); e.wait()
    n = 16 * S
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        engs[i % S].submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i % S].data_ptr())
    for e in engs: e.wait()
    dt = (time.perf_counter() - t0) / n * 1e3
    print("streams %2d: %.3f ms/pair  %.0f pairs/s" % (S, dt, 1e3 / dt), flush=True)
    for e in engs: e.close()
