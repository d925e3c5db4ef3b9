"""Throughput cost of the stages: pairs/s (64 handles, device-resident 1080p, preset 3) with parts of the
parameter set switched off.  Tells which kernels actually consume the GPU when pairs overlap."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from tests.synth import synth_pair
import torch
w, h, S = 1920, 1080, 64
a, b, _ = synth_pair(w, h, seed=1)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
do = torch.empty((S, h, w, 2), dtype=torch.float32, device="cuda")
base = F.Params.preset(3, 1920, verbosity=0)
for name, p in (("full", base), ("no variational", base.copy(usetvref=0)),
                ("1 GN iteration", base.copy(maxiter=1, miniter=1)),
                ("1 GN iteration, no variational", base.copy(maxiter=1, miniter=1, usetvref=0)),
                ("8 GN iterations", base.copy(maxiter=8, miniter=8))):
    engs = [F.Engine(p, w, h) for _ in range(S)]
    for i, e in enumerate(engs):
        e.submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i].data_ptr()); e.wait()
    n = 16 * S
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        engs[i % S].submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i % S].data_ptr())
    for e in engs: e.wait()
    dt = (time.perf_counter() - t0) / n * 1e3
    print("%-32s %.4f ms/pair  %.0f pairs/s" % (name, dt, 1e3 / dt), flush=True)
    for e in engs: e.close()
