"""Throughput cost of the stages with batched handles (32 handles x 8 pairs per launch), 1080p preset 3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from flowonthego_b200 import api
from tests.synth import synth_pair
import torch
w, h, S, nb = 1920, 1080, 32, 8
a, b, _ = synth_pair(w, h, seed=1)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
do = torch.empty((S, nb, h, w, 2), dtype=torch.float32, device="cuda")
A, B = [da.data_ptr()] * nb, [db.data_ptr()] * nb
outs = [[do[s, i].data_ptr() for i in range(nb)] for s in range(S)]
base = F.Params.preset(3, 1920, verbosity=0)
cases = [("full", base, 0), ("full, SOR kG=16", base, 16), ("no variational", base.copy(usetvref=0), 0),
         ("1 GN iteration", base.copy(maxiter=1, miniter=1), 0),
         ("1 GN iteration, no variational", base.copy(maxiter=1, miniter=1, usetvref=0), 0),
         ("tv_solverit 1", base.copy(tv_solverit=1), 0), ("tv_solverit 6", base.copy(tv_solverit=6), 0)]
for name, p, grp in cases:
    engs = [F.Engine(p, w, h, batch=nb) for _ in range(S)]
    for s, e in enumerate(engs):
        e.set_option(api.OPT_SOR_GROUP, grp)
        e.submit_u8_device_batch(A, B, w, h, w, outs[s]); e.wait()
    reps = 8
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for r in range(reps):
        for s, e in enumerate(engs):
            e.submit_u8_device_batch(A, B, w, h, w, outs[s])
    for e in engs: e.wait()
    dt = (time.perf_counter() - t0) / (reps * S * nb) * 1e3
    print("%-32s %.4f ms/pair  %.0f pairs/s" % (name, dt, 1e3 / dt), flush=True)
    for e in engs: e.close()
