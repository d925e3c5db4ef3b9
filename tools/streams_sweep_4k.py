"""4K (C4a) throughput vs number of handles in flight, device-resident."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from flowonthego_b200 import api
from tests.synth import synth_pair
import torch
w, h = 3840, 2160
p = F.Params.from_argv("7 0 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0".split())
a, b, _ = synth_pair(w, h, seed=2)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
for S, grp in ((4, 0), (8, 0), (16, 0), (32, 0), (16, 8), (32, 8)):
    engs = [F.Engine(p, w, h) for _ in range(S)]
    for e in engs: e.set_option(api.OPT_SOR_GROUP, grp)
    do = torch.empty((S, h, w, 2), dtype=torch.float32, device="cuda")
    for i, e in enumerate(engs):
        e.submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i].data_ptr()); e.wait()
    n = 4 * S
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        engs[i % S].submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i % S].data_ptr())
    for e in engs: e.wait()
    dt = (time.perf_counter() - t0) / n * 1e3
    print("4K handles %2d sor_group %2d: %.3f ms/pair  %.1f pairs/s" % (S, grp, dt, 1e3 / dt), flush=True)
    for e in engs: e.close()
    del do
