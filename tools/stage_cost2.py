"""Throughput-mode cost of launches and of the variational parameters (64 handles, 1080p, preset 3)."""
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
def run(name, p):
    engs = [F.Engine(p, w, h) for _ in range(S)]
    for i, e in enumerate(engs):
        e.submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i].data_ptr()); e.wait()
    n = 16 * S
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        engs[i % S].submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i % S].data_ptr())
    for e in engs: e.wait()
    dt = (time.perf_counter() - t0) / n * 1e3
    print("%-40s launches/pair %3d  %.4f ms/pair  %.0f pairs/s" % (name, engs[0].timings()["launches"], dt, 1e3 / dt), flush=True)
    for e in engs: e.close()
which = sys.argv[1] if len(sys.argv) > 1 else "params"
if which == "params":
    run("full", base)
    run("tv_solverit 1", base.copy(tv_solverit=1))
    run("tv_solverit 6", base.copy(tv_solverit=6))
    run("tv_innerit 2", base.copy(tv_innerit=2))
else:
    run("extra launches: " + os.environ.get("DIS_DEBUG_EXTRA_LAUNCHES", "0"), base)
