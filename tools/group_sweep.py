"""Throughput with n pairs per graph launch (dis_group_*), G groups in flight; 1080p preset 3, device-resident."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from tests.synth import synth_pair
import torch
w, h = 1920, 1080
p = F.Params.preset(3, 1920, verbosity=0)
a, b, _ = synth_pair(w, h, seed=1)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
with F.Engine(p, w, h) as e:
    ref = torch.from_numpy(e.run_u8(a, b)).cuda()
combos = [(32, 1), (32, 2), (32, 4), (16, 4), (16, 8), (8, 8), (32, 8), (8, 16), (4, 32)]
if len(sys.argv) > 1:
    combos = [tuple(map(int, c.split("x"))) for c in sys.argv[1:]]
for G, n in combos:
    groups = [F.EngineGroup(p, w, h, n) for _ in range(G)]
    do = torch.zeros((G, n, h, w, 2), dtype=torch.float32, device="cuda")
    A, B = [da.data_ptr()] * n, [db.data_ptr()] * n
    outs = [[do[g, i].data_ptr() for i in range(n)] for g in range(G)]
    for g, gr in enumerate(groups):
        gr.submit_u8_device(A, B, w, h, w, outs[g]); gr.wait()
    ok = bool((do == ref).all().item())
    reps = max(2, 1024 // (G * n))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for r in range(reps):
        for g, gr in enumerate(groups):
            gr.submit_u8_device(A, B, w, h, w, outs[g])
    for gr in groups: gr.wait()
    dt = (time.perf_counter() - t0) / (reps * G * n) * 1e3
    print("groups %2d x %2d pairs: %.4f ms/pair  %.0f pairs/s  bit-exact=%s" % (G, n, dt, 1e3 / dt, ok), flush=True)
    for gr in groups: gr.close()
    del do
