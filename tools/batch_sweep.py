"""Throughput with batched handles (dis_create_batch): S handles x nb pairs per launch; 1080p preset 3."""
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
combos = [(64, 1), (32, 2), (16, 4), (32, 4), (64, 4), (8, 8), (16, 8), (32, 8)]
if len(sys.argv) > 1:
    combos = [tuple(map(int, c.split("x"))) for c in sys.argv[1:]]
for S, nb in combos:
    engs = [F.Engine(p, w, h, batch=nb) for _ in range(S)]
    do = torch.zeros((S, nb, h, w, 2), dtype=torch.float32, device="cuda")
    A, B = [da.data_ptr()] * nb, [db.data_ptr()] * nb
    outs = [[do[s, i].data_ptr() for i in range(nb)] for s in range(S)]
    for s, e in enumerate(engs):
        e.submit_u8_device_batch(A, B, w, h, w, outs[s]); e.wait()
    ok = bool((do == ref).all().item())
    reps = max(2, 2048 // (S * nb))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for r in range(reps):
        for s, e in enumerate(engs):
            e.submit_u8_device_batch(A, B, w, h, w, outs[s])
    for e in engs: e.wait()
    dt = (time.perf_counter() - t0) / (reps * S * nb) * 1e3
    print("handles %2d x batch %d: %.4f ms/pair  %.0f pairs/s  launches/call %d  bit-exact=%s"
          % (S, nb, dt, 1e3 / dt, engs[0].timings()["launches"], ok), flush=True)
    for e in engs: e.close()
    del do
