"""Where does the host-buffer path (dis_submit_u8 / dis_wait) spend its time?  Prints pairs/s and the
in-stream phase times (CUDA events) for several handle counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flowonthego_b200 as F
from tests.synth import synth_pair
W, H, B = 1920, 1080, 128
p = F.Params.preset(3, W, verbosity=0)
a, b, _ = synth_pair(W, H, seed=1)
for S in (4, 8, 16, 32, 64):
    eng = [F.Engine(p, W, H) for _ in range(S)]
    ha = [F.pinned_empty((H, W), np.uint8) for _ in range(S)]
    hb = [F.pinned_empty((H, W), np.uint8) for _ in range(S)]
    ho = [F.pinned_empty((H, W, 2), np.float32) for _ in range(S)]
    for i in range(S):
        ha[i][...] = a; hb[i][...] = b
    def step():
        tsub = 0.0
        for i in range(B):
            e = eng[i % S]
            if i >= S:
                e.wait()
            t = time.perf_counter()
            e.submit_u8(ha[i % S], hb[i % S], ho[i % S])
            tsub += time.perf_counter() - t
        for e in eng:
            e.wait()
        return tsub
    step()
    t = time.perf_counter(); tsub = step() + step(); dt = (time.perf_counter() - t) / 2
    tm = eng[0].timings()
    print("S=%2d  %.0f pairs/s  %.3f ms/pair  host submit %.3f ms/pair | last pair in-stream: h2d %.2f  d2h %.2f  total %.2f ms"
          % (S, B / dt, dt / B * 1e3, tsub / 2 / B * 1e3, tm["h2d_ms"], tm["d2h_ms"], tm["total_ms"]))
    for e in eng:
        e.close()
