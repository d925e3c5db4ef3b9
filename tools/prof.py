"""Per-kernel CUDA-event profile of one pair (dev tool; run under gpurun)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flowonthego_b200 as F
from tests.synth import synth_pair

def run(label, w, h, p, npairs=3, streams=4):
    a, b, _ = synth_pair(w, h, seed=1)
    with F.Engine(p, w, h) as e:
        e.run_u8(a, b)
        t0 = time.perf_counter()
        for _ in range(5):
            e.run_u8(a, b)
        lat = (time.perf_counter() - t0) / 5 * 1e3
        e.enable_kernel_profile(True)
        for _ in range(npairs):
            e.run_u8(a, b)
        prof = e.kernel_profile()
        e.enable_kernel_profile(False)
    tot = sum(r["ms"] for r in prof) / npairs
    print("== %s  %dx%d  graph latency incl. copies %.3f ms; sum of kernels %.3f ms" % (label, w, h, lat, tot))
    byk = {}
    for r in prof:
        k = byk.setdefault(r["name"], [0.0, 0, 0.0])
        k[0] += r["ms"] / npairs; k[1] += r["launches"] // npairs; k[2] += r["alg_bytes"] / npairs
    for n, (ms, ln, by) in sorted(byk.items(), key=lambda kv: -kv[1][0]):
        print("  %-18s %8.4f ms  %4d launches  %6.1f%%  alg %8.1f MB  %7.1f GB/s" % (n, ms, ln, 100 * ms / tot, by / 1e6, by / ms / 1e6))
    if "-v" in sys.argv:
        for r in sorted(prof, key=lambda r: (r["name"], -r["level"])):
            print("     %-18s L%d %8.4f ms x%d" % (r["name"], r["level"], r["ms"] / r["launches"], r["launches"] // npairs))
    # throughput with several engines
    engs = [F.Engine(p, w, h) for _ in range(streams)]
    outs = [F.pinned_empty((h, w, 2), np.float32) for _ in range(streams)]
    import torch
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    do = torch.empty((streams, h, w, 2), dtype=torch.float32, device="cuda")
    for i, e in enumerate(engs):
        e.submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i].data_ptr()); e.wait()
    n = 8 * streams
    t0 = time.perf_counter()
    for i in range(n):
        engs[i % streams].submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i % streams].data_ptr())
    for e in engs: e.wait()
    print("  throughput with %d streams: %.3f ms/pair" % (streams, (time.perf_counter() - t0) / n * 1e3))
    for e in engs: e.close()

if __name__ == "__main__":
    which = [x for x in sys.argv[1:] if not x.startswith("-")] or ["c3", "c4"]
    if "c1" in which:
        run("C1+TV preset2", 1024, 436, F.Params.preset(2, 1024, verbosity=0), streams=8)
    if "c3" in which:
        run("C3 preset3", 1920, 1080, F.Params.preset(3, 1920, verbosity=0), streams=8)
    if "c4" in which:
        run("C4a", 3840, 2160, F.Params.from_argv("7 0 16 16 0.05 0.95 0 12 0.75 0 1 0 1 10 10 5 1 3 1.6 0".split()), streams=4)
