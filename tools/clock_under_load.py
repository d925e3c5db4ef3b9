"""SM clock, power and utilisation while the device-resident 1080p workload runs at full throughput."""
import os, sys, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flowonthego_b200 as F
from tests.synth import synth_pair
import torch
w, h, S = 1920, 1080, 64
p = F.Params.preset(3, 1920, verbosity=0)
a, b, _ = synth_pair(w, h, seed=1)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
do = torch.empty((S, h, w, 2), dtype=torch.float32, device="cuda")
engs = [F.Engine(p, w, h) for _ in range(S)]
for i, e in enumerate(engs):
    e.submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i].data_ptr()); e.wait()
samples = []
stop = False
def sampler():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,utilization.gpu,utilization.memory,clocks_throttle_reasons.active",
                              "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.2)
th = threading.Thread(target=sampler); th.start()
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 8.0:
    for i in range(S):
        engs[i].wait()
        engs[i].submit_u8_device(da.data_ptr(), db.data_ptr(), w, h, w, do[i].data_ptr())
    n += S
for e in engs: e.wait()
dt = time.perf_counter() - t0
stop = True; th.join()
print("%.0f pairs/s over %.1f s" % (n / dt, dt))
for s in samples[::3]: print(s)
