"""Host<->device copy bandwidth of the box (pinned memory), to put the bench's e2e number in context."""
import torch, time
dev = torch.device("cuda", 0)
def bw(nbytes, direction, streams=1, reps=20):
    hs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(streams)]
    ds = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    def go():
        for h, d, s in zip(hs, ds, ss):
            with torch.cuda.stream(s):
                if direction in ("h2d", "both"):
                    d.copy_(h, non_blocking=True)
                if direction in ("d2h", "both"):
                    h.copy_(d, non_blocking=True)
    go(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        go()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    mult = 2 if direction == "both" else 1
    return nbytes * streams * reps * mult / dt / 1e9
for n in (2 << 20, 16 << 20, 256 << 20):
    for d in ("h2d", "d2h", "both"):
        print(n >> 20, "MiB", d, "1 stream %.1f GB/s" % bw(n, d, 1), " 4 streams %.1f GB/s" % bw(n, d, 4))
