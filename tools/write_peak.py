"""Write-only stream bandwidth of this GPU (development aid): torch fill of rotating buffers, CUDA events.
The label planes are a pure store stream, so this -- not the copy peak -- is what bounds k_raster."""
import json
import sys

import torch

res = {}
for mb in (123, 369, 1024):
    n = mb * 1000 * 1000 // 4
    bufs = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(3 if mb < 1000 else 2)]
    for b in bufs:
        b.fill_(1.0)
    torch.cuda.synchronize()
    best, tot = 1e9, 0.0
    iters = 30
    for i in range(iters):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        bufs[i % len(bufs)].fill_(float(i))
        e.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(e)
        best = min(best, t)
        tot += t
    res["fill_%dMB" % mb] = {"best_us": round(best * 1e3, 2), "mean_us": round(tot / iters * 1e3, 2),
                             "best_gbs": round(n * 4 / best / 1e6, 1), "mean_gbs": round(n * 4 / (tot / iters) / 1e6, 1)}
# back to back (no sync between launches), like a kernel inside a step
n = 123 * 1000 * 1000 // 4
bufs = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(3)]
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for i in range(60):
    bufs[i % 3].fill_(float(i))
e.record()
torch.cuda.synchronize()
t = a.elapsed_time(e) / 60
res["fill_123MB_back_to_back"] = {"us": round(t * 1e3, 2), "gbs": round(n * 4 / t / 1e6, 1)}
print(json.dumps(res))
