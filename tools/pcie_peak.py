"""Host<->device copy rates of this box for the e2e step's byte counts (development aid): pinned buffers,
one direction at a time and both at once on two streams; CUDA events + wall clock."""
import json
import time

import torch

H2D, D2H = 139_027_712, 230_011_904          # bench.py e2e bytes per 256-sample step
hin = torch.empty(H2D, dtype=torch.uint8).pin_memory()
hout = torch.empty(D2H, dtype=torch.uint8).pin_memory()
din = torch.empty(H2D, dtype=torch.uint8, device="cuda")
dout = torch.empty(D2H, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def h2d():
    with torch.cuda.stream(s1):
        din.copy_(hin, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hout.copy_(dout, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn, nbytes in (("h2d", h2d, H2D), ("d2h", d2h, D2H), ("both", both, H2D + D2H)):
    ms = timed(fn)
    res[name] = {"ms": round(ms, 3), "gbs": round(nbytes / ms / 1e6, 1)}
print(json.dumps(res))
