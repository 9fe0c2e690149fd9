"""A/B of the host-buffer GT call (development aid): wall-clock of rmpe_gt_batch_host on the bench batch, pinned buffers.
   RMPE_HOST_CHUNKS=32 python tools/e2e_ab.py"""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, rmpe_b200
rmpe_b200.lib.ensure_init(0)
hb = bench.make_gt_inputs(rmpe_b200, 0, bench.BATCH, bench.PERSONS)
pin = {}
keep = []
for k, dt in (("imgs", torch.uint8), ("masks", torch.uint8), ("joints", torch.float64)):
    t, a = bench.pinned(hb[k].shape, dt); a[...] = hb[k]; keep.append(t); pin[k] = a
out = {}
for k, shape, dt in (("img", (bench.BATCH, 368, 368, 3), torch.uint8), ("mask", (bench.BATCH, 46, 46), torch.float32),
                     ("labels", (bench.BATCH, 57, 46, 46), torch.float32), ("joints", (bench.BATCH, bench.PERSONS, 18, 3), torch.float64)):
    t, a = bench.pinned(shape, dt); keep.append(t); out[k] = a
def step():
    rmpe_b200.batch.gt_batch_host(pin["imgs"], pin["masks"], pin["joints"], hb["n_persons"], hb["M"], hb["flip"], out=out)
for _ in range(3): step()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
print(os.environ.get("RMPE_HOST_CHUNKS"), "median ms %.3f  min %.3f" % (np.median(ts) * 1e3, min(ts) * 1e3))
