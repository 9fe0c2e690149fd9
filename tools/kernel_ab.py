"""A/B of one GT kernel variant switch (development aid): per-kernel CUDA-event times of the bench's GT batch
   RMPE_RASTER_VARIANT=2 python tools/kernel_ab.py [runs=40]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import rmpe_b200  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rmpe_b200.lib.ensure_init(0)
hb = bench.make_gt_inputs(rmpe_b200, 0, bench.BATCH, bench.PERSONS)
plans = []
for _ in range(3):                      # rotate three buffer sets (> L2) like the bench
    p = rmpe_b200.batch.GtDevicePlan(bench.BATCH, bench.PERSONS, bench.SRC_HW)
    p.upload(hb["imgs"], hb["masks"], hb["joints"], hb["n_persons"], hb["M"], hb["flip"])
    plans.append(p)
for i in range(6):
    plans[i % 3].run()
torch.cuda.synchronize()
L = rmpe_b200.lib
L.profile_enable(True)
for i in range(runs):
    plans[i % 3].run()
torch.cuda.synchronize()
L.profile_enable(False, reset=False)
env = {k: v for k, v in os.environ.items() if k.startswith("RMPE_")}
print(env, {k: round(t / max(n, 1), 5) for k, (t, n) in L.profile_read().items()}, flush=True)
