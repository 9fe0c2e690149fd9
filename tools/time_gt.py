"""Per-kernel device time of one GT configuration: time_gt.py [batch] [persons] [iters].  Development aid (A/B of kernel
variants through the RMPE_* environment switches), not the bench."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import rmpe_b200  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    persons = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    L = rmpe_b200.lib
    L.ensure_init(0)
    plans = []
    for k in range(3):
        hb = bench.make_gt_inputs(rmpe_b200, 1000 * k, batch, persons)
        p = rmpe_b200.batch.GtDevicePlan(batch, persons, bench.SRC_HW)
        p.upload(hb["imgs"], hb["masks"], hb["joints"], hb["n_persons"], hb["M"], hb["flip"])
        plans.append(p)
    for i in range(5):
        plans[i % 3].run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        plans[i % 3].run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    L.profile_enable(True)
    for i in range(iters):
        plans[i % 3].run()
    torch.cuda.synchronize()
    L.profile_enable(False, reset=False)
    prof = {k: round(v[0] / max(v[1], 1) * 1e3, 2) for k, v in L.profile_read().items()}
    L.profile_enable(False, reset=True)
    tag = " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("RMPE_"))
    print("GT batch %d, %d persons [%s]: step %.4f ms (%.0f samples/s); kernels us: %s" % (
        batch, persons, tag, ms, batch / ms * 1e3, prof), flush=True)


if __name__ == "__main__":
    main()
