"""Quick device-side timings (CUDA events) of the two paths; development aid, not the bench."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmpe_b200  # noqa: E402


def ev_time(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    rmpe_b200.lib.ensure_init(0)
    B, P = 256, 3
    t0 = time.time()
    b = rmpe_b200.synth.gt_batch(B, n_persons=P, seed0=0)
    flip = np.array([a[0] for a in b["augs"]], np.uint8)
    M = rmpe_b200.batch.aug_affine(flip, [a[1] for a in b["augs"]], [a[2] for a in b["augs"]],
                                   [a[3] for a in b["augs"]], b["centers"], b["scale_self"])
    print("synth %.1fs" % (time.time() - t0), flush=True)
    plan = rmpe_b200.batch.GtDevicePlan(B, P)
    plan.upload(b["imgs"], b["masks"], b["joints"], b["n_persons"], M, flip)
    for simple in (True, False):
        ms = ev_time(lambda: plan.run(simple=simple))
        print("GT batch %d %s: %.3f ms  -> %.0f samples/s" % (B, "simple" if simple else "tile", ms, B / ms * 1e3), flush=True)
    L = rmpe_b200.lib
    L.profile_enable(True)
    for _ in range(5):
        plan.run()
    torch.cuda.synchronize()
    L.profile_enable(False, reset=False)
    for k, (tms, n) in L.profile_read().items():
        print("   %-20s %.4f ms/launch" % (k, tms / max(n, 1)))
    L.profile_enable(False, reset=True)
    # per-kernel split via flags
    d = plan.desc_struct
    full_flags = d.flags
    d.flags = full_flags | L.GT_NO_WARP
    ms = ev_time(lambda: plan.run())
    print("  mask46 + raster only: %.3f ms" % ms, flush=True)
    d.flags = full_flags
    for P2 in (20,):
        b2 = rmpe_b200.synth.gt_batch(64, n_persons=P2, seed0=5000)
        flip2 = np.array([a[0] for a in b2["augs"]], np.uint8)
        M2 = rmpe_b200.batch.aug_affine(flip2, [a[1] for a in b2["augs"]], [a[2] for a in b2["augs"]],
                                        [a[3] for a in b2["augs"]], b2["centers"], b2["scale_self"])
        plan2 = rmpe_b200.batch.GtDevicePlan(64, P2)
        plan2.upload(b2["imgs"], b2["masks"], b2["joints"], b2["n_persons"], M2, flip2)
        ms = ev_time(lambda: plan2.run())
        print("GT batch 64, %d persons: %.3f ms -> %.0f samples/s" % (P2, ms, 64 / ms * 1e3), flush=True)

    H, W = 674, 712
    h, w = rmpe_b200.synth.single_scale_grid(H, W)
    for P3, nf in ((3, 8), (20, 8)):
        frames = []
        for i in range(nf):
            paf, heat, _ = rmpe_b200.synth.decode_blobs(900 + i, (H, W), (h, w), P3)
            frames.append(dict(H=H, W=W, scales=[(paf, heat, 0, 0)]))
        dp = rmpe_b200.batch.DecodeDevicePlan(frames)
        ms = ev_time(dp.run, iters=5, warm=2)
        print("decode single-scale ski x%d, %d persons: %.3f ms -> %.1f frames/s" % (nf, P3, ms, nf / ms * 1e3), flush=True)
    from oracle import decode_oracle as do
    H, W = 480, 640
    frames = []
    for i in range(4):
        shapes = do.multi_scale_feed_shapes(H, W)
        _, _, persons = rmpe_b200.synth.decode_blobs(700 + i, (H, W), (4, 4), 3)
        sc = []
        for (Hs, Ws, pd, pr, hs, ws) in shapes:
            paf, heat, _ = rmpe_b200.synth.decode_blobs(700 + i + 1000 * len(sc), (H, W), (hs, ws), 3, persons=persons,
                                                        stride=8.0 * H / Hs)
            sc.append((paf, heat, pd, pr))
        frames.append(dict(H=H, W=W, scales=sc))
    dp = rmpe_b200.batch.DecodeDevicePlan(frames)
    ms = ev_time(dp.run, iters=3, warm=1)
    print("decode multi-scale 480x640 x4: %.3f ms -> %.1f frames/s" % (ms, 4 / ms * 1e3), flush=True)
    L.profile_enable(True)
    for _ in range(3):
        dp.run()
    torch.cuda.synchronize()
    L.profile_enable(False, reset=False)
    for k, (tms, n) in L.profile_read().items():
        print("   %-20s %.3f ms/launch x %d" % (k, tms / max(n, 1), n // 3))
    L.profile_enable(False, reset=True)


if __name__ == "__main__":
    main()
