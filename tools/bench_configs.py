"""Device-side throughput of the BASELINE.json configs that are not bench.py's headline (configs[2..4]),
one rank's share of each, CUDA-event timed.  Prints one JSON line per config.  Round-1 development tool: since round 2
bench.py itself carries these workloads as the `configs.*` legs of its one JSON line (with cpu_baseline and e2e each).

  config 2: single-scale decode, ski.jpg-shaped frames, batch 64 over 8 GPUs  -> 8 frames per GPU
  config 3: multi-scale (4 scales) decode over 1k COCO2014-Val-shaped images  -> shapes drawn from
            tests/golden/val2014_1k_shapes.json, --frames per GPU (default 32)
  config 4: crowded scene, 20 persons: GT batch 64 per GPU (512 over 8) and single-scale decode of 8 frames
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmpe_b200  # noqa: E402
from bench import make_gt_inputs  # noqa: E402


def ev_time(fn, iters, warm=3):
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


def kernels(L, fn, iters=3):
    L.profile_enable(True)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    L.profile_enable(False, reset=False)
    out = {k: round(v[0] / max(v[1], 1), 5) for k, v in L.profile_read().items()}
    L.profile_enable(False, reset=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    L = rmpe_b200.lib
    L.ensure_init(0)
    S = rmpe_b200.synth

    # ---- config 2 (and the decode half of config 4) ----
    H, W = 674, 712
    h, w = S.single_scale_grid(H, W)
    for persons, tag in ((3, "config2_single_scale_ski_8frames"), (20, "config4_crowded_decode_ski_8frames")):
        frames = []
        for i in range(8):
            paf, heat, _ = S.decode_blobs(9000 + i, (H, W), (h, w), persons)
            frames.append(dict(H=H, W=W, scales=[(paf, heat, 0, 0)]))
        dp = rmpe_b200.batch.DecodeDevicePlan(frames, max_peaks=128, max_cand=2048 if persons > 8 else 1024)
        ms = ev_time(dp.run, args.iters)
        res = dp.results()
        print(json.dumps({"config": tag, "persons": persons, "frames": 8, "ms": ms, "frames_per_s": 8 / ms * 1e3,
                          "persons_found": [len(r["subset"]) for r in res], "status": [r["status"] for r in res],
                          "kernels_ms": kernels(L, dp.run)}), flush=True)

    # ---- config 3: multi-scale over COCO-val-shaped frames ----
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "val2014_1k_shapes.json")))["shapes"]
    rng = np.random.RandomState(0)
    pool = [(hh, ww) for hh, ww, c in shapes for _ in range(c)]
    pick = [pool[i] for i in rng.choice(len(pool), size=args.frames, replace=False)]
    frames = [S.multi_scale_frame(700 + i, H, W, 3) for i, (H, W) in enumerate(pick)]
    dp = rmpe_b200.batch.DecodeDevicePlan(frames)
    ms = ev_time(dp.run, max(3, args.iters // 2), warm=2)
    res = dp.results()
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dp.run()
    host_ms = (time.perf_counter() - t0) * 1e3          # time the host needs to enqueue one pass (no synchronisation)
    torch.cuda.synchronize()
    print(json.dumps({"config": "config3_multi_scale_coco_val_shapes", "frames": len(frames), "ms": ms,
                      "host_enqueue_ms": host_ms,
                      "frames_per_s": len(frames) / ms * 1e3, "shapes": sorted(set(pick))[:6],
                      "persons_found_mean": float(np.mean([len(r["subset"]) for r in res])),
                      "status_nonzero": int(sum(1 for r in res if r["status"])),
                      "kernels_ms": kernels(L, dp.run)}), flush=True)

    # ---- config 4: crowded GT ----
    B, P = 64, 20
    hb = make_gt_inputs(rmpe_b200, 5000, B, P)
    plan = rmpe_b200.batch.GtDevicePlan(B, P)
    plan.upload(hb["imgs"], hb["masks"], hb["joints"], hb["n_persons"], hb["M"], hb["flip"])
    ms = ev_time(plan.run, args.iters * 2)
    print(json.dumps({"config": "config4_crowded_gt_batch64_20persons", "ms": ms, "samples_per_s": B / ms * 1e3,
                      "kernels_ms": kernels(L, plan.run)}), flush=True)


if __name__ == "__main__":
    main()
