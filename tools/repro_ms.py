"""Multi-scale decode over a handful of COCO-val-shaped frames (memcheck target)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmpe_b200
from tools.bench_configs import multi_scale_feed_shapes
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
one = len(sys.argv) > 2
S = rmpe_b200.synth
rmpe_b200.lib.ensure_init(0)
shapes = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "val2014_1k_shapes.json")))["shapes"]
rng = np.random.RandomState(0)
pool = [(hh, ww) for hh, ww, c in shapes for _ in range(c)]
pick = [pool[i] for i in rng.choice(len(pool), size=32, replace=False)][:n]
frames = []
for i, (H, W) in enumerate(pick):
    _, _, persons = S.decode_blobs(700 + i, (H, W), (4, 4), 3)
    sc = []
    for (Hs, pd, pr, hs, ws) in multi_scale_feed_shapes(H, W):
        paf, heat, _ = S.decode_blobs(700 + i + 1000 * len(sc), (H, W), (hs, ws), 3, persons=persons, stride=8.0 * H / Hs)
        sc.append((paf, heat, pd, pr))
    frames.append(dict(H=H, W=W, scales=sc))
print(pick, flush=True)
if one:
    for i, f in enumerate(frames):
        dp = rmpe_b200.batch.DecodeDevicePlan([f])
        dp.run(); torch.cuda.synchronize()
        print("frame", i, pick[i], "ok", dp.results()[0]["status"], flush=True)
else:
    dp = rmpe_b200.batch.DecodeDevicePlan(frames)
    dp.run(); torch.cuda.synchronize()
    print("ok", [r["status"] for r in dp.results()])
import time
if not one:
    for _ in range(3):
        dp.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        dp.run()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("host submit per call %.3f ms, total per call %.3f ms" % ((t1 - t0) / 20 * 1e3, (t2 - t0) / 20 * 1e3))
