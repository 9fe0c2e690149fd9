"""One warm + one measured pass of the multi-scale decode over 64 COCO-val-shaped frames (one chunk): the short command
ncu is wrapped around for the multi-scale variants of k_screen_pairs (profiles/README.md)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rmpe_b200  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rmpe_b200.lib.ensure_init(0)
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "val2014_1k_shapes.json")))["shapes"]
    rng = np.random.RandomState(0)
    pool = [(hh, ww) for hh, ww, c in shapes for _ in range(c)]
    pick = [pool[i] for i in rng.choice(len(pool), size=n, replace=False)]
    frames = [rmpe_b200.synth.multi_scale_frame(700 + i, H, W, 3) for i, (H, W) in enumerate(pick)]
    dp = rmpe_b200.batch.DecodeDevicePlan(frames)
    for _ in range(2):
        dp.run()
    torch.cuda.synchronize()
    res = dp.results()
    print("prof_ms done:", n, "frames,", sum(len(r["subset"]) for r in res), "persons")


if __name__ == "__main__":
    main()
