"""Tiny end-to-end pass of both paths for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_once.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmpe_b200  # noqa: E402


def main():
    rmpe_b200.lib.ensure_init(0)
    for hw, P in (((368, 368), 3), ((240, 322), 2), ((96, 128), 6)):     # aligned, unaligned pitch, > 4 persons
        b = rmpe_b200.synth.gt_batch(3, n_persons=P, seed0=40, src_hw=hw)
        flip = np.array([a[0] for a in b["augs"]], np.uint8)
        ss = b["scale_self"] * np.array([0.5, 1.0, 2.2])
        M = rmpe_b200.batch.aug_affine(flip, [a[1] for a in b["augs"]], [a[2] for a in b["augs"]], [a[3] for a in b["augs"]],
                                       b["centers"], ss)
        r = rmpe_b200.batch.gt_batch_host(b["imgs"], b["masks"], b["joints"], b["n_persons"], M, flip, want_count=True)
        assert r["status"].max() < 256
    H, W = 160, 208
    h, w = rmpe_b200.synth.single_scale_grid(H, W)
    paf, heat, _ = rmpe_b200.synth.decode_blobs(3, (H, W), (h, w), 4)
    res = rmpe_b200.batch.decode_batch_host([dict(H=H, W=W, scales=[(paf, heat, 0, 0)])] * 2)
    assert res[0]["status"] == 0
    # multi scale (4 scales, eval_method 1): two frames of different shapes
    frames = [rmpe_b200.synth.multi_scale_frame(50 + i, H, W, 2) for i, (H, W) in enumerate(((120, 160), (97, 141)))]
    res2 = rmpe_b200.batch.decode_batch_host(frames)
    assert all(r["status"] == 0 for r in res2)
    print("sanitize_once done:", len(res[0]["subset"]), "persons")


if __name__ == "__main__":
    main()
