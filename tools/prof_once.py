"""One warm + one measured pass of each path (GT batch 256; single-scale decode of 8 ski-shaped
frames): the short command ncu is wrapped around (profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import rmpe_b200  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    persons = int(sys.argv[2]) if len(sys.argv) > 2 else bench.PERSONS
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else bench.BATCH
    rmpe_b200.lib.ensure_init(0)
    if what in ("all", "gt"):
        hb = bench.make_gt_inputs(rmpe_b200, 0, batch, persons)
        plan = rmpe_b200.batch.GtDevicePlan(batch, persons, bench.SRC_HW)
        plan.upload(hb["imgs"], hb["masks"], hb["joints"], hb["n_persons"], hb["M"], hb["flip"])
        for _ in range(2):
            plan.run()
        torch.cuda.synchronize()
    if what in ("all", "decode"):
        H, W = bench.DEC_HW
        h, w = rmpe_b200.synth.single_scale_grid(H, W)
        frames = []
        for i in range(bench.DEC_FRAMES):
            paf, heat, _ = rmpe_b200.synth.decode_blobs(9000 + i, (H, W), (h, w), persons)
            frames.append(dict(H=H, W=W, scales=[(paf, heat, 0, 0)]))
        dp = rmpe_b200.batch.DecodeDevicePlan(frames)
        for _ in range(2):
            dp.run()
        torch.cuda.synchronize()
    print("prof_once done, launches:", rmpe_b200.lib.load().rmpe_launch_count())


if __name__ == "__main__":
    main()
