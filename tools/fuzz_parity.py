"""Large randomised parity sweep of both paths against the oracle (a one-off on the GPU box; the committed tests hold
the small fixed subset).  Prints one JSON line; profiles/ keeps the line of the end-of-round run.

   python tools/fuzz_parity.py [n_gt_batches=40] [n_decode_frames=120] [seed=0]

GT: batches of 16 random samples (source sizes 40..700 with unaligned pitches, 0..8 persons, person scales 0.12..2.5 of the
crop, rotations in and beyond the augmentation range, flips, centres near and outside the frame) through ONE
rmpe_gt_batch_host call each; warped image, 46x46 mask, joints and PAF counts must be bit-identical to the oracle, labels
within 1e-5.  Decode: random frame sizes, 1..12 persons, single and multi scale; candidate and subset arrays must be equal."""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gt_oracle_one(args):
    from oracle import gt_oracle as go
    img, mask, joints, M, flip = args
    oimg, omask, oj = go.transform(img, mask, joints, M, bool(flip))
    olab, ocnt = go.create_heatmaps(oj, omask, return_count=True)
    return oimg, omask, oj, olab, ocnt


def decode_oracle_one(case):
    from cases import decode_case_inputs
    from oracle import decode_oracle as do
    name, H, W, P, seed, multi = case
    blobs = decode_case_inputs(case)
    if multi:
        return do.multi_scale(blobs, H, W, detail=False)
    return do.single_scale(blobs[0][0], blobs[0][1], H, W, detail=False)


def main():
    n_gt = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    n_dec = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    pool = mp.get_context("fork").Pool(min(32, os.cpu_count() or 8))      # before CUDA is initialised
    import rmpe_b200
    from cases import frames_of
    rmpe_b200.lib.ensure_init(0)
    rng = np.random.RandomState(seed)
    t0 = time.time()
    out = {"gt_samples": 0, "gt_img_bad": 0, "gt_mask_bad": 0, "gt_joint_bad": 0, "gt_count_bad": 0,
           "gt_label_maxabs": 0.0, "gt_status_nonzero": 0, "decode_frames": 0, "decode_bad": 0, "decode_status_nonzero": 0}

    B = 16
    for _ in range(n_gt):
        H, W = int(rng.randint(40, 700)), int(rng.randint(40, 700))
        if rng.rand() < 0.3:
            H, W = 8 * int(rng.randint(8, 80)), 8 * int(rng.randint(8, 80))     # 8-byte aligned pitches (64-bit staging)
        P = int(rng.randint(0, 9))
        imgs = rng.randint(0, 256, size=(B, H, W, 3)).astype(np.uint8)
        masks = (rng.randint(0, 5, size=(B, H, W)) > 0).astype(np.uint8) * 255
        if rng.rand() < 0.3:
            masks = rng.randint(0, 256, size=(B, H, W)).astype(np.uint8)
        joints = np.zeros((B, max(P, 1), 18, 3))
        joints[..., 2] = 2.0
        if P:
            joints[:, :P, :, 0] = rng.uniform(-0.1 * W, 1.1 * W, (B, P, 18))
            joints[:, :P, :, 1] = rng.uniform(-0.1 * H, 1.1 * H, (B, P, 18))
            joints[:, :P, :, 2] = rng.choice([0.0, 1.0, 2.0], (B, P, 18), p=[0.2, 0.6, 0.2])
            if rng.rand() < 0.3:
                joints[:, :P, :, :2] = np.round(joints[:, :P, :, :2])          # tie-prone integer joints
        n_persons = np.full(B, P, np.int32)
        flip = rng.randint(0, 2, B).astype(np.uint8)
        deg = np.where(rng.rand(B) < 0.2, rng.uniform(-180, 180, B), rng.uniform(-40, 40, B))
        scale_self = np.exp(rng.uniform(np.log(0.12), np.log(2.5), B))
        centers = np.stack([rng.uniform(-0.2, 1.2, B) * W, rng.uniform(-0.2, 1.2, B) * H], 1)
        crops = [(int(rng.randint(-40, 41)), int(rng.randint(-40, 41))) for _ in range(B)]
        M = rmpe_b200.batch.aug_affine(flip, list(deg), crops, [1.0] * B, centers, scale_self)
        r = rmpe_b200.batch.gt_batch_host(imgs, masks, joints, n_persons, M, flip, f64=False, want_count=True)
        ora = pool.map(gt_oracle_one, [(imgs[i], masks[i], joints[i, :P], M[i], flip[i]) for i in range(B)])
        for i, (oimg, omask, oj, olab, ocnt) in enumerate(ora):
            out["gt_samples"] += 1
            out["gt_img_bad"] += int(not np.array_equal(r["img"][i], oimg))
            out["gt_mask_bad"] += int(not np.array_equal(r["mask"][i], omask.astype(np.float32)))
            out["gt_joint_bad"] += int(P > 0 and not np.array_equal(r["joints"][i, :P], oj))
            out["gt_count_bad"] += int(not np.array_equal(r["count"][i], ocnt))
            out["gt_label_maxabs"] = max(out["gt_label_maxabs"], float(np.abs(r["labels"][i] - olab).max()))
            out["gt_status_nonzero"] += int(r["status"][i] & ~1 != 0)           # bit 0 = zero-length limb (the reference prints)
    out["gt_seconds"] = round(time.time() - t0, 1)

    t1 = time.time()
    cases = []
    for k in range(n_dec):
        H, W = int(rng.randint(40, 520)), int(rng.randint(40, 720))
        P = int(rng.randint(1, 13))
        cases.append(("f%d" % k, H, W, P, 20000 + 7 * k + seed, bool(k % 3 == 0)))
    ora = pool.map(decode_oracle_one, cases, chunksize=1)
    for c0 in range(0, n_dec, 8):
        chunk = cases[c0:c0 + 8]
        res = rmpe_b200.batch.decode_batch_host([frames_of(c) for c in chunk])
        for c, rr, o in zip(chunk, res, ora[c0:c0 + 8]):
            out["decode_frames"] += 1
            out["decode_status_nonzero"] += int(rr["status"] != 0)
            ok = np.array_equal(rr["candidate"], o[0]) and np.array_equal(rr["subset"], o[1])
            out["decode_bad"] += int(not ok)
            if not ok:
                print("decode mismatch:", c, file=sys.stderr)
    out["decode_seconds"] = round(time.time() - t1, 1)
    pool.close()
    print(json.dumps(out))
    bad = out["gt_img_bad"] + out["gt_mask_bad"] + out["gt_joint_bad"] + out["gt_count_bad"] + out["decode_bad"]
    sys.exit(1 if bad or out["gt_label_maxabs"] > 1e-5 else 0)


if __name__ == "__main__":
    main()
