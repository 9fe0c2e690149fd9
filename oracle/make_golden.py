"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_shim.py) on the seeded synthetic inputs of
`synth.py`.  Run in the build container only:  python oracle/make_golden.py

The reference ships no golden vectors of its own, so these files are the
pin: inputs are regenerated from seeds at test time (their sha256 is stored to detect
generator drift); outputs are stored in full where small, and as sha256 of the exact bytes
where large (warped images, f64 label stacks), plus full float32 label stacks for 3 cases.
"""
import hashlib
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "adapting-rgb-pose-estimation-to-new-domains_b200"
synth = importlib.import_module(PKG + ".synth")
from oracle import ref_shim  # noqa: E402
from oracle import decode_oracle as do  # noqa: E402

GT_CASES = [
    # name, seed, persons, src_hw, augment, integer_joints, aug override
    ("g0", 0, 3, (368, 368), True, False, None),
    ("g1", 1, 3, (368, 368), True, False, None),
    ("g2", 2, 3, (368, 368), True, False, None),
    ("g3", 3, 3, (368, 368), True, False, None),
    ("g4", 4, 3, (368, 368), True, False, None),
    ("g5", 5, 3, (368, 368), True, False, None),
    ("g6", 6, 20, (368, 368), True, False, None),
    ("g7", 7, 3, (300, 420), True, False, None),
    ("g8", 8, 3, (368, 368), False, False, None),                       # C1: unrandom
    ("g9", 9, 4, (368, 368), False, True, (True, 0.0, (8, -16), 1.0)),  # tie-prone integer joints
    ("g10", 10, 0, (368, 368), True, False, None),                      # no persons
    ("g11", 11, 5, (480, 640), True, False, (False, 33.0, (-30, 21), 0.7)),
]

DECODE_CASES = [
    # name, H, W, persons, seed, multi
    ("d0", 674, 712, 3, 1, False),
    ("d1", 674, 712, 20, 2, False),
    ("d2", 240, 320, 2, 3, False),
    ("d3", 333, 251, 6, 4, False),
    ("d4", 480, 640, 3, 5, True),
    ("d5", 427, 640, 8, 6, True),
    ("d6", 120, 160, 0, 7, False),
]

FULL_LABEL_CASES = ("g0", "g6", "g9")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gt_case_inputs(case):
    name, seed, P, hw, augment, integer, aug_override = case
    s = synth.gt_sample(seed, P, hw, augment, integer)
    if aug_override is not None:
        s["aug"] = aug_override
    return s


def decode_case_inputs(case):
    name, H, W, P, seed, multi = case
    if not multi:
        h, w = synth.single_scale_grid(H, W)
        paf, heat, _ = synth.decode_blobs(seed, (H, W), (h, w), P)
        return [(paf, heat, 0, 0)]
    shapes = do.multi_scale_feed_shapes(H, W)
    _, _, persons = synth.decode_blobs(seed, (H, W), (4, 4), P)
    blobs = []
    for (Hs, Ws, pd, pr, hs, ws) in shapes:
        paf, heat, _ = synth.decode_blobs(seed + 1000 * len(blobs), (H, W), (hs, ws), P,
                                          persons=persons, stride=8.0 * H / Hs)
        blobs.append((paf, heat, pd, pr))
    return blobs


def main():
    import cv2
    ref = ref_shim.load()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    gt = {}
    for case in GT_CASES:
        name = case[0]
        s = gt_case_inputs(case)
        flip, deg, crop, scale = s["aug"]
        aug = ref.AugmentSelection(flip, deg, crop, scale)
        M = aug.affine(s["objpos"][0], s["scale_provided"][0])
        meta = dict(objpos=s["objpos"], scale_provided=s["scale_provided"], joints=s["joints"].copy())
        img, mask, meta = ref.Transformer.transform(s["img"], s["mask"], meta, aug)
        labels = ref.Heatmapper().create_heatmaps(meta["joints"], mask)
        gt[name + "_in_sha"] = np.array(sha(s["img"]) + sha(s["mask"]) + sha(s["joints"]))
        gt[name + "_M"] = M
        gt[name + "_img_sha"] = np.array(sha(img))
        gt[name + "_img_rows"] = img[::23].copy()          # 16 full rows for a readable diff
        gt[name + "_mask46"] = np.rint(mask * 255.).astype(np.uint8)
        gt[name + "_mask_sha"] = np.array(sha(mask))
        gt[name + "_joints"] = meta["joints"]
        gt[name + "_labels_sha"] = np.array(sha(labels))
        gt[name + "_labels_sum"] = labels.sum(axis=(1, 2))
        if name in FULL_LABEL_CASES:
            gt[name + "_labels_f32"] = labels.astype(np.float32)
        print(name, "done", img.shape, labels.shape)
    np.savez_compressed(os.path.join(out_dir, "gt_golden.npz"), **gt)

    dec = {}
    params = {'scale_search': [.5, 1, 1.5, 2], 'thre1': .1, 'thre2': .05}
    mparams = {'boxsize': 368, 'stride': 8, 'padValue': 128}
    tmp = tempfile.mkdtemp()
    for case in DECODE_CASES:
        name, H, W, P, seed, multi = case
        blobs = decode_case_inputs(case)
        path = os.path.join(tmp, name + ".png")
        cv2.imwrite(path, np.zeros((H, W, 3), np.uint8))
        fm = ref_shim.FakeModel(lambda hh, ww, i: (blobs[i][0], blobs[i][1]))
        fn = ref.eval.process_multi_scale if multi else ref.eval.process_single_scale
        canvas, candidate, subset = fn(path, fm, dict(params), mparams)
        candidate = np.asarray(candidate, dtype=np.float64).reshape(-1, 4)
        dec[name + "_in_sha"] = np.array("".join(sha(b[0]) + sha(b[1]) for b in blobs))
        dec[name + "_candidate"] = candidate
        dec[name + "_subset"] = np.asarray(subset, dtype=np.float64).reshape(-1, 20)
        print(name, "done", candidate.shape, subset.shape)
    np.savez_compressed(os.path.join(out_dir, "decode_golden.npz"), **dec)


if __name__ == "__main__":
    main()
