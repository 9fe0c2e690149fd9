"""Writes tests/golden/datum_3samples.h5: three samples in the on-disk format of the reference's training set
(training/generate_hdf5_coco2014.py:313, 356-366): /datum/<7-digit key> = (6, H, W) u8 [B, G, R, legacy meta plane,
mask_miss, mask_all], attribute 'meta' = JSON with joints (1 + nop, 17, 3) in COCO order, objpos, scale_provided and the
three paths.  h5py is not in this image, so the file is written by the package's own writer (h5lite.write_datum_file);
the reader is also checked against a file written by the HDF5 library itself (tests/test_h5lite.py)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rmpe_b200  # noqa: E402,F401

h5lite = importlib.import_module("adapting-rgb-pose-estimation-to-new-domains_b200.py_rmpe_server.h5lite")


def samples():
    out = {}
    for i, (H, W, nop) in enumerate(((120, 160, 0), (144, 128, 2), (96, 200, 1))):
        rng = np.random.RandomState(100 + i)
        datum = rng.randint(0, 256, size=(6, H, W)).astype(np.uint8)
        datum[4] = 255
        datum[4, H // 4:H // 2, W // 3:W // 2] = 0          # mask_miss with a hole
        joints = []
        for p in range(1 + nop):
            j = np.zeros((17, 3))
            j[:, 0] = rng.uniform(0.1 * W, 0.9 * W, 17).round(1)
            j[:, 1] = rng.uniform(0.1 * H, 0.9 * H, 17).round(1)
            j[:, 2] = rng.choice([0.0, 1.0, 2.0], size=17, p=[0.2, 0.7, 0.1])
            joints.append(j.tolist())
        meta = {"joints": joints,
                "objpos": [[round(float(rng.uniform(0.3 * W, 0.7 * W)), 1), round(float(rng.uniform(0.3 * H, 0.7 * H)), 1)]
                           for _ in range(1 + nop)],
                "scale_provided": [round(float(rng.uniform(0.2, 0.5)), 3) for _ in range(1 + nop)],
                "img_path": "val2014/COCO_val2014_%012d.jpg" % (1000 + i),
                "mask_all_path": "mask2014/val2014_mask_all_%012d.png" % (1000 + i),
                "mask_miss_path": "mask2014/val2014_mask_miss_%012d.png" % (1000 + i)}
        out["%07d" % i] = (datum, meta)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "datum_3samples.h5")
    h5lite.write_datum_file(path, samples())
    print("wrote", path, os.path.getsize(path), "bytes")
