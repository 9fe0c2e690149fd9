"""Parity of the GPU target generator in the REFERENCE'S OWN TERMS (py_rmpe_server/rmpe_server_comparator.py:19-111):
per sample L1, L2 and exact-match fraction (AC) of the image, of the mask (x255) and of each of the 57 label layers
(x255) -- here between this package (one rmpe_gt_batch_host call) and the oracle, on seeded synthetic samples.

   python tools/parity_report.py [n_samples=256] [persons=3] [out_prefix=gpurun_out/parity_report]

Writes <prefix>.tsv (the reference's weights.tsv layout, one row per sample) and prints / writes a JSON summary."""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _oracle_one(args):
    from oracle import gt_oracle as go
    img, mask, joints, M, flip = args
    oimg, omask, oj = go.transform(img, mask, joints, M, bool(flip))
    return oimg, omask, go.create_heatmaps(oj, omask)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    persons = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    prefix = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "parity_report")
    import bench
    import rmpe_b200
    hb = bench.make_gt_inputs(rmpe_b200, 31000, n, persons)
    # oracle first (fork pool before CUDA)
    with mp.get_context("fork").Pool(bench.host_cores()) as pool:
        ora = pool.map(_oracle_one, [(hb["imgs"][i], hb["masks"][i], hb["joints"][i], hb["M"][i], hb["flip"][i]) for i in range(n)])
    cmpm = rmpe_b200.sub("py_rmpe_server.rmpe_server_comparator")
    r = rmpe_b200.batch.gt_batch_host(hb["imgs"], hb["masks"], hb["joints"], hb["n_persons"], hb["M"], hb["flip"], f64=True,
                                      chw=True)
    rows = []
    for i in range(n):
        oimg, omask, olab = ora[i]
        augs = {"b200": (r["img"][i], r["mask"][i], r["labels"][i]), "oracle": (np.transpose(oimg, (2, 0, 1)), omask, olab)}
        rows += cmpm.step(i, augs)
    rows = np.array(rows)
    cmpm.write_tsv(rows, prefix + ".tsv")
    cols = cmpm.columns()
    lay_l1 = rows[:, [cols.index("Layer%dL1" % k) for k in range(57)]]
    lay_ac = rows[:, [cols.index("Layer%dAC" % k) for k in range(57)]]
    summary = {"samples": n, "persons": persons,
               "image": {"L1_max": float(rows[:, 0].max()), "AC_min": float(rows[:, 2].min())},
               "mask_x255": {"L1_max": float(rows[:, 3].max()), "AC_min": float(rows[:, 5].min())},
               "layers_x255": {"L1_max_over_samples_and_layers": float(lay_l1.max()), "L1_mean": float(lay_l1.mean()),
                               "AC_mean": float(lay_ac.mean()), "AC_min": float(lay_ac.min()),
                               "paf_layers_AC_min": float(lay_ac[:, :38].min()), "heat_layers_AC_mean": float(lay_ac[:, 38:].mean())},
               "max_abs_label_diff": float(max(np.abs(r["labels"][i] - ora[i][2]).max() for i in range(n))),
               "note": "AC = fraction of exactly equal elements.  Image and mask are bit-exact; the labels of the GPU path hold "
                       "float32 values (returned as f64 here), so a layer's AC is the share of its cells that are exactly 0 / "
                       "exactly representable: non-zero heat and PAF values agree with the f64 oracle to float32 rounding"}
    json.dump(summary, open(prefix + ".json", "w"), indent=1)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
