"""Drop-in for py_rmpe_server/rmpe_server_comparator.py: the reference's own A/B parity harness.

It pulls the same samples from two data servers (the reference compared its Python server with an external C++
server, :14, :113-154) and reports per sample the mean absolute difference (L1), the root mean square difference (L2)
and the fraction of exactly equal elements (AC) of the image, of the 46x46 mask (differences scaled by 255) and of
each of the 57 label layers (scaled by 255), then writes `weights.tsv` with the reference's column names.  Here the
metric functions are importable (no module-level run), the picture dumps are optional and the servers are whatever
speaks the reference's wire format: this package's GPU-backed `rmpe_server.Server`, the reference's own server, or both.
tools/parity_report.py applies the same metrics to the GPU path against the oracle without any server.
"""
import os

import numpy as np

from .py_rmpe_config import RmpeGlobalConfig

servers = [('py-server', 'localhost', 5556), ('cpp-server', 'localhost', 5557)]


def _l1_l2_ac(lhs, rhs, scale):
    diff = (np.asarray(lhs, dtype=float) - np.asarray(rhs, dtype=float)) * scale
    return float(np.average(np.abs(diff))), float(np.sqrt(np.average(diff ** 2))), float(np.average(lhs == rhs))


def cmp_pics(num, lhsd, rhsd, lhsn="lhs", rhsn="rhs", save_to=None):
    """(L1, L2, AC) of two images (3, 368, 368) u8 (reference :19-35); with save_to, the two pictures and their
    absolute difference are written like the reference does."""
    res = _l1_l2_ac(lhsd, rhsd, 1.0)
    if save_to:
        import cv2
        d = os.path.join(save_to, "%5d" % num)
        os.makedirs(d, exist_ok=True)
        diff = np.abs(lhsd.astype(float) - rhsd.astype(float)).transpose([1, 2, 0]).astype(np.uint8)
        cv2.imwrite(os.path.join(d, "%07dimage.%s.png" % (num, lhsn)), lhsd.transpose([1, 2, 0]))
        cv2.imwrite(os.path.join(d, "%07dimage.%s.png" % (num, rhsn)), rhsd.transpose([1, 2, 0]))
        cv2.imwrite(os.path.join(d, "%07dimagediff.png" % num), diff)
    return res


def cmp_masks(num, lhsd, rhsd, lhsn="lhs", rhsn="rhs", save_to=None):
    """(L1, L2, AC) of two (46, 46) masks in [0, 1], differences in 1/255 units (reference :37-64)."""
    return _l1_l2_ac(lhsd, rhsd, 255.0)


def cmp_layers(num, lhsd_all, rhsd_all, lhsn="lhs", rhsn="rhs", save_to=None):
    """[L1, L2, AC] * 57 of two (57, 46, 46) label stacks, differences in 1/255 units (reference :66-111)."""
    result = []
    for layer in range(RmpeGlobalConfig.num_layers):
        result += list(_l1_l2_ac(lhsd_all[layer], rhsd_all[layer], 255.0))
    return result


def columns():
    cols = ["ImageL1", "ImageL2", "ImageAC", "MaskL1", "MaskL2", "MaskAC"]
    for layer in range(RmpeGlobalConfig.num_layers):
        cols += ["Layer%dL1" % layer, "Layer%dL2" % layer, "Layer%dAC" % layer]
    return cols


def step(num, augs, save_to=None):
    """augs: {server name: (image, mask, labels, ...)}; one result row per pair of servers (reference :113-131)."""
    all_res = []
    names = list(augs)
    for i, lhs in enumerate(names):
        for rhs in names[i + 1:]:
            res = list(cmp_pics(num, augs[lhs][0], augs[rhs][0], lhs, rhs, save_to))
            res += cmp_masks(num, augs[lhs][1], augs[rhs][1], lhs, rhs, save_to)
            res += cmp_layers(num, augs[lhs][2], augs[rhs][2], lhs, rhs, save_to)
            all_res.append(res)
    return all_res


def write_tsv(rows, path="weights.tsv"):
    cols = columns()
    with open(path, "w") as fh:
        fh.write("\t" + "\t".join(cols) + "\n")
        for i, r in enumerate(rows):
            fh.write(str(i) + "\t" + "\t".join(repr(float(v)) for v in r) + "\n")


def main(servers=servers, batch_size=20, n=2645, save_to=None, out="weights.tsv"):
    from ..training.ds_generators import DataGeneratorClient
    clients = {name: DataGeneratorClient(port=port, host=host, hwm=1, batch_size=batch_size).gen_raw()
               for (name, host, port) in servers}
    res_all = []
    for i in range(n):
        augs = {name: next(gen) for name, gen in clients.items()}
        res_all += step(i, augs, save_to)
    write_tsv(res_all, out)
    return np.array(res_all)


if __name__ == "__main__":
    np.set_printoptions(precision=1, linewidth=1000, suppress=True, threshold=100000)
    main()
