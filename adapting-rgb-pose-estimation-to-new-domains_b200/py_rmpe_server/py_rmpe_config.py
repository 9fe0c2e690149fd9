"""Constants of the target-generation path -- same names, values and channel layout as the
reference's py_rmpe_server/py_rmpe_config.py:12-55 (RmpeGlobalConfig, TransformationParams)
and :58-95 (RmpeCocoConfig.convert).  The CUDA side holds the same tables as __constant__ arrays in
csrc/rmpe_common.cuh (c_limb_from / c_limb_to / c_flip_partner); tests/test_host_logic.py::test_config_tables
checks the two agree.
"""
import numpy as np


class RmpeGlobalConfig:
    width = 368
    height = 368
    stride = 8

    parts = ["nose", "neck", "Rsho", "Relb", "Rwri", "Lsho", "Lelb", "Lwri", "Rhip", "Rkne",
             "Rank", "Lhip", "Lkne", "Lank", "Reye", "Leye", "Rear", "Lear"]
    num_parts = len(parts)                       # 18
    parts_dict = {name: i for i, name in enumerate(parts)}
    parts = parts + ["background"]
    num_parts_with_background = len(parts)       # 19

    # parts exchanged by a horizontal flip: L/R sho, elb, wri, hip, kne, ank, eye, ear
    leftParts = [5, 6, 7, 11, 12, 13, 15, 17]
    rightParts = [2, 3, 4, 8, 9, 10, 14, 16]

    # 0-based (from, to) part indices of the 19 limbs, PAF channels (2k, 2k+1)
    limbs_conn = [(1, 8), (8, 9), (9, 10), (1, 11), (11, 12), (12, 13), (1, 2), (2, 3), (3, 4),
                  (2, 16), (1, 5), (5, 6), (6, 7), (5, 17), (1, 0), (0, 14), (0, 15), (14, 16),
                  (15, 17)]
    limb_from = [2, 9, 10, 2, 12, 13, 2, 3, 4, 3, 2, 6, 7, 6, 2, 1, 1, 15, 16]   # 1-based
    limb_to = [9, 10, 11, 12, 13, 14, 3, 4, 5, 17, 6, 7, 8, 18, 1, 15, 16, 17, 18]

    paf_layers = 2 * len(limbs_conn)             # 38
    heat_layers = num_parts                      # 18
    num_layers = paf_layers + heat_layers + 1    # 57

    paf_start = 0
    heat_start = paf_layers                      # 38
    bkg_start = paf_layers + heat_layers         # 56

    data_shape = (3, height, width)
    mask_shape = (height // stride, width // stride)
    parts_shape = (num_layers, height // stride, width // stride)


class TransformationParams:
    target_dist = 0.6
    scale_prob = 1          # "scale improbability": 1 = scale augmentation never fires
    scale_min = 0.5
    scale_max = 1.1
    max_rotate_degree = 40.
    center_perterb_max = 40.
    flip_prob = 0.5
    sigma = 7.
    paf_thre = 8.


class RmpeCocoConfig:
    parts = ['nose', 'Leye', 'Reye', 'Lear', 'Rear', 'Lsho', 'Rsho', 'Lelb', 'Relb', 'Lwri',
             'Rwri', 'Lhip', 'Rhip', 'Lkne', 'Rkne', 'Lank', 'Rank']
    num_parts = len(parts)
    parts_dict = {name: i for i, name in enumerate(parts)}

    @staticmethod
    def convert(joints):
        """COCO-17 -> internal-18 order with a synthesised neck (mean of the shoulders when both
        are known, visibility = min).  Host-side index shuffle; reference :71-95."""
        joints = np.asarray(joints, dtype=np.float64)
        out = np.zeros((joints.shape[0], RmpeGlobalConfig.num_parts, 3), dtype=np.float64)
        out[:, :, 2] = 2.
        src = [RmpeCocoConfig.parts_dict[p] for p in RmpeCocoConfig.parts]
        dst = [RmpeGlobalConfig.parts_dict[p] for p in RmpeCocoConfig.parts]
        out[:, dst, :] = joints[:, src, :]
        neck = RmpeGlobalConfig.parts_dict['neck']
        rs = RmpeCocoConfig.parts_dict['Rsho']
        ls = RmpeCocoConfig.parts_dict['Lsho']
        both = (joints[:, ls, 2] < 2) & (joints[:, rs, 2] < 2)
        out[both, neck, 0:2] = (joints[both, rs, 0:2] + joints[both, ls, 0:2]) / 2
        out[both, neck, 2] = np.minimum(joints[both, rs, 2], joints[both, ls, 2])
        return out


def check_layer_dictionary():
    """Every one of the 57 layers is named exactly once (reference :117-133)."""
    names = [None] * RmpeGlobalConfig.paf_layers + list(RmpeGlobalConfig.parts)
    for k, (fr, to) in enumerate(RmpeGlobalConfig.limbs_conn):
        base = "%s->%s" % (RmpeGlobalConfig.parts[fr], RmpeGlobalConfig.parts[to])
        for c, axis in ((2 * k, "x"), (2 * k + 1, "y")):
            assert names[c] is None
            names[c] = base + ":" + axis
    assert all(n is not None for n in names) and len(names) == RmpeGlobalConfig.num_layers
    return names
