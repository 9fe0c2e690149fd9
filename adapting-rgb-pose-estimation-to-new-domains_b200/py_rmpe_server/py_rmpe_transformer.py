"""Drop-in for the reference's py_rmpe_server/py_rmpe_transformer.py (AugmentSelection :10-78,
Transformer.transform :83-114).  Same names, arguments, return values and in-place mutation of
meta['joints']; the arithmetic runs in the sm_100a kernels behind include/rmpe_b200.h
(k_warp_fused: image + mask warp and the 368->46 mask reduction; k_raster_small / k_raster_roles: keypoint
transform) -- nothing here computes pixels on the CPU."""
import random

import numpy as np

from .. import batch as _batch
from .py_rmpe_config import RmpeGlobalConfig, TransformationParams


class AugmentSelection:

    def __init__(self, flip=False, degree=0., crop=(0, 0), scale=1.):
        self.flip = flip
        self.degree = degree  # rotate
        self.crop = crop      # shift actually
        self.scale = scale

    @staticmethod
    def random():
        # same draw order from python's `random` as the reference (:19-27)
        flip = random.uniform(0., 1.) > TransformationParams.flip_prob
        degree = random.uniform(-1., 1.) * TransformationParams.max_rotate_degree
        scale = (TransformationParams.scale_max - TransformationParams.scale_min) * random.uniform(0., 1.) \
            + TransformationParams.scale_min \
            if random.uniform(0., 1.) > TransformationParams.scale_prob else 1.
        x_offset = int(random.uniform(-1., 1.) * TransformationParams.center_perterb_max)
        y_offset = int(random.uniform(-1., 1.) * TransformationParams.center_perterb_max)
        return AugmentSelection(flip, degree, (x_offset, y_offset), scale)

    @staticmethod
    def unrandom():
        return AugmentSelection(False, 0., (0, 0), 1.)

    def affine(self, center, scale_self):
        """(2,3) float64 forward matrix: translate -> rotate -> scale -> flip -> recentre."""
        M = _batch.aug_affine([1 if self.flip else 0], [self.degree], [self.crop], [self.scale],
                              [center], [scale_self])
        return M[0]


class Transformer:

    @staticmethod
    def transform(img, mask, meta, aug=None):
        """Returns (img (368,368,3) u8, mask (46,46) f64 in [0,1], meta); meta['joints'] is
        overwritten in place with the transformed (and, on flip, left/right swapped) joints."""
        if aug is None:
            aug = AugmentSelection.random()
        M = aug.affine(meta['objpos'][0], meta['scale_provided'][0])
        joints = np.asarray(meta['joints'], dtype=np.float64)
        P = joints.shape[0]
        res = _batch.gt_batch_host(np.asarray(img)[None], np.asarray(mask)[None], joints[None],
                                   [P], M[None], [1 if aug.flip else 0], f64=True, want_labels=False)
        if P:
            meta['joints'][:, :, :] = res["joints"][0]
        return res["img"][0], res["mask"][0], meta
