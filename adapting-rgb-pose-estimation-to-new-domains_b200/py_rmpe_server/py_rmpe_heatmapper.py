"""Drop-in for the reference's py_rmpe_server/py_rmpe_heatmapper.py (Heatmapper :8-138,
distances :144-155).  create_heatmaps() hands joints + mask to the sm_100a rasteriser
(k_raster_small / k_raster_roles, csrc/rmpe_gt.cu) and returns the (57,46,46) float64 stack;
there is no NumPy rasteriser here.  sigma and thre travel through the C ABI (RmpeGtBatchHost.sigma / .thre)."""
import numpy as np

from .. import batch as _batch
from .py_rmpe_config import RmpeGlobalConfig, TransformationParams


class Heatmapper:

    def __init__(self, sigma=TransformationParams.sigma, thre=TransformationParams.paf_thre):
        if not (float(sigma) > 0. and float(thre) > 0.):
            raise ValueError("Heatmapper needs sigma > 0 and thre > 0")
        self.sigma = float(sigma)
        self.double_sigma2 = 2 * sigma * sigma
        self.thre = thre
        stride = RmpeGlobalConfig.stride
        # kept for callers that inspect them (reference :22-25): cell centres / cell top-left
        self.grid_x = np.arange(RmpeGlobalConfig.width // stride) * stride + stride / 2 - 0.5
        self.grid_y = np.arange(RmpeGlobalConfig.height // stride) * stride + stride / 2 - 0.5
        self.Y, self.X = np.mgrid[0:RmpeGlobalConfig.height:stride, 0:RmpeGlobalConfig.width:stride]

    def create_heatmaps(self, joints, mask, return_count=False, paf_average=False):
        """(57,46,46) float64 like the reference (:32-44).  return_count adds put_vector_maps' local `count`
        (19,46,46); paf_average selects the NON-reference averaging variant the reference keeps commented out
        (:119-126)."""
        joints = np.asarray(joints, dtype=np.float64)
        P = joints.shape[0]
        res = _batch.heatmaps_host(joints.reshape(1, P, 18, 3), [P], np.asarray(mask, dtype=np.float64)[None],
                                   f64=True, want_count=return_count, sigma=self.sigma, thre=float(self.thre),
                                   paf_average=paf_average)
        if res["status"][0] & 1:
            print("Parts are too close to each other. Length is zero. Skipping")  # reference :81-84
        if return_count:
            return res["labels"][0], res["count"][0]
        return res["labels"][0]


def distances(X, Y, x1, y1, x2, y2):
    """Point-to-line distance helper of the reference (:144-155); host-side convenience only --
    the rasteriser evaluates the same un-fused f64 expression on the device."""
    xD = (x2 - x1)
    yD = (y2 - y1)
    norm2 = np.sqrt(xD ** 2 + yD ** 2)
    dist = xD * (y1 - Y) - (x1 - X) * yD
    return np.abs(dist / norm2)
