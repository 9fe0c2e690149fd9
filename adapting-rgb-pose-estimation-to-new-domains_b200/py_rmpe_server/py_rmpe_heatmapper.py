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


    # ---- the reference's building blocks (:47-138), same in-place semantics on a caller-owned (57,46,46) stack ----
    # Each one is a call of the same rasteriser with an all-ones mask: Gaussian layers are max-merged into what the stack
    # already holds, PAF layers are overwritten only where a limb band hits (the kernel's hit count says where).
    _MAXP = 64      # persons per rasteriser call (RmpeGtBatch.max_persons)

    def _raster(self, persons, want_count=False):
        ones = np.ones((1, 46, 46))
        P = persons.shape[0]
        res = _batch.heatmaps_host(persons.reshape(1, P, 18, 3), [P], ones, f64=True, want_count=want_count,
                                   sigma=self.sigma, thre=float(self.thre))
        if res["status"][0] & 1:
            print("Parts are too close to each other. Length is zero. Skipping")  # reference :81-84
        return res["labels"][0], (res["count"][0] if want_count else None)

    def put_gaussian_maps(self, heatmaps, layer, joints):
        """heatmaps[heat_start + layer] = max(itself, Gaussians of joints (n, 2)) (reference :47-61)."""
        joints = np.asarray(joints, dtype=np.float64).reshape(-1, 2)
        plane = heatmaps[RmpeGlobalConfig.heat_start + layer]
        for i0 in range(0, joints.shape[0], self._MAXP):
            j = joints[i0:i0 + self._MAXP]
            persons = np.zeros((j.shape[0], 18, 3))
            persons[:, :, 2] = 2.0
            persons[:, layer, 0:2] = j
            persons[:, layer, 2] = 1.0
            lab, _ = self._raster(persons)
            np.maximum(plane, lab[RmpeGlobalConfig.heat_start + layer], out=plane)

    def put_joints(self, heatmaps, joints):
        """All 18 Gaussian layers of joints (P, 18, 3), visibility < 2 (reference :63-67)."""
        joints = np.asarray(joints, dtype=np.float64)
        sl = slice(RmpeGlobalConfig.heat_start, RmpeGlobalConfig.heat_start + RmpeGlobalConfig.heat_layers)
        for i0 in range(0, joints.shape[0], self._MAXP):
            persons = joints[i0:i0 + self._MAXP].copy()
            lab, _ = self._raster(persons)
            np.maximum(heatmaps[sl], lab[sl], out=heatmaps[sl])

    def put_vector_maps(self, heatmaps, layerX, layerY, joint_from, joint_to):
        """Limb bands of the pairs joint_from[i] -> joint_to[i] (n, 2) into the layers layerX / layerY: a cell inside a
        band is overwritten with that pair's unit vector, later pairs win, other cells keep their value (reference :70-127)."""
        jf = np.asarray(joint_from, dtype=np.float64).reshape(-1, 2)
        jt = np.asarray(joint_to, dtype=np.float64).reshape(-1, 2)
        fr, to = RmpeGlobalConfig.limbs_conn[0]
        ps = RmpeGlobalConfig.paf_start
        for i0 in range(0, jf.shape[0], self._MAXP):
            n = min(self._MAXP, jf.shape[0] - i0)
            persons = np.zeros((n, 18, 3))
            persons[:, :, 2] = 2.0
            persons[:, fr, 0:2] = jf[i0:i0 + n]
            persons[:, to, 0:2] = jt[i0:i0 + n]
            persons[:, fr, 2] = persons[:, to, 2] = 1.0
            lab, cnt = self._raster(persons, want_count=True)
            hit = cnt[0] > 0
            heatmaps[layerX][hit] = lab[ps][hit]
            heatmaps[layerY][hit] = lab[ps + 1][hit]

    def put_limbs(self, heatmaps, joints):
        """All 19 limbs of joints (P, 18, 3) whose two parts are visible (reference :129-138)."""
        joints = np.asarray(joints, dtype=np.float64)
        ps = RmpeGlobalConfig.paf_start
        for i0 in range(0, joints.shape[0], self._MAXP):
            lab, cnt = self._raster(joints[i0:i0 + self._MAXP].copy(), want_count=True)
            for i in range(len(RmpeGlobalConfig.limbs_conn)):
                hit = cnt[i] > 0
                heatmaps[ps + 2 * i][hit] = lab[ps + 2 * i][hit]
                heatmaps[ps + 2 * i + 1][hit] = lab[ps + 2 * i + 1][hit]


def distances(X, Y, x1, y1, x2, y2):
    """Point-to-line distance helper of the reference (:144-155); host-side convenience only --
    the rasteriser evaluates the same un-fused f64 expression on the device."""
    xD = (x2 - x1)
    yD = (y2 - y1)
    norm2 = np.sqrt(xD ** 2 + yD ** 2)
    dist = xD * (y1 - Y) - (x1 - X) * yD
    return np.abs(dist / norm2)
