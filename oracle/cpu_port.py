"""ORACLE-SIDE CPU BASELINE (test infrastructure / bench.py cpu_baseline + --impl reference only).

The reference's own CPU path restated with the same third-party calls it makes -- cv2.warpAffine,
cv2.resize (py_rmpe_transformer.py:90-92), np.matmul (:102), the NumPy rasteriser of
py_rmpe_heatmapper.py:32-138, cv2.resize + scipy gaussian_filter + the pure-Python limb loops of
eval/eval_coco2014_multi_modes.py:263-415 -- so that its speed is the reference's speed on this
host.  It is pinned against the golden vectors in tests/test_cpu_port.py.  The product never
imports this file."""
import numpy as np

from . import gt_oracle as go
from . import decode_oracle as do


def gt_sample(img, mask, joints, M, flip):
    """Transformer.transform + Heatmapper.create_heatmaps for one sample, on cv2 + numpy."""
    import cv2
    M = np.asarray(M, dtype=np.float64)
    out = cv2.warpAffine(img, M, (go.HEIGHT, go.WIDTH), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT,
                         borderValue=(127, 127, 127))
    m = cv2.warpAffine(mask, M, (go.HEIGHT, go.WIDTH), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT,
                       borderValue=255)
    m = cv2.resize(m, (go.GRID, go.GRID), interpolation=cv2.INTER_CUBIC).astype(np.float64) / 255.
    j = np.array(joints, dtype=np.float64, copy=True)
    pts = j.copy()
    pts[:, :, 2] = 1
    j[:, :, 0:2] = np.matmul(M, pts.transpose([0, 2, 1])).transpose([0, 2, 1])
    if flip:
        left = j[:, go.LEFT_PARTS, :].copy()
        right = j[:, go.RIGHT_PARTS, :].copy()
        j[:, go.LEFT_PARTS, :] = right
        j[:, go.RIGHT_PARTS, :] = left
    labels = go.create_heatmaps(j, m)
    return out, m, j, labels


def decode_frame(scales, H, W, thre1=0.1, thre2=0.05):
    """process_single_scale / process_multi_scale from the blobs on, with cv2.resize and
    scipy.ndimage.gaussian_filter doing what they do in the reference."""
    import cv2
    from scipy.ndimage import gaussian_filter
    if len(scales) == 1:
        paf, heat = scales[0][0], scales[0][1]
        heat_up = cv2.resize(heat, (W, H), interpolation=cv2.INTER_CUBIC)
        paf_up = cv2.resize(paf, (W, H), interpolation=cv2.INTER_CUBIC)
    else:
        heat_up = np.zeros((H, W, 19))
        paf_up = np.zeros((H, W, 38))
        for paf, heat, pd, pr in scales:
            hs, ws = heat.shape[:2]
            hh = cv2.resize(heat, (0, 0), fx=8, fy=8, interpolation=cv2.INTER_CUBIC)[:hs * 8 - pd, :ws * 8 - pr, :]
            hh = cv2.resize(hh, (W, H), interpolation=cv2.INTER_CUBIC)
            pp = cv2.resize(paf, (0, 0), fx=8, fy=8, interpolation=cv2.INTER_CUBIC)[:hs * 8 - pd, :ws * 8 - pr, :]
            pp = cv2.resize(pp, (W, H), interpolation=cv2.INTER_CUBIC)
            heat_up = heat_up + hh / len(scales)
            paf_up = paf_up + pp / len(scales)
    all_peaks = []
    counter = 0
    for part in range(18):
        ori = heat_up[:, :, part]
        m = gaussian_filter(ori, sigma=3)
        z = np.zeros(m.shape)
        a = z.copy(); a[1:, :] = m[:-1, :]
        b = z.copy(); b[:-1, :] = m[1:, :]
        c = z.copy(); c[:, 1:] = m[:, :-1]
        d = z.copy(); d[:, :-1] = m[:, 1:]
        binary = np.logical_and.reduce((m >= a, m >= b, m >= c, m >= d, m > thre1))
        ys, xs = np.nonzero(binary)
        all_peaks.append([(int(x), int(y), ori[y, x], counter + i) for i, (x, y) in enumerate(zip(xs, ys))])
        counter += len(xs)
    conn, special, _ = do.score_limbs(lambda ch, y, x: paf_up[y, x, ch], all_peaks, H, thre2)
    cand, sub, _ = do.assemble(all_peaks, conn, special)
    return cand, sub
