"""Seeded synthetic COCO-shaped inputs for the five BASELINE.json configs.

No dataset or checkpoint is reachable, so GT samples and decode blobs are synthesised with the
shapes the reference handles: 368x368x3 u8 crops + u8 miss-mask + (P,18,3) joints for the
target generator, and (h,w,38) PAF / (h,w,19) heat fp32 NHWC blobs for the decoder.
"""
import random as _pyrandom

import numpy as np

from .py_rmpe_server.py_rmpe_config import RmpeGlobalConfig

# 18-joint template (nose..Lear in RmpeGlobalConfig.parts order) in a unit box: x in [-0.5,0.5]
# (image-left is the person's right side), y in [0,1] top to bottom.
_TEMPLATE = np.array([
    [0.00, 0.08],   # nose
    [0.00, 0.20],   # neck
    [-0.20, 0.20],  # Rsho
    [-0.30, 0.38],  # Relb
    [-0.34, 0.54],  # Rwri
    [0.20, 0.20],   # Lsho
    [0.30, 0.38],   # Lelb
    [0.34, 0.54],   # Lwri
    [-0.12, 0.55],  # Rhip
    [-0.14, 0.77],  # Rkne
    [-0.15, 0.98],  # Rank
    [0.12, 0.55],   # Lhip
    [0.14, 0.77],   # Lkne
    [0.15, 0.98],   # Lank
    [-0.04, 0.05],  # Reye
    [0.04, 0.05],   # Leye
    [-0.09, 0.07],  # Rear
    [0.09, 0.07],   # Lear
], dtype=np.float64)


def template_person(rng, frame_h, frame_w, min_frac=0.35, max_frac=0.8, integer=False):
    """One person: template scaled to height U(min,max)*frame_h, aspect 0.5, origin uniform."""
    ph = rng.uniform(min_frac, max_frac) * frame_h
    pw = 0.5 * ph * 2.0  # template x spans [-0.5,0.5] of a box half as wide as tall, doubled
    ox = rng.uniform(0.15 * frame_w, 0.85 * frame_w)
    oy = rng.uniform(0.0, max(1.0, frame_h - ph))
    j = np.zeros((18, 3), dtype=np.float64)
    j[:, 0] = ox + _TEMPLATE[:, 0] * pw * 0.5
    j[:, 1] = oy + _TEMPLATE[:, 1] * ph
    if integer:
        j[:, :2] = np.rint(j[:, :2])
    return j


def gt_sample(seed, n_persons=3, src_hw=(368, 368), augment=True, integer_joints=False):
    """One synthetic GT sample (SURVEY.md 8d 'Synthetic sample, GT').

    Returns dict(img u8 HxWx3, mask u8 HxW, joints (P,18,3) f64, objpos, scale_provided,
    aug=(flip, degree, (cx,cy), scale))."""
    rng = np.random.RandomState(seed)
    H, W = src_hw
    img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    mask = np.full((H, W), 255, dtype=np.uint8)
    if seed % 4 == 0:
        mask[100:180, 200:300] = 0
    joints = np.stack([template_person(rng, H, W, integer=integer_joints) for _ in range(n_persons)]) \
        if n_persons > 0 else np.zeros((0, 18, 3))
    if n_persons > 0:
        vis = np.where(rng.uniform(size=(n_persons, 18)) < 0.1, 2.0, 1.0)
        joints[:, :, 2] = vis
    if augment:
        aug = random_aug(seed)
    else:
        aug = (False, 0., (0, 0), 1.)
    return dict(img=img, mask=mask, joints=joints, objpos=[[W / 2.0, H / 2.0]],
                scale_provided=[0.6], aug=aug)


def random_aug(seed):
    """AugmentSelection.random() draw order with python's `random` seeded by the sample index
    (py_rmpe_transformer.py:19-27): flip, degree, scale-condition, [scale], x_off, y_off."""
    from .py_rmpe_server.py_rmpe_config import TransformationParams as TP
    r = _pyrandom.Random(seed)
    flip = r.uniform(0., 1.) > TP.flip_prob
    degree = r.uniform(-1., 1.) * TP.max_rotate_degree
    scale = (TP.scale_max - TP.scale_min) * r.uniform(0., 1.) + TP.scale_min \
        if r.uniform(0., 1.) > TP.scale_prob else 1.
    x_off = int(r.uniform(-1., 1.) * TP.center_perterb_max)
    y_off = int(r.uniform(-1., 1.) * TP.center_perterb_max)
    return (flip, degree, (x_off, y_off), scale)


def gt_batch(batch, n_persons=3, seed0=0, src_hw=(368, 368), augment=True):
    """Stacked arrays for `batch` samples: imgs (B,H,W,3) u8, masks (B,H,W) u8, joints
    (B,P,18,3) f64, n_persons (B,) i32, aug tuple list, centers (B,2), scale_self (B,)."""
    samples = [gt_sample(seed0 + i, n_persons, src_hw, augment) for i in range(batch)]
    return dict(
        imgs=np.stack([s["img"] for s in samples]),
        masks=np.stack([s["mask"] for s in samples]),
        joints=np.stack([s["joints"] for s in samples]),
        n_persons=np.full((batch,), n_persons, dtype=np.int32),
        augs=[s["aug"] for s in samples],
        centers=np.array([s["objpos"][0] for s in samples], dtype=np.float64),
        scale_self=np.array([s["scale_provided"][0] for s in samples], dtype=np.float64),
    )


# ------------------------------------------------------------------------------------------
# decode blobs
# ------------------------------------------------------------------------------------------
_LIMBS = RmpeGlobalConfig.limbs_conn


def decode_blobs(seed, frame_hw, grid_hw, n_persons=3, noise=0.01, persons=None, stride=None):
    """Synthetic network output for one frame on a (h,w) grid: reference rasteriser semantics
    (Gaussians sigma=7 px, PAF band 8 px limited to the segment +-8 px) evaluated at grid-cell
    positions for template persons placed at image scale, plus N(0, noise^2).

    Returns paf (h,w,38) f32, heat (h,w,19) f32, persons (P,18,3)."""
    rng = np.random.RandomState(seed)
    H, W = frame_hw
    h, w = grid_hw
    if persons is None:
        persons = np.stack([template_person(rng, H, W, 0.3, 0.75) for _ in range(n_persons)]) \
            if n_persons > 0 else np.zeros((0, 18, 3))
        persons[:, :, 2] = np.where(rng.uniform(size=(n_persons, 18)) < 0.1, 2.0, 1.0)
    sy = H / float(h) if stride is None else float(stride)
    sx = W / float(w) if stride is None else float(stride)
    gy = (np.arange(h) + 0.5) * sy - 0.5
    gx = (np.arange(w) + 0.5) * sx - 0.5
    heat = np.zeros((h, w, 19), dtype=np.float64)
    for p in range(persons.shape[0]):
        for i in range(18):
            if persons[p, i, 2] >= 2:
                continue
            ex = np.exp(-(gx - persons[p, i, 0]) ** 2 / 98.0)
            ey = np.exp(-(gy - persons[p, i, 1]) ** 2 / 98.0)
            heat[:, :, i] = np.maximum(heat[:, :, i], np.outer(ey, ex))
    heat[:, :, 18] = 1.0 - heat[:, :, :18].max(axis=2)
    paf = np.zeros((h, w, 38), dtype=np.float64)
    GY, GX = np.meshgrid(gy, gx, indexing="ij")
    for k, (fr, to) in enumerate(_LIMBS):
        for p in range(persons.shape[0]):
            if persons[p, fr, 2] >= 2 or persons[p, to, 2] >= 2:
                continue
            x1, y1 = persons[p, fr, :2]
            x2, y2 = persons[p, to, :2]
            dx, dy = x2 - x1, y2 - y1
            n = np.hypot(dx, dy)
            if n == 0:
                continue
            ux, uy = dx / n, dy / n
            along = (GX - x1) * ux + (GY - y1) * uy
            perp = np.abs((GX - x1) * uy - (GY - y1) * ux)
            on = (perp <= 8.0) & (along >= -8.0) & (along <= n + 8.0)
            paf[:, :, 2 * k][on] = ux
            paf[:, :, 2 * k + 1][on] = uy
    heat += rng.normal(0.0, noise, size=heat.shape)
    paf += rng.normal(0.0, noise, size=paf.shape)
    return paf.astype(np.float32), heat.astype(np.float32), persons


def single_scale_grid(H, W):
    """Output grid of the testing model for an unpadded (H,W) feed: three floor-halvings
    (SURVEY.md 8a D0)."""
    return ((H // 2) // 2) // 2, ((W // 2) // 2) // 2


def multi_scale_feed_shapes(H, W, scale_search=(0.5, 1, 1.5, 2), boxsize=368, stride=8):
    """eval_coco2014_multi_modes.py:61,69-71: per scale (resized height, pad_down, pad_right, blob rows, blob columns)."""
    out = []
    for x in scale_search:
        m = x * boxsize / H
        Ws, Hs = int(np.rint(W * m)), int(np.rint(H * m))
        pd = 0 if Hs % stride == 0 else stride - Hs % stride
        pr = 0 if Ws % stride == 0 else stride - Ws % stride
        out.append((Hs, pd, pr, (Hs + pd) // stride, (Ws + pr) // stride))
    return out


def multi_scale_frame(seed, H, W, n_persons=3):
    """One synthetic frame for process_multi_scale-style decode: the same persons rendered on the blob grid of every scale."""
    _, _, persons = decode_blobs(seed, (H, W), (4, 4), n_persons)
    sc = []
    for (Hs, pd, pr, hs, ws) in multi_scale_feed_shapes(H, W):
        paf, heat, _ = decode_blobs(seed + 1000 * len(sc), (H, W), (hs, ws), n_persons, persons=persons, stride=8.0 * H / Hs)
        sc.append((paf, heat, pd, pr))
    return dict(H=H, W=W, scales=sc)



def dense_frame(seed, H, W, coarse=12):
    """Worst case of the decoder's screening: a smooth random field around 0.3 in every heat channel and random PAFs --
    every (tile, part) pair is above thre1 somewhere, nothing is culled, ~25 peaks per part on a ski.jpg-sized frame.
    (Real network output sits between this and the sparse template frames of decode_blobs.)"""
    import cv2
    rng = np.random.RandomState(seed)
    h, w = single_scale_grid(H, W)

    def field(c, amp, base):
        z = rng.normal(size=(h // coarse + 2, w // coarse + 2, c)).astype(np.float32)
        z = cv2.resize(z, (w, h), interpolation=cv2.INTER_CUBIC)
        return (base + amp * z.reshape(h, w, c)).astype(np.float32)

    return dict(H=H, W=W, scales=[(field(38, 0.3, 0.0), field(19, 0.15, 0.3), 0, 0)])
