"""ORACLE (test infrastructure, never the product path): CPU restatement of the reference's
training-target path -- AugmentSelection.affine, cv2.warpAffine (INTER_CUBIC, fixed point),
the 368->46 mask resize, the keypoint transform and Heatmapper.create_heatmaps.

Every function cites the reference lines it follows (paths relative to /root/reference).
The third-party arithmetic (OpenCV 4.13 warpAffine / resize) is restated from its published
algorithm (imgwarp.cpp: initInterTab2D, remapBicubic, WarpAffineInvoker; resize.cpp:
HResizeCubic / VResizeCubic) and pinned against cv2 itself in tests/test_oracle_pin.py and
against outputs of the real reference in tests/golden/ (made by oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes
import ctypes.util
import math

import numpy as np

# C99 fma from the host libm (python < 3.13 has no math.fma)
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.fma.restype = ctypes.c_double
_libm.fma.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double]


def fma64(a, b, c):
    return _libm.fma(float(a), float(b), float(c))


# ---- constants: py_rmpe_server/py_rmpe_config.py:12-55 ----------------------------------
WIDTH = 368
HEIGHT = 368
STRIDE = 8
GRID = 46
NUM_PARTS = 18
LIMB_FROM = [2, 9, 10, 2, 12, 13, 2, 3, 4, 3, 2, 6, 7, 6, 2, 1, 1, 15, 16]
LIMB_TO = [9, 10, 11, 12, 13, 14, 3, 4, 5, 17, 6, 7, 8, 18, 1, 15, 16, 17, 18]
LIMBS_CONN = [(f - 1, t - 1) for f, t in zip(LIMB_FROM, LIMB_TO)]
PAF_LAYERS = 38
HEAT_START = 38
BKG_START = 56
NUM_LAYERS = 57
LEFT_PARTS = [5, 6, 7, 11, 12, 13, 15, 17]
RIGHT_PARTS = [2, 3, 4, 8, 9, 10, 14, 16]
TARGET_DIST = 0.6
SIGMA = 7.0
PAF_THRE = 8.0


# ---- T1: AugmentSelection.affine (py_rmpe_transformer.py:39-78) ---------------------------
def affine_chain(flip, degree, crop, scale, center, scale_self):
    """Literal restatement: five 3x3 matrices combined with four .dot() calls (:76)."""
    A = scale * math.cos(degree / 180. * math.pi)
    B = scale * math.sin(degree / 180. * math.pi)
    scale_size = TARGET_DIST / scale_self * scale
    cx = center[0] + crop[0]
    cy = center[1] + crop[1]
    c2z = np.array([[1., 0., -cx], [0., 1., -cy], [0., 0., 1.]])
    rot = np.array([[A, B, 0], [-B, A, 0], [0, 0, 1.]])
    sc = np.array([[scale_size, 0, 0], [0, scale_size, 0], [0, 0, 1.]])
    fl = np.array([[-1 if flip else 1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
    c2c = np.array([[1., 0., WIDTH // 2], [0., 1., HEIGHT // 2], [0., 0., 1.]])
    return c2c.dot(fl).dot(sc).dot(rot).dot(c2z)[0:2]


def affine_closed_form(flip, degree, crop, scale, center, scale_self):
    """Closed form of the chain above with the rounding sequence the numpy/BLAS 3x3 dots
    produce in this container; this is what the C host layer does."""
    A = scale * math.cos(degree / 180. * math.pi)
    B = scale * math.sin(degree / 180. * math.pi)
    s = TARGET_DIST / scale_self * scale
    cx = float(center[0] + crop[0])
    cy = float(center[1] + crop[1])
    f = -1.0 if flip else 1.0
    fs = f * s
    m00 = fs * A
    m01 = fs * B
    m10 = s * (-B)
    m11 = s * A
    m02 = fma64(m01, -cy, m00 * (-cx)) + float(WIDTH // 2)
    m12 = fma64(m11, -cy, m10 * (-cx)) + float(HEIGHT // 2)
    return np.array([[m00, m01, m02], [m10, m11, m12]], dtype=np.float64)


# ---- T2: cv2.warpAffine INTER_CUBIC BORDER_CONSTANT (py_rmpe_transformer.py:90-91) --------
_CUBIC_A = np.float32(-0.75)


def cubic_coeffs_f32(t):
    """OpenCV interpolateCubic (imgwarp.cpp): 1-D taps in float32, A=-0.75."""
    t = np.float32(t)
    one = np.float32(1)
    A = _CUBIC_A
    c0 = ((A * (t + one) - np.float32(5) * A) * (t + one) + np.float32(8) * A) * (t + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * t - (A + np.float32(3))) * t * t + one
    u = one - t
    c2 = ((A + np.float32(2)) * u - (A + np.float32(3))) * u * u + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=np.float32)


_tab_cache = {}


def bicubic_tab_i16():
    """OpenCV initInterTab2D(INTER_CUBIC, fixpt=true): [ay][ax][ky][kx] int16 summing to 32768."""
    if "tab" in _tab_cache:
        return _tab_cache["tab"]
    tab1 = np.stack([cubic_coeffs_f32(np.float32(i) * np.float32(1.0 / 32)) for i in range(32)])
    out = np.zeros((32, 32, 4, 4), dtype=np.int16)
    for i in range(32):
        for j in range(32):
            isum = 0
            it = np.zeros((4, 4), dtype=np.int64)
            for k1 in range(4):
                vy = tab1[i, k1]
                for k2 in range(4):
                    v = np.float32(vy * tab1[j, k2])
                    iv = int(np.rint(np.float32(v * np.float32(32768))))
                    iv = max(-32768, min(32767, iv))  # saturate_cast<short>
                    it[k1, k2] = iv
                    isum += iv
            if isum != 32768:
                diff = isum - 32768
                Mk1 = Mk2 = mk1 = mk2 = 2
                for k1 in range(2, 4):
                    for k2 in range(2, 4):
                        if it[k1, k2] < it[mk1, mk2]:
                            mk1, mk2 = k1, k2
                        elif it[k1, k2] > it[Mk1, Mk2]:
                            Mk1, Mk2 = k1, k2
                if diff < 0:
                    it[Mk1, Mk2] = it[Mk1, Mk2] - diff
                else:
                    it[mk1, mk2] = it[mk1, mk2] - diff
            out[i, j] = it.astype(np.int16)
    _tab_cache["tab"] = out
    return out


def invert_affine(M):
    """cv::warpAffine's inversion of the forward 2x3 matrix (imgwarp.cpp, f64)."""
    M = np.asarray(M, dtype=np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11 = M[1, 1] * D
    A22 = M[0, 0] * D
    i00 = A11
    i01 = M[0, 1] * (-D)
    i10 = M[1, 0] * (-D)
    i11 = A22
    b1 = -i00 * M[0, 2] - i01 * M[1, 2]
    b2 = -i10 * M[0, 2] - i11 * M[1, 2]
    return np.array([[i00, i01, b1], [i10, i11, b2]], dtype=np.float64)


def warp_coords(M, out_w=WIDTH, out_h=HEIGHT):
    """Fixed-point source coordinates (5 fractional bits) of every destination pixel."""
    iM = invert_affine(M)
    xs = np.arange(out_w, dtype=np.float64)
    ys = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(iM[0, 0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(iM[1, 0] * xs * 1024).astype(np.int64)
    X0 = np.rint((iM[0, 1] * ys + iM[0, 2]) * 1024).astype(np.int64) + 16
    Y0 = np.rint((iM[1, 1] * ys + iM[1, 2]) * 1024).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    return X, Y


def warp_affine_cubic_u8(src, M, border, out_w=WIDTH, out_h=HEIGHT, positions=None):
    """cv2.warpAffine(src, M, (out_w,out_h), INTER_CUBIC, BORDER_CONSTANT, border), u8.

    src (H,W) or (H,W,C) uint8.  `positions` = optional (ys, xs) index arrays to evaluate only
    a subset of destination pixels (used for the fused 46x46 mask)."""
    src = np.asarray(src)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    H, W, C = src.shape
    X, Y = warp_coords(M, out_w, out_h)
    # OpenCV saturates the int16 map; coordinates this far out are all-border anyway
    X = np.clip(X, -32768 * 32, 32767 * 32 + 31)
    Y = np.clip(Y, -32768 * 32, 32767 * 32 + 31)
    if positions is not None:
        X = X[positions]
        Y = Y[positions]
    sx = (X >> 5) - 1
    sy = (Y >> 5) - 1
    ax = X & 31
    ay = Y & 31
    tab = bicubic_tab_i16().astype(np.int64)
    w = tab[ay, ax]  # (..., 4, 4)
    acc = np.zeros(X.shape + (C,), dtype=np.int64)
    bval = np.broadcast_to(np.asarray(border, dtype=np.int64), (C,))
    for ky in range(4):
        yy = sy + ky
        for kx in range(4):
            xx = sx + kx
            inside = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            v = src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
            v = np.where(inside[..., None], v, bval)
            acc += w[..., ky, kx][..., None] * v
    out = np.clip((acc + 16384) >> 15, 0, 255).astype(np.uint8)
    if squeeze:
        out = out[..., 0]
    return out


# ---- T3: cv2.resize(mask, (46,46), INTER_CUBIC) (py_rmpe_transformer.py:92) ---------------
_MASK_W = np.array([-192, 1216, 1216, -192], dtype=np.int64)


def mask_resize_46(mask368):
    """Exact 8:1 bicubic of a 368x368 u8 plane: taps 8d+2..8d+5, horizontal pass exact int32,
    vertical pass a float32 FMA chain (OpenCV VResizeCubic on the int32 rows with
    beta*2^-22), rint, saturate."""
    m = np.asarray(mask368).astype(np.int64)
    idx = (np.arange(GRID) * 8)[:, None] + np.arange(2, 6)[None, :]  # (46,4) never clamps
    hor = (m[:, idx] * _MASK_W[None, None, :]).sum(axis=2)  # (368,46) int
    rows = hor[idx]  # (46,4,46)
    b = (_MASK_W.astype(np.float32) * np.float32(2.0 ** -22)).astype(np.float32)
    S = rows.astype(np.float32)
    v = S[:, 3, :] * b[3]
    v = _fma32(S[:, 2, :], b[2], v)
    v = _fma32(S[:, 1, :], b[1], v)
    v = _fma32(S[:, 0, :], b[0], v)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def _fma32(a, b, c):
    """float32 fused multiply-add emulated through float64 (exact product, one rounding:
    |a*b| < 2^53 here and double->float rounding of an exact f64 sum is the fma result except
    for double rounding, which cannot occur because a*b+c is exactly representable in f64
    for these magnitudes (24-bit x 12-bit products plus 24-bit addend))."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def mask46_from_src(src_mask, M):
    """Fused T2+T3 for the miss-mask: only rows/cols 8d+2..8d+5 of the warped mask are read by
    the resize, so evaluate the warp at those 184x184 positions only."""
    idx = ((np.arange(GRID) * 8)[:, None] + np.arange(2, 6)[None, :]).reshape(-1)  # 184
    ys, xs = np.meshgrid(idx, idx, indexing="ij")
    sub = warp_affine_cubic_u8(src_mask, M, 255, positions=(ys, xs)).astype(np.int64)  # (184,184)
    sub = sub.reshape(184, GRID, 4)
    hor = (sub * _MASK_W[None, None, :]).sum(axis=2)  # (184,46)
    rows = hor.reshape(GRID, 4, GRID)
    b = (_MASK_W.astype(np.float32) * np.float32(2.0 ** -22)).astype(np.float32)
    S = rows.astype(np.float32)
    v = S[:, 3, :] * b[3]
    v = _fma32(S[:, 2, :], b[2], v)
    v = _fma32(S[:, 1, :], b[1], v)
    v = _fma32(S[:, 0, :], b[0], v)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


# ---- T4: keypoint transform + flip swap (py_rmpe_transformer.py:100-111) -------------------
def transform_joints(joints, M, flip):
    """np.matmul(M, [x,y,1]) in the rounding order numpy produces (mul, fma, add); visibility
    untouched; left/right rows exchanged on flip."""
    j = np.array(joints, dtype=np.float64, copy=True)
    M = np.asarray(M, dtype=np.float64)
    x = j[:, :, 0].copy()
    y = j[:, :, 1].copy()
    for r in range(2):
        t = M[r, 0] * x
        t = np.array([fma64(M[r, 1], yy, tt) for yy, tt in zip(y.ravel(), t.ravel())]).reshape(t.shape)
        j[:, :, r] = t + M[r, 2]
    if flip:
        left = j[:, LEFT_PARTS, :].copy()
        right = j[:, RIGHT_PARTS, :].copy()
        j[:, LEFT_PARTS, :] = right
        j[:, RIGHT_PARTS, :] = left
    return j


def transform(img, mask, joints, M, flip):
    """Transformer.transform (py_rmpe_transformer.py:83-114) given the matrix."""
    out_img = warp_affine_cubic_u8(img, M, 127)
    m368 = warp_affine_cubic_u8(mask, M, 255)
    m46 = mask_resize_46(m368)
    return out_img, m46.astype(np.float64) / 255., transform_joints(joints, M, flip)


# ---- H1-H5: Heatmapper.create_heatmaps (py_rmpe_heatmapper.py:10-155) -----------------------
def _py_round(v):
    return int(round(float(v)))  # python-3 round = half to even (py_rmpe_heatmapper.py:95-98)


def create_heatmaps(joints, mask, sigma=SIGMA, thre=PAF_THRE, return_count=False, paf_average=False):
    """labels (57,46,46) f64.  count (19,46,46) int64 is the per-limb local `count` of
    put_vector_maps (:71,:122), which the reference computes and discards.  paf_average=True is the NON-reference
    variant spelt out in the reference's comments (:119-126): `+=` instead of `=` and `/= count` at the end."""
    joints = np.asarray(joints, dtype=np.float64)
    double_sigma2 = 2 * sigma * sigma
    grid = np.arange(GRID) * STRIDE + STRIDE / 2 - 0.5          # :22-23 cell centres
    Yg, Xg = np.mgrid[0:HEIGHT:STRIDE, 0:WIDTH:STRIDE]           # :25 cell top-left (ints)
    heat = np.zeros((NUM_LAYERS, GRID, GRID), dtype=np.float64)
    counts = np.zeros((len(LIMBS_CONN), GRID, GRID), dtype=np.int64)
    # put_joints / put_gaussian_maps (:47-66)
    for i in range(NUM_PARTS):
        for p in range(joints.shape[0]):
            if not (joints[p, i, 2] < 2):
                continue
            ex = np.exp(-(grid - joints[p, i, 0]) ** 2 / double_sigma2)
            ey = np.exp(-(grid - joints[p, i, 1]) ** 2 / double_sigma2)
            heat[HEAT_START + i] = np.maximum(heat[HEAT_START + i], np.outer(ey, ex))
    heat[BKG_START] = 1. - np.amax(heat[HEAT_START:HEAT_START + NUM_PARTS], axis=0)  # :37-38
    # put_limbs / put_vector_maps (:69-138)
    for k, (fr, to) in enumerate(LIMBS_CONN):
        for p in range(joints.shape[0]):
            if not (joints[p, fr, 2] < 2 and joints[p, to, 2] < 2):
                continue
            x1, y1 = joints[p, fr, 0], joints[p, fr, 1]
            x2, y2 = joints[p, to, 0], joints[p, to, 1]
            dx = x2 - x1
            dy = y2 - y1
            dnorm = math.sqrt(dx * dx + dy * dy)
            if dnorm == 0:
                continue
            ux = dx / dnorm
            uy = dy / dnorm
            min_sx, max_sx = (x1, x2) if x1 < x2 else (x2, x1)
            min_sy, max_sy = (y1, y2) if y1 < y2 else (y2, y1)
            min_sx = _py_round((min_sx - thre) / STRIDE)
            min_sy = _py_round((min_sy - thre) / STRIDE)
            max_sx = _py_round((max_sx + thre) / STRIDE)
            max_sy = _py_round((max_sy + thre) / STRIDE)
            if max_sy < 0 or max_sx < 0:
                continue
            min_sx = max(min_sx, 0)
            min_sy = max(min_sy, 0)
            sl = (slice(min_sy, max_sy), slice(min_sx, max_sx))
            X = Xg[sl]
            Y = Yg[sl]
            # distances() (:144-155): un-fused f64, line (not segment) distance
            xD = x2 - x1
            yD = y2 - y1
            norm2 = math.sqrt(xD ** 2 + yD ** 2)
            dist = xD * (y1 - Y) - (x1 - X) * yD
            dist = dist / norm2
            on = np.abs(dist) <= thre
            if paf_average:
                heat[2 * k][sl][on] += ux          # "# += dist * dx" (:120)
                heat[2 * k + 1][sl][on] += uy
            else:
                heat[2 * k][sl][on] = ux
                heat[2 * k + 1][sl][on] = uy
            counts[k][sl][on] += 1
        if paf_average:                            # "# heatmaps[layerX, :, :][count > 0] /= count[count > 0]" (:125-126)
            nz = counts[k] > 0
            heat[2 * k][nz] /= counts[k][nz]
            heat[2 * k + 1][nz] /= counts[k][nz]
    heat *= np.asarray(mask, dtype=np.float64)  # :42
    if return_count:
        return heat, counts
    return heat


# ---- next row: DataIteratorBase.gen batch assembly (training/ds_generators.py:47-77) ---------
def keras_batch(labels, mask):
    """labels (B,57,46,46), mask (B,46,46) -> x1, x2, y1, y2 exactly as the reference builds them per
    sample (np.repeat of the mask :52-53, label split at 38 + HWC transposes :59-62, np.concatenate :75-79)."""
    x1, x2, y1, y2 = [], [], [], []
    for lab, m in zip(labels, mask):
        x1.append(np.repeat(m[:, :, np.newaxis], PAF_LAYERS, axis=2)[np.newaxis, ...])
        x2.append(np.repeat(m[:, :, np.newaxis], NUM_LAYERS - PAF_LAYERS, axis=2)[np.newaxis, ...])
        y1.append(np.transpose(lab[:PAF_LAYERS, :, :], (1, 2, 0))[np.newaxis, ...])
        y2.append(np.transpose(lab[PAF_LAYERS:, :, :], (1, 2, 0))[np.newaxis, ...])
    return np.concatenate(x1), np.concatenate(x2), np.concatenate(y1), np.concatenate(y2)
