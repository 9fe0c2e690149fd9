"""ORACLE (test infrastructure, never the product path): CPU restatement of the reference's
inference decode -- eval/eval_coco2014_multi_modes.py process_single_scale (:263-415) and
process_multi_scale (:58-231), starting at the network's output blobs, and util.py
padRightDownCorner (:57-77).

Third-party arithmetic restated from its published algorithm and pinned against the installed
libraries in tests/test_oracle_pin.py:
  * cv2.resize INTER_CUBIC on many-channel float32 (OpenCV 4.13 resize.cpp HResizeCubic /
    VResizeCubic + VResizeCubicVec_32f, SSE baseline: no FMA; vector body sums taps 3->0, the
    scalar row tail sums 0->3)  -- bit-exact.
  * scipy.ndimage.gaussian_filter sigma=3 (ni_filters.c NI_Correlate1D symmetric branch,
    reflect mode, f64 accumulation, output dtype = input dtype) -- bit-exact.
Pinned against the real reference's outputs in tests/golden/ (oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import math

import numpy as np

from .gt_oracle import cubic_coeffs_f32

f32 = np.float32

# eval/eval_coco2014_multi_modes.py:28-35
LIMB_SEQ = [[2, 3], [2, 6], [3, 4], [4, 5], [6, 7], [7, 8], [2, 9], [9, 10],
            [10, 11], [2, 12], [12, 13], [13, 14], [2, 1], [1, 15], [15, 17],
            [1, 16], [16, 18], [3, 17], [6, 18]]
MAP_IDX = [[31, 32], [39, 40], [33, 34], [35, 36], [41, 42], [43, 44], [19, 20], [21, 22],
           [23, 24], [25, 26], [27, 28], [29, 30], [47, 48], [49, 50], [53, 54], [51, 52],
           [55, 56], [37, 38], [45, 46]]
MID_NUM = 10


# ---- U1: util.padRightDownCorner (util.py:57-77) -------------------------------------------
def pad_right_down_corner(img, stride, pad_value):
    h, w = img.shape[0], img.shape[1]
    pad = [0, 0, 0 if h % stride == 0 else stride - h % stride,
           0 if w % stride == 0 else stride - w % stride]
    out = np.full((h + pad[2], w + pad[3]) + img.shape[2:], pad_value, dtype=img.dtype)
    out[:h, :w] = img
    return out, pad


# ---- D1: cv2.resize(blob, (W,H), INTER_CUBIC) on (h,w,C) float32 ------------------------------
def resize_axis_table(dst, src, inv_scale=None):
    """Per destination index: first tap s-1 (unclamped), 4 float32 cubic coeffs.
    resize.cpp: scale = 1/inv_scale (double); f = float((d+0.5)*scale-0.5); s=floor(f)."""
    if inv_scale is None:
        inv_scale = float(dst) / float(src)
    scale = 1.0 / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    t = (f - s.astype(np.float32)).astype(np.float32)
    co = np.stack([cubic_coeffs_f32(tt) for tt in t]) if dst > 0 else np.zeros((0, 4), f32)
    return s, co.astype(np.float32)


def resize_cubic_f32(src, W, H, inv_fx=None, inv_fy=None):
    """(h,w,C) float32 -> (H,W,C) float32, replicate border, horizontal then vertical."""
    src = np.ascontiguousarray(src, dtype=np.float32)
    if src.ndim == 2:
        src = src[:, :, None]
    h, w, C = src.shape
    sx, cx = resize_axis_table(W, w, inv_fx)
    sy, cy = resize_axis_table(H, h, inv_fy)
    ix = np.clip(sx[:, None] + np.arange(-1, 3)[None, :], 0, w - 1)
    iy = np.clip(sy[:, None] + np.arange(-1, 3)[None, :], 0, h - 1)
    # horizontal: ((S0*a0 + S1*a1) + S2*a2) + S3*a3, every op rounded to float32
    S = src[:, ix, :]                      # (h,W,4,C)
    a = cx[None, :, :, None]
    hp = S[:, :, 0] * a[:, :, 0]
    for k in range(1, 4):
        hp = hp + S[:, :, k] * a[:, :, k]
    # vertical: vector body S0*b0 + (S1*b1 + (S2*b2 + S3*b3)); scalar tail left-to-right
    R = hp[iy]                             # (H,4,W,C)
    b = cy[:, :, None, None]
    out = R[:, 3] * b[:, 3]
    for k in (2, 1, 0):
        out = R[:, k] * b[:, k] + out
    tail = (W * C) % 4
    if tail:
        alt = R[:, 0] * b[:, 0]
        for k in range(1, 4):
            alt = alt + R[:, k] * b[:, k]
        o2 = out.reshape(H, W * C)
        o2[:, W * C - tail:] = alt.reshape(H, W * C)[:, W * C - tail:]
        out = o2.reshape(H, W, C)
    assert out.dtype == np.float32
    return out


# ---- D2: gaussian_filter(sigma=3) + 4-neighbour peak test ------------------------------------
def gaussian_weights(sigma=3.0, truncate=4.0):
    radius = int(truncate * sigma + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum(), radius


def _correlate1d_sym(x64, w, radius, axis):
    pad = [(0, 0)] * x64.ndim
    pad[axis] = (radius, radius)
    xp = np.pad(x64, pad, mode="symmetric")
    n = x64.shape[axis]

    def sl(off):
        idx = [slice(None)] * x64.ndim
        idx[axis] = slice(radius + off, radius + off + n)
        return tuple(idx)

    tmp = xp[sl(0)] * w[radius]
    for j in range(-radius, 0):
        tmp = tmp + (xp[sl(j)] + xp[sl(-j)]) * w[radius + j]
    return tmp


def gaussian_filter_sigma3(m):
    """scipy.ndimage.gaussian_filter(m, sigma=3) for a 2-D map (axis 0 then axis 1); the line
    buffers are f64, each axis' result is stored in m.dtype (eval...:99, :285)."""
    w, r = gaussian_weights()
    dt = m.dtype
    a = _correlate1d_sym(m.astype(np.float64), w, r, 0).astype(dt)
    return _correlate1d_sym(a.astype(np.float64), w, r, 1).astype(dt)


def find_peaks(heat_hw_c, thre1):
    """eval...:97-118 / :283-304.  heat (H,W,>=18).  Returns all_peaks (list of 18 lists of
    (x, y, score, id)), candidate (N,4) f64."""
    all_peaks = []
    counter = 0
    for part in range(18):
        ori = heat_hw_c[:, :, part]
        m = gaussian_filter_sigma3(ori)
        z = np.zeros(m.shape)
        left = z.copy(); left[1:, :] = m[:-1, :]
        right = z.copy(); right[:-1, :] = m[1:, :]
        up = z.copy(); up[:, 1:] = m[:, :-1]
        down = z.copy(); down[:, :-1] = m[:, 1:]
        binary = (m >= left) & (m >= right) & (m >= up) & (m >= down) & (m > thre1)
        ys, xs = np.nonzero(binary)
        peaks = [(int(x), int(y), ori[y, x], counter + i) for i, (x, y) in enumerate(zip(xs, ys))]
        counter += len(peaks)
        all_peaks.append(peaks)
    return all_peaks


def _py_round(v):
    return int(round(float(v)))


def score_limbs(paf_at, all_peaks, img_h, thre2):
    """eval...:126-176 / :310-360.  paf_at(c, y, x) -> value of PAF channel c at integer (y,x)
    (float32 single-scale / float64 multi-scale); products and sums are f64 (numpy>=2
    promotion of f32-array * np.float64 scalar).  Returns connection_all, special_k, and the
    raw candidate lists (before sorting) for parity checks."""
    connection_all, special_k, cand_all = [], [], []
    for k in range(len(MAP_IDX)):
        cxi, cyi = MAP_IDX[k][0] - 19, MAP_IDX[k][1] - 19
        candA = all_peaks[LIMB_SEQ[k][0] - 1]
        candB = all_peaks[LIMB_SEQ[k][1] - 1]
        nA, nB = len(candA), len(candB)
        if nA == 0 or nB == 0:
            special_k.append(k)
            connection_all.append([])
            cand_all.append([])
            continue
        cands = []
        for i in range(nA):
            for j in range(nB):
                vx = candB[j][0] - candA[i][0]
                vy = candB[j][1] - candA[i][1]
                norm = math.sqrt(vx * vx + vy * vy)
                if norm == 0:
                    continue
                ux = float(vx) / norm
                uy = float(vy) / norm
                ax, ay, bx, by = candA[i][0], candA[i][1], candB[j][0], candB[j][1]
                stepx = (bx - ax) / 9.0
                stepy = (by - ay) / 9.0
                s_sum = 0
                n_ok = 0
                for I in range(MID_NUM):
                    if I == MID_NUM - 1:
                        px, py = float(bx), float(by)
                    else:
                        px = I * stepx + ax
                        py = I * stepy + ay
                    xi, yi = _py_round(px), _py_round(py)
                    s = float(paf_at(cxi, yi, xi)) * ux + float(paf_at(cyi, yi, xi)) * uy
                    s_sum = s_sum + s
                    if s > thre2:
                        n_ok += 1
                prior = min(0.5 * img_h / norm - 1, 0)
                score = s_sum / MID_NUM + prior
                if n_ok > 0.8 * MID_NUM and score > 0:
                    cands.append([i, j, score, score + float(candA[i][2]) + float(candB[j][2])])
        cand_all.append(cands)
        srt = sorted(cands, key=lambda x: x[2], reverse=True)     # stable
        conn = np.zeros((0, 5))
        for c in srt:
            i, j, s = c[0:3]
            if i not in conn[:, 3] and j not in conn[:, 4]:
                conn = np.vstack([conn, [candA[i][3], candB[j][3], s, i, j]])
                if len(conn) >= min(nA, nB):
                    break
        connection_all.append(conn)
    return connection_all, special_k, cand_all


def assemble(all_peaks, connection_all, special_k):
    """eval...:180-231 / :364-415.  Returns (candidate (N,4), subset (M,20), overflow) where
    overflow=True marks the reference's unguarded `found > 2` IndexError."""
    candidate = np.array([item for sub in all_peaks for item in sub], dtype=np.float64).reshape(-1, 4)
    subset = -1 * np.ones((0, 20))
    overflow = False
    for k in range(len(MAP_IDX)):
        if k in special_k:
            continue
        conn = connection_all[k]
        partAs = conn[:, 0]
        partBs = conn[:, 1]
        indexA, indexB = np.array(LIMB_SEQ[k]) - 1
        for i in range(len(conn)):
            found = 0
            subset_idx = [-1, -1]
            for j in range(len(subset)):
                if subset[j][indexA] == partAs[i] or subset[j][indexB] == partBs[i]:
                    if found >= 2:
                        overflow = True
                        continue
                    subset_idx[found] = j
                    found += 1
            if found == 1:
                j = subset_idx[0]
                if subset[j][indexB] != partBs[i]:
                    subset[j][indexB] = partBs[i]
                    subset[j][-1] += 1
                    subset[j][-2] += candidate[int(partBs[i]), 2] + conn[i][2]
            elif found == 2:
                j1, j2 = subset_idx
                membership = ((subset[j1] >= 0).astype(int) + (subset[j2] >= 0).astype(int))[:-2]
                if len(np.nonzero(membership == 2)[0]) == 0:
                    subset[j1][:-2] += (subset[j2][:-2] + 1)
                    subset[j1][-2:] += subset[j2][-2:]
                    subset[j1][-2] += conn[i][2]
                    subset = np.delete(subset, j2, 0)
                else:
                    subset[j1][indexB] = partBs[i]
                    subset[j1][-1] += 1
                    subset[j1][-2] += candidate[int(partBs[i]), 2] + conn[i][2]
            elif not found and k < 17:
                row = -1 * np.ones(20)
                row[indexA] = partAs[i]
                row[indexB] = partBs[i]
                row[-1] = 2
                row[-2] = sum(candidate[conn[i, :2].astype(int), 2]) + conn[i][2]
                subset = np.vstack([subset, row])
    keep = [i for i in range(len(subset))
            if not (subset[i][-1] < 4 or subset[i][-2] / subset[i][-1] < 0.4)]
    subset = subset[keep] if len(subset) else subset
    return candidate, subset, overflow


def decode_maps(heat_up, paf_at, img_h, thre1=0.1, thre2=0.05, detail=False):
    all_peaks = find_peaks(heat_up, thre1)
    connection_all, special_k, cand_all = score_limbs(paf_at, all_peaks, img_h, thre2)
    candidate, subset, overflow = assemble(all_peaks, connection_all, special_k)
    if detail:
        return dict(all_peaks=all_peaks, connection_all=connection_all, special_k=special_k,
                    limb_candidates=cand_all, candidate=candidate, subset=subset, overflow=overflow)
    return candidate, subset


def single_scale(paf, heat, H, W, thre1=0.1, thre2=0.05, detail=False):
    """process_single_scale from the blobs on (eval...:277-415). paf (h,w,38), heat (h,w,19)."""
    heat_up = resize_cubic_f32(heat, W, H)
    paf_up = resize_cubic_f32(paf, W, H)
    r = decode_maps(heat_up, lambda c, y, x: paf_up[y, x, c], H, thre1, thre2, detail)
    if detail:
        r["heat_up"] = heat_up
        r["paf_up"] = paf_up
    return r


def multi_scale_average(blobs, H, W, stride=8):
    """eval...:79-92.  blobs = list over scales of (paf (hs,ws,38), heat (hs,ws,19), pad_down,
    pad_right) with hs*8, ws*8 the padded feed size.  Returns heat_avg, paf_avg (f64)."""
    heat_avg = np.zeros((H, W, 19))
    paf_avg = np.zeros((H, W, 38))
    n = len(blobs)
    for paf, heat, pad_d, pad_r in blobs:
        hs, ws = heat.shape[:2]
        hh = resize_cubic_f32(heat, ws * stride, hs * stride, float(stride), float(stride))
        hh = hh[:hs * stride - pad_d, :ws * stride - pad_r, :]
        hh = resize_cubic_f32(hh, W, H)
        pp = resize_cubic_f32(paf, ws * stride, hs * stride, float(stride), float(stride))
        pp = pp[:hs * stride - pad_d, :ws * stride - pad_r, :]
        pp = resize_cubic_f32(pp, W, H)
        heat_avg = heat_avg + hh / n
        paf_avg = paf_avg + pp / n
    return heat_avg, paf_avg


def multi_scale(blobs, H, W, thre1=0.1, thre2=0.05, stride=8, detail=False):
    heat_avg, paf_avg = multi_scale_average(blobs, H, W, stride)
    r = decode_maps(heat_avg, lambda c, y, x: paf_avg[y, x, c], H, thre1, thre2, detail)
    if detail:
        r["heat_up"] = heat_avg
        r["paf_up"] = paf_avg
    return r


def multi_scale_feed_shapes(H, W, scale_search=(0.5, 1, 1.5, 2), boxsize=368, stride=8):
    """Per scale: resized image size (cv2.resize fx=fy=m -> cvRound) and the padded feed size
    (eval...:61,69-71): returns list of (Hs, Ws, pad_down, pad_right, hs, ws)."""
    out = []
    for x in scale_search:
        m = x * boxsize / H
        Ws = int(np.rint(W * m))
        Hs = int(np.rint(H * m))
        pd = 0 if Hs % stride == 0 else stride - Hs % stride
        pr = 0 if Ws % stride == 0 else stride - Ws % stride
        out.append((Hs, Ws, pd, pr, (Hs + pd) // stride, (Ws + pr) // stride))
    return out
