"""Batched entry points over the C ABI.

  gt_batch_host / decode_batch_host  : numpy (host) buffers in and out -- what the reference-style
                                       classes (Transformer, Heatmapper, process_*_scale) call;
                                       copies happen inside the C call.
  GtDevicePlan / DecodeDevicePlan    : torch CUDA tensors resident in HBM, asynchronous on the
                                       current torch stream -- what bench.py's `value` leg times.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .py_rmpe_server.py_rmpe_config import RmpeGlobalConfig as G

NL = G.num_layers
GRID = G.height // G.stride


# ------------------------------------------------------------------------------------------
# T1 helpers (host, C)
# ------------------------------------------------------------------------------------------
def aug_affine(flip, degree, crop_xy, scale, center_xy, scale_self):
    """Vectorised AugmentSelection.affine: returns (n,2,3) float64."""
    lib = L.load()
    flip = L.c_contig(flip, np.uint8).reshape(-1)
    n = flip.shape[0]
    degree = L.c_contig(degree, np.float64).reshape(n)
    crop_xy = L.c_contig(crop_xy, np.int32).reshape(n, 2)
    scale = L.c_contig(scale, np.float64).reshape(n)
    center_xy = L.c_contig(center_xy, np.float64).reshape(n, 2)
    scale_self = L.c_contig(scale_self, np.float64).reshape(n)
    M = np.empty((n, 2, 3), dtype=np.float64)
    L.check(lib.rmpe_aug_affine(n, L.ptr(flip), L.ptr(degree), L.ptr(crop_xy), L.ptr(scale),
                                L.ptr(center_xy), L.ptr(scale_self), L.ptr(M)))
    return M


def aug_random(seeds):
    """AugmentSelection.random() for each seed (random.seed(seed) semantics): returns
    flip (n,) u8, degree (n,), crop_xy (n,2) i32, scale (n,)."""
    lib = L.load()
    seeds = L.c_contig(seeds, np.uint64).reshape(-1)
    n = seeds.shape[0]
    flip = np.empty(n, np.uint8)
    degree = np.empty(n, np.float64)
    crop = np.empty((n, 2), np.int32)
    scale = np.empty(n, np.float64)
    L.check(lib.rmpe_aug_random(n, L.ptr(seeds), L.ptr(flip), L.ptr(degree), L.ptr(crop), L.ptr(scale)))
    return flip, degree, crop, scale


# ------------------------------------------------------------------------------------------
# GT, host buffers
# ------------------------------------------------------------------------------------------
def gt_batch_host(imgs, masks, joints, n_persons, M, flip, *, f64=False, chw=False, want_img=True,
                  want_labels=True, want_count=False, simple=False, out=None, sigma=None, thre=None,
                  keras=False, keras_weights=True, paf_average=False):
    """imgs (B,H,W,3) u8, masks (B,H,W) u8, joints (B,P,18,3) f64, n_persons (B,) i32,
    M (B,2,3) f64, flip (B,) u8.  Returns dict(img, mask, labels, joints, count, status).
    `out` may carry preallocated (e.g. pinned) output arrays under the same keys.
    keras=True adds the NHWC tensors of DataIteratorBase.gen (training/ds_generators.py:52-63), written by the
    rasteriser itself: y1 (B,46,46,38), y2 (B,46,46,19) and, with keras_weights, x1 / x2 (the mask repeated)."""
    lib = L.ensure_init()
    masks = L.c_contig(masks, np.uint8)
    B, H, W = masks.shape
    if want_img:
        imgs = L.c_contig(imgs, np.uint8)
        assert imgs.shape == (B, H, W, 3), imgs.shape
    joints = L.c_contig(joints, np.float64)
    P = joints.shape[1] if joints.ndim == 4 else 0
    joints = joints.reshape(B, P, 18, 3)
    n_persons = L.c_contig(n_persons, np.int32).reshape(B)
    M = L.c_contig(M, np.float64).reshape(B, 6)
    flip = L.c_contig(flip, np.uint8).reshape(B)
    ft = np.float64 if f64 else np.float32
    out = dict(out or {})

    def buf(key, shape, dtype):
        a = out.get(key)
        if a is None:
            a = np.empty(shape, dtype)
        assert a.shape == tuple(shape) and a.dtype == dtype and a.flags.c_contiguous, (key, a.shape, a.dtype)
        return a

    res = {}
    res["img"] = buf("img", (B, 3, G.height, G.width) if chw else (B, G.height, G.width, 3), np.uint8) if want_img else None
    res["mask"] = buf("mask", (B, GRID, GRID), ft)
    res["labels"] = buf("labels", (B, NL, GRID, GRID), ft) if want_labels else None
    res["joints"] = buf("joints", (B, P, 18, 3), np.float64)
    res["count"] = np.empty((B, 19, GRID, GRID), np.int32) if want_count else None
    res["status"] = np.zeros(B, np.int32)
    for k, c in (("y1", 38), ("y2", 19), ("x1", 38), ("x2", 19)):
        res[k] = buf(k, (B, GRID, GRID, c), ft) if keras and (keras_weights or k[0] == "y") else None
    h = L.GtBatchHost()
    h.batch = B
    h.max_persons = P
    h.flags = (L.GT_LABELS_F64 if f64 else 0) | (L.GT_IMG_CHW if chw else 0) | \
        (0 if want_img else L.GT_NO_WARP) | (L.GT_SIMPLE_KERNELS if simple else 0) | \
        (L.GT_PAF_AVERAGE if paf_average else 0)
    h.src_height, h.src_width = H, W
    h.src_img = L.ptr(imgs) if want_img else None
    h.src_mask = L.ptr(masks)
    h.joints = L.ptr(joints) if P else None
    h.n_persons = L.ptr(n_persons)
    h.M = L.ptr(M)
    h.flip = L.ptr(flip)
    h.out_img = L.ptr(res["img"])
    h.out_mask = L.ptr(res["mask"])
    h.out_labels = L.ptr(res["labels"])
    h.out_joints = L.ptr(res["joints"]) if P else None
    h.out_count = L.ptr(res["count"])
    h.status = L.ptr(res["status"])
    h.sigma = float(sigma) if sigma is not None else 0.0
    h.thre = float(thre) if thre is not None else 0.0
    h.out_vec_label, h.out_heat_label = L.ptr(res["y1"]), L.ptr(res["y2"])
    h.out_vec_weights, h.out_heat_weights = L.ptr(res["x1"]), L.ptr(res["x2"])
    L.check(lib.rmpe_gt_batch_host(C.byref(h)))
    return res


def heatmaps_host(joints, n_persons, mask, *, f64=True, want_count=False, sigma=None, thre=None, paf_average=False):
    """Heatmapper(sigma, thre).create_heatmaps for a batch: joints (B,P,18,3) already in output coordinates,
    mask (B,46,46) in [0,1] (same dtype as the labels)."""
    lib = L.ensure_init()
    ft = np.float64 if f64 else np.float32
    mask = np.array(mask, dtype=ft, order="C", copy=True)
    B = mask.shape[0]
    joints = L.c_contig(joints, np.float64)
    P = joints.shape[1]
    n_persons = L.c_contig(n_persons, np.int32).reshape(B)
    labels = np.empty((B, NL, GRID, GRID), ft)
    count = np.empty((B, 19, GRID, GRID), np.int32) if want_count else None
    status = np.zeros(B, np.int32)
    h = L.GtBatchHost()
    h.batch = B
    h.max_persons = P
    h.flags = L.GT_NO_TRANSFORM | (L.GT_LABELS_F64 if f64 else 0) | (L.GT_PAF_AVERAGE if paf_average else 0)
    h.sigma = float(sigma) if sigma is not None else 0.0
    h.thre = float(thre) if thre is not None else 0.0
    h.joints = L.ptr(joints) if P else None
    h.n_persons = L.ptr(n_persons)
    h.out_mask = L.ptr(mask)
    h.out_labels = L.ptr(labels)
    h.out_count = L.ptr(count)
    h.status = L.ptr(status)
    L.check(lib.rmpe_gt_batch_host(C.byref(h)))
    return dict(labels=labels, count=count, status=status)


def keras_batch_host(labels, mask):
    """DataIteratorBase.gen's batch assembly (training/ds_generators.py:47-77) for a whole batch:
    labels (B,57,46,46), mask (B,46,46) of one float dtype -> dict(x1 (B,46,46,38) mask repeated,
    x2 (B,46,46,19), y1 (B,46,46,38) = labels[:, :38] as NHWC, y2 (B,46,46,19))."""
    lib = L.ensure_init()
    labels = np.ascontiguousarray(labels)
    f64 = labels.dtype == np.float64
    ft = np.float64 if f64 else np.float32
    labels = L.c_contig(labels, ft)
    mask = L.c_contig(mask, ft)
    B = labels.shape[0]
    assert labels.shape == (B, NL, GRID, GRID) and mask.shape == (B, GRID, GRID)
    out = dict(x1=np.empty((B, GRID, GRID, 38), ft), x2=np.empty((B, GRID, GRID, 19), ft),
               y1=np.empty((B, GRID, GRID, 38), ft), y2=np.empty((B, GRID, GRID, 19), ft))
    k = L.KerasBatch()
    k.batch, k.flags = B, (L.GT_LABELS_F64 if f64 else 0)
    k.labels, k.mask = L.ptr(labels), L.ptr(mask)
    k.vec_weights, k.heat_weights, k.vec_label, k.heat_label = (L.ptr(out["x1"]), L.ptr(out["x2"]), L.ptr(out["y1"]),
                                                                L.ptr(out["y2"]))
    L.check(lib.rmpe_keras_batch_host(C.byref(k)))
    return out


# ------------------------------------------------------------------------------------------
# GT, device-resident (torch tensors)
# ------------------------------------------------------------------------------------------
class GtDevicePlan:
    """Preallocated device buffers for batches of fixed geometry; `run()` enqueues the kernels on
    the current torch stream and returns immediately."""

    def __init__(self, batch, max_persons, src_hw=(368, 368), f64=False, chw=False, want_count=False,
                 device=None, keras=False, sigma=None, thre=None):
        import torch
        self.torch = torch
        self.lib = L.ensure_init(device)
        dev = torch.device("cuda", L._inited_device)
        self.device = dev
        self.B, self.P = batch, max_persons
        H, W = src_hw
        self.H, self.W = H, W
        self.f64, self.chw = f64, chw
        ft = torch.float64 if f64 else torch.float32
        pad = 64  # slack behind the sources: bulk row copies read whole 16-byte groups
        self.src_img = torch.zeros(batch * H * W * 3 + pad, dtype=torch.uint8, device=dev)
        self.src_mask = torch.zeros(batch * H * W + pad, dtype=torch.uint8, device=dev)
        desc = np.zeros(batch, dtype=L.SRC_DESC_DTYPE)
        desc["img_offset"] = np.arange(batch, dtype=np.int64) * (H * W * 3)
        desc["mask_offset"] = np.arange(batch, dtype=np.int64) * (H * W)
        desc["height"], desc["width"] = H, W
        desc["img_pitch"], desc["mask_pitch"] = 3 * W, W
        self.desc = torch.from_numpy(desc.view(np.uint8).reshape(batch, -1).copy()).to(dev)
        self.joints = torch.zeros((batch, max(max_persons, 1), 18, 3), dtype=torch.float64, device=dev)
        self.n_persons = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.M = torch.zeros((batch, 6), dtype=torch.float64, device=dev)
        self.flip = torch.zeros(batch, dtype=torch.uint8, device=dev)
        self.out_img = torch.empty((batch, 3, G.height, G.width) if chw else (batch, G.height, G.width, 3),
                                   dtype=torch.uint8, device=dev)
        self.out_mask = torch.empty((batch, GRID, GRID), dtype=ft, device=dev)
        self.out_labels = torch.empty((batch, NL, GRID, GRID), dtype=ft, device=dev)
        self.out_joints = torch.empty_like(self.joints)
        self.out_count = torch.empty((batch, 19, GRID, GRID), dtype=torch.int32, device=dev) if want_count else None
        self.status = torch.zeros(batch, dtype=torch.int32, device=dev)
        d = L.GtBatch()
        d.batch, d.max_persons = batch, max_persons
        d.flags = (L.GT_LABELS_F64 if f64 else 0) | (L.GT_IMG_CHW if chw else 0)
        d.src_img, d.src_mask, d.src_desc = L.ptr(self.src_img), L.ptr(self.src_mask), L.ptr(self.desc)
        d.joints, d.n_persons, d.M, d.flip = L.ptr(self.joints), L.ptr(self.n_persons), L.ptr(self.M), L.ptr(self.flip)
        d.out_img, d.out_mask, d.out_labels = L.ptr(self.out_img), L.ptr(self.out_mask), L.ptr(self.out_labels)
        d.out_joints = L.ptr(self.out_joints)
        d.out_count = L.ptr(self.out_count)
        d.status = L.ptr(self.status)
        d.sigma = float(sigma) if sigma is not None else 0.0
        d.thre = float(thre) if thre is not None else 0.0
        self.desc_struct = d
        self.keras = keras
        if keras:   # Keras-ready NHWC tensors (DataIteratorBase.gen) written by the rasteriser itself
            self.x1 = torch.empty((batch, GRID, GRID, 38), dtype=ft, device=dev)
            self.x2 = torch.empty((batch, GRID, GRID, 19), dtype=ft, device=dev)
            self.y1 = torch.empty((batch, GRID, GRID, 38), dtype=ft, device=dev)
            self.y2 = torch.empty((batch, GRID, GRID, 19), dtype=ft, device=dev)
            d.out_vec_label, d.out_heat_label = L.ptr(self.y1), L.ptr(self.y2)
            d.out_vec_weights, d.out_heat_weights = L.ptr(self.x1), L.ptr(self.x2)
            if keras == "only":       # no planar labels at all
                d.out_labels = None

    def upload(self, imgs, masks, joints, n_persons, M, flip, non_blocking=False):
        t = self.torch
        B, H, W = self.B, self.H, self.W
        self.src_img[:B * H * W * 3].copy_(t.as_tensor(imgs).reshape(-1), non_blocking=non_blocking)
        self.src_mask[:B * H * W].copy_(t.as_tensor(masks).reshape(-1), non_blocking=non_blocking)
        if self.P:
            self.joints.copy_(t.as_tensor(joints).reshape(self.joints.shape), non_blocking=non_blocking)
        self.n_persons.copy_(t.as_tensor(n_persons), non_blocking=non_blocking)
        self.M.copy_(t.as_tensor(M).reshape(B, 6), non_blocking=non_blocking)
        self.flip.copy_(t.as_tensor(flip), non_blocking=non_blocking)

    def run(self, simple=False, stream=None):
        d = self.desc_struct
        base = d.flags & ~L.GT_SIMPLE_KERNELS
        d.flags = base | (L.GT_SIMPLE_KERNELS if simple else 0)
        if stream is None:
            stream = self.torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.rmpe_gt_batch(C.byref(d), C.c_void_p(stream)))


# ------------------------------------------------------------------------------------------
# decode
# ------------------------------------------------------------------------------------------
def make_frames(frames):
    """frames: list of dict(H, W, scales=[(paf (h,w,38) f32, heat (h,w,19) f32, pad_down, pad_right), ...]).
    Returns (desc structured array, heat flat f32, paf flat f32)."""
    desc = np.zeros(len(frames), dtype=L.FRAME_DESC_DTYPE)
    heats, pafs = [], []
    ho = po = 0
    for i, f in enumerate(frames):
        desc[i]["height"], desc[i]["width"] = f["H"], f["W"]
        sc = f["scales"]
        assert 1 <= len(sc) <= L.MAX_SCALES
        desc[i]["n_scales"] = len(sc)
        for s, (paf, heat, pd, pr) in enumerate(sc):
            paf = np.ascontiguousarray(paf, np.float32)
            heat = np.ascontiguousarray(heat, np.float32)
            assert paf.ndim == 3 and paf.shape[2] == 38 and heat.shape[2] == 19 and paf.shape[:2] == heat.shape[:2]
            desc[i]["grid_h"][s], desc[i]["grid_w"][s] = heat.shape[0], heat.shape[1]
            desc[i]["pad_down"][s], desc[i]["pad_right"][s] = pd, pr
            desc[i]["heat_offset"][s], desc[i]["paf_offset"][s] = ho, po
            heats.append(heat.reshape(-1))
            pafs.append(paf.reshape(-1))
            ho += heat.size
            po += paf.size
    return desc, np.concatenate(heats), np.concatenate(pafs)


def _unpack_decode(B, MP, MC, MS, cand, npk, conn, nconn, lc, nlc, sub, nsub, status):
    out = []
    for i in range(B):
        n = int(npk[i].sum())
        r = dict(candidate=cand[i, :n].copy(), subset=sub[i, :int(nsub[i])].copy(),
                 n_peaks=npk[i].copy(), status=int(status[i]))
        r["connections"] = [None if nconn[i, k] < 0 else conn[i, k, :nconn[i, k]].copy() for k in range(19)]
        r["special_k"] = [k for k in range(19) if nconn[i, k] < 0]
        if lc is not None:
            r["limb_candidates"] = [lc[i, k, :nlc[i, k]].copy() for k in range(19)]
        out.append(r)
    return out


def decode_batch_host(frames, thre1=0.1, thre2=0.05, stride=8, max_peaks=128, max_cand=1024,
                      max_persons=64, want_limb_candidates=False):
    """Decode a list of frames (see make_frames) from host blobs; returns a list of dicts with
    candidate (N,4) f64, subset (M,20) f64, connections, special_k, status."""
    desc, heat, paf = make_frames(frames)
    return decode_batch_host_raw(desc, heat, paf, thre1, thre2, stride, max_peaks, max_cand, max_persons,
                                 want_limb_candidates)


def decode_batch_host_raw(desc, heat, paf, thre1=0.1, thre2=0.05, stride=8, max_peaks=128, max_cand=1024,
                          max_persons=64, want_limb_candidates=False):
    """decode_batch_host on an already flattened batch: desc (FRAME_DESC_DTYPE array), heat / paf flat f32."""
    lib = L.ensure_init()
    B = len(desc)
    MP, MC, MS = max_peaks, max_cand, max_persons
    cand = np.zeros((B, 18 * MP, 4), np.float64)
    npk = np.zeros((B, 18), np.int32)
    conn = np.zeros((B, 19, MP, 5), np.float64)
    nconn = np.zeros((B, 19), np.int32)
    lc = np.zeros((B, 19, MC, 4), np.float64) if want_limb_candidates else None
    nlc = np.zeros((B, 19), np.int32)
    sub = np.zeros((B, MS, 20), np.float64)
    nsub = np.zeros(B, np.int32)
    status = np.zeros(B, np.int32)
    h = L.DecodeBatchHost()
    h.batch, h.max_peaks, h.max_cand, h.max_persons, h.stride, h.flags = B, MP, MC, MS, stride, 0
    h.thre1, h.thre2 = float(thre1), float(thre2)
    h.heat, h.paf, h.heat_elems, h.paf_elems = L.ptr(heat), L.ptr(paf), heat.size, paf.size
    h.frames = L.ptr(desc)
    h.candidate, h.n_peaks, h.connections, h.n_conn = L.ptr(cand), L.ptr(npk), L.ptr(conn), L.ptr(nconn)
    h.limb_cand, h.n_limb_cand = L.ptr(lc), L.ptr(nlc)
    h.subset, h.n_subset, h.status = L.ptr(sub), L.ptr(nsub), L.ptr(status)
    L.check(lib.rmpe_decode_batch_host(C.byref(h)))
    return _unpack_decode(B, MP, MC, MS, cand, npk, conn, nconn, lc, nlc, sub, nsub, status)


class DecodeHostPlan:
    """rmpe_decode_batch_host over a fixed list of frames with every host buffer allocated once (pinned when torch is
    importable): what a serving loop that re-uses its buffers calls, and what bench.py's decode `e2e` times.
    `run()` = one C call: blobs host->device, the whole decode pipeline, the filled prefix of every table device->host."""

    def __init__(self, frames, thre1=0.1, thre2=0.05, stride=8, max_peaks=128, max_cand=1024, max_persons=64, pinned=True):
        self.lib = L.ensure_init()
        desc, heat, paf = make_frames(frames)
        B = self.B = len(desc)
        self.MP, self.MC, self.MS = max_peaks, max_cand, max_persons
        self._keep = []

        def alloc(shape, dtype):
            if pinned:
                import torch
                t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
                self._keep.append(t)
                return t.numpy()
            return np.empty(shape, dtype)

        self.heat, self.paf = alloc(heat.shape, np.float32), alloc(paf.shape, np.float32)
        self.heat[...] = heat
        self.paf[...] = paf
        self.desc = desc
        self.cand = alloc((B, 18 * max_peaks, 4), np.float64)
        self.npk = alloc((B, 18), np.int32)
        self.conn = alloc((B, 19, max_peaks, 5), np.float64)
        self.nconn = alloc((B, 19), np.int32)
        self.nlc = alloc((B, 19), np.int32)
        self.sub = alloc((B, max_persons, 20), np.float64)
        self.nsub = alloc((B,), np.int32)
        self.status = alloc((B,), np.int32)
        h = L.DecodeBatchHost()
        h.batch, h.max_peaks, h.max_cand, h.max_persons, h.stride, h.flags = B, max_peaks, max_cand, max_persons, stride, 0
        h.thre1, h.thre2 = float(thre1), float(thre2)
        h.heat, h.paf, h.heat_elems, h.paf_elems = L.ptr(self.heat), L.ptr(self.paf), heat.size, paf.size
        h.frames = L.ptr(desc)
        h.candidate, h.n_peaks, h.connections, h.n_conn = L.ptr(self.cand), L.ptr(self.npk), L.ptr(self.conn), L.ptr(self.nconn)
        h.limb_cand, h.n_limb_cand = None, L.ptr(self.nlc)
        h.subset, h.n_subset, h.status = L.ptr(self.sub), L.ptr(self.nsub), L.ptr(self.status)
        self.h = h
        self.h2d_bytes = int(heat.nbytes + paf.nbytes + desc.nbytes)

    def run(self):
        L.check(self.lib.rmpe_decode_batch_host(C.byref(self.h)))

    def d2h_bytes(self):
        """Bytes the last run() brought back: the count vectors and the filled prefix of every table (rmpe_host.cu)."""
        B, MP, MS = self.B, self.MP, self.MS
        rows = int(np.minimum(self.npk, MP).sum(axis=1).max()) if B else 0
        mconn = int(np.clip(self.nconn, 0, MP).max()) if B else 0
        msub = int(np.minimum(self.nsub, MS).max()) if B else 0
        return int(B * (18 + 19 + 1 + 1) * 4 + B * rows * 32 + B * 19 * mconn * 40 + B * msub * 160)

    def results(self):
        return _unpack_decode(self.B, self.MP, self.MC, self.MS, self.cand, self.npk, self.conn, self.nconn, None,
                              self.nlc, self.sub, self.nsub, self.status)


class DecodeDevicePlan:
    """Device-resident decode of a fixed list of frames: blobs live in HBM, `run()` enqueues the
    whole pipeline on the current torch stream, `results()` reads back and unpacks."""

    def __init__(self, frames, thre1=0.1, thre2=0.05, stride=8, max_peaks=128, max_cand=1024,
                 max_persons=64, want_limb_candidates=False, device=None, workspace_bytes=None):
        import torch
        self.torch = torch
        self.lib = L.ensure_init(device)
        dev = torch.device("cuda", L._inited_device)
        self.device = dev
        desc, heat, paf = make_frames(frames)
        self.desc_host = desc
        B = self.B = len(frames)
        self.MP, self.MC, self.MS = max_peaks, max_cand, max_persons
        self.heat = torch.from_numpy(heat).to(dev)
        self.paf = torch.from_numpy(paf).to(dev)
        self.heat_host_bytes = heat.nbytes
        self.paf_host_bytes = paf.nbytes
        self.frames_dev = torch.from_numpy(desc.view(np.uint8).reshape(B, -1).copy()).to(dev)
        f64 = torch.float64
        self.cand = torch.zeros((B, 18 * max_peaks, 4), dtype=f64, device=dev)
        self.npk = torch.zeros((B, 18), dtype=torch.int32, device=dev)
        self.conn = torch.zeros((B, 19, max_peaks, 5), dtype=f64, device=dev)
        self.nconn = torch.zeros((B, 19), dtype=torch.int32, device=dev)
        self.lc = torch.zeros((B, 19, max_cand, 4), dtype=f64, device=dev) if want_limb_candidates else None
        self.nlc = torch.zeros((B, 19), dtype=torch.int32, device=dev)
        self.sub = torch.zeros((B, max_persons, 20), dtype=f64, device=dev)
        self.nsub = torch.zeros(B, dtype=torch.int32, device=dev)
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        need = int(self.lib.rmpe_decode_workspace_bytes(B, L.ptr(desc), max_peaks, max_cand, stride))
        if workspace_bytes is None:
            workspace_bytes = need
        self.workspace = torch.empty(int(workspace_bytes), dtype=torch.uint8, device=dev)
        d = L.DecodeBatch()
        d.batch, d.max_peaks, d.max_cand, d.max_persons, d.stride, d.flags = B, max_peaks, max_cand, max_persons, stride, 0
        d.thre1, d.thre2 = float(thre1), float(thre2)
        d.heat, d.paf, d.frames, d.frames_host = L.ptr(self.heat), L.ptr(self.paf), L.ptr(self.frames_dev), L.ptr(desc)
        d.candidate, d.n_peaks, d.connections, d.n_conn = L.ptr(self.cand), L.ptr(self.npk), L.ptr(self.conn), L.ptr(self.nconn)
        d.limb_cand, d.n_limb_cand = L.ptr(self.lc), L.ptr(self.nlc)
        d.subset, d.n_subset, d.status = L.ptr(self.sub), L.ptr(self.nsub), L.ptr(self.status)
        d.workspace, d.workspace_bytes = L.ptr(self.workspace), int(workspace_bytes)
        self.d = d

    def run(self, stream=None):
        if stream is None:
            stream = self.torch.cuda.current_stream(self.device).cuda_stream
        L.check(self.lib.rmpe_decode_batch(C.byref(self.d), C.c_void_p(stream)))
        # the operator tables in the workspace depend on the frame geometry only: later runs of this plan reuse them
        self.d.flags |= L.DECODE_REUSE_TABLES

    def results(self):
        self.torch.cuda.synchronize(self.device)
        g = lambda t: None if t is None else t.cpu().numpy()
        return _unpack_decode(self.B, self.MP, self.MC, self.MS, g(self.cand), g(self.npk), g(self.conn),
                              g(self.nconn), g(self.lc), g(self.nlc), g(self.sub), g(self.nsub), g(self.status))
