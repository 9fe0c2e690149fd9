"""GPU parity of the inference decode, through the C ABI.  Bars: peak indices, limb-candidate
sets, connections and person assignments bit-exact; the up-sampled / smoothed maps bit-exact too
(the cv2.resize and scipy gaussian arithmetic is reproduced operation by operation)."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from cases import DECODE_CASES, decode_case_inputs, frames_of
from oracle import decode_oracle as do

pytestmark = pytest.mark.gpu


def _oracle(case, detail=True):
    name, H, W, P, seed, multi = case
    blobs = decode_case_inputs(case)
    if multi:
        return do.multi_scale(blobs, H, W, detail=detail)
    return do.single_scale(blobs[0][0], blobs[0][1], H, W, detail=detail)


@pytest.mark.parametrize("case", DECODE_CASES, ids=[c[0] for c in DECODE_CASES])
def test_decode_matches_reference_golden(rmpe, decode_golden, case):
    name = case[0]
    r = rmpe.batch.decode_batch_host([frames_of(case)], max_peaks=128, max_cand=1024, max_persons=64)[0]
    assert r["status"] == 0
    assert np.array_equal(r["candidate"], decode_golden[name + "_candidate"])
    assert np.array_equal(r["subset"], decode_golden[name + "_subset"])


@pytest.mark.parametrize("case", [DECODE_CASES[2], DECODE_CASES[3], DECODE_CASES[5]], ids=["d2", "d3", "d5"])
def test_decode_stages_match_oracle(rmpe, case):
    """limb-candidate sets (generation order), connections and special_k, not just the end result."""
    o = _oracle(case)
    r = rmpe.batch.decode_batch_host([frames_of(case)], want_limb_candidates=True)[0]
    assert r["special_k"] == o["special_k"]
    for k in range(19):
        oc = np.array(o["limb_candidates"][k], dtype=np.float64).reshape(-1, 4)
        assert np.array_equal(r["limb_candidates"][k], oc), "limb %d candidates" % k
        if k in o["special_k"]:
            assert r["connections"][k] is None
        else:
            assert np.array_equal(r["connections"][k], np.asarray(o["connection_all"][k]).reshape(-1, 5)), "limb %d" % k
    assert np.array_equal(r["candidate"], o["candidate"]) and np.array_equal(r["subset"], o["subset"])


@pytest.mark.parametrize("case", [DECODE_CASES[2], DECODE_CASES[3], DECODE_CASES[5]], ids=["d2", "d3", "d5"])
def test_heat_maps_bit_exact(rmpe, case):
    """D1 (cv2.resize chain / 4-scale f64 average) and D2 (scipy gaussian) as materialised maps."""
    import torch
    name, H, W, P, seed, multi = case
    o = _oracle(case)
    desc, heat, paf = rmpe.batch.make_frames([frames_of(case)])
    dev = torch.device("cuda", 0)
    dt = torch.float64 if multi else torch.float32
    up = torch.zeros((18, H, W), dtype=dt, device=dev)
    sm = torch.zeros((18, H, W), dtype=dt, device=dev)
    heat_d = torch.from_numpy(heat).to(dev)
    lib = rmpe.lib.load()
    rmpe.lib.check(lib.rmpe_debug_heat_maps(desc.ctypes.data, heat_d.data_ptr(), up.data_ptr(), sm.data_ptr(), None))
    torch.cuda.synchronize()
    want_up = np.transpose(o["heat_up"][:, :, :18], (2, 0, 1))
    got_up = up.cpu().numpy()
    assert got_up.dtype == want_up.dtype
    assert np.array_equal(got_up, want_up), "up-sampled heat: %d of %d differ, max %g" % (
        (got_up != want_up).sum(), want_up.size, np.abs(got_up - want_up).max())
    want_sm = np.stack([do.gaussian_filter_sigma3(o["heat_up"][:, :, p]) for p in range(18)])
    got_sm = sm.cpu().numpy()
    assert np.array_equal(got_sm, want_sm), "smoothed heat: %d differ, max %g" % (
        (got_sm != want_sm).sum(), np.abs(got_sm - want_sm).max())
    # PAF sampled on the fly == the materialised up-sampled PAF at random integer points
    rng = np.random.RandomState(0)
    n = 4000
    cyx = np.stack([rng.randint(0, 38, n), rng.randint(0, H, n), rng.randint(0, W, n)], axis=1).astype(np.int32)
    cyx[:200, 2] = W - 1          # row tails
    cyx[200:400, 1] = H - 1
    cyx[400:500, 1:] = 0
    cyx_d = torch.from_numpy(cyx).to(dev)
    paf_d = torch.from_numpy(paf).to(dev)
    out = torch.zeros(n, dtype=torch.float64, device=dev)
    rmpe.lib.check(lib.rmpe_debug_paf_points(desc.ctypes.data, paf_d.data_ptr(), n, cyx_d.data_ptr(), out.data_ptr(), None))
    torch.cuda.synchronize()
    want = o["paf_up"][cyx[:, 1], cyx[:, 2], cyx[:, 0]].astype(np.float64)
    assert np.array_equal(out.cpu().numpy(), want)


def test_mixed_batch_and_device_plan(rmpe, decode_golden):
    """One call over frames of different sizes, modes and person counts; device-resident plan."""
    cases = [DECODE_CASES[2], DECODE_CASES[4], DECODE_CASES[6], DECODE_CASES[3], DECODE_CASES[2]]
    frames = [frames_of(c) for c in cases]
    plan = rmpe.batch.DecodeDevicePlan(frames)
    plan.run()
    res = plan.results()
    for c, r in zip(cases, res):
        assert r["status"] == 0
        assert np.array_equal(r["candidate"], decode_golden[c[0] + "_candidate"]), c[0]
        assert np.array_equal(r["subset"], decode_golden[c[0] + "_subset"]), c[0]
    # idempotent: a second run over the same buffers gives the same answer
    plan.run()
    res2 = plan.results()
    for a, b in zip(res, res2):
        assert np.array_equal(a["candidate"], b["candidate"]) and np.array_equal(a["subset"], b["subset"])
    # small workspace forces chunking frame by frame
    need1 = max(int(rmpe.lib.load().rmpe_decode_workspace_bytes(1, rmpe.batch.make_frames([f])[0].ctypes.data, 128, 1024))
                for f in frames)
    need_all = int(rmpe.lib.load().rmpe_decode_workspace_bytes(len(frames), plan.desc_host.ctypes.data, 128, 1024))
    plan3 = rmpe.batch.DecodeDevicePlan(frames, workspace_bytes=min(need_all, need1 + (4 << 20)))
    plan3.run()
    for a, b in zip(res, plan3.results()):
        assert np.array_equal(a["candidate"], b["candidate"]) and np.array_equal(a["subset"], b["subset"])


def test_empty_and_capacity(rmpe):
    H, W = 96, 120
    paf = np.zeros((12, 15, 38), np.float32)
    heat = np.zeros((12, 15, 19), np.float32)
    r = rmpe.batch.decode_batch_host([dict(H=H, W=W, scales=[(paf, heat, 0, 0)])])[0]
    assert r["candidate"].shape == (0, 4) and r["subset"].shape == (0, 20) and len(r["special_k"]) == 19
    # a saturated plateau produces more peaks than the capacity: flagged, never silently wrong
    heat2 = np.ones((12, 15, 19), np.float32)
    r = rmpe.batch.decode_batch_host([dict(H=H, W=W, scales=[(paf, heat2, 0, 0)])], max_peaks=16)[0]
    assert r["status"] & 0x2
    with pytest.raises(OverflowError):
        rmpe.decode._raise_status(r["status"])


def test_process_single_and_multi_scale_drop_in(rmpe, decode_golden):
    """The reference-facing signatures: image path + model.predict stand-in + params dicts."""
    import cv2
    params = {'scale_search': [.5, 1, 1.5, 2], 'thre1': .1, 'thre2': .05}
    mparams = {'boxsize': 368, 'stride': 8, 'padValue': 128}
    tmp = tempfile.mkdtemp()

    class FakeModel:
        def __init__(self, blobs):
            self.blobs, self.i = blobs, 0

        def predict(self, x):
            b = self.blobs[self.i]
            self.i += 1
            assert x.shape[1] // 8 == b[1].shape[0] or x.shape[1] >= b[1].shape[0] * 8
            return [b[0][None], b[1][None]]

    for case in (DECODE_CASES[2], DECODE_CASES[4]):
        name, H, W, P, seed, multi = case
        path = os.path.join(tmp, name + ".png")
        cv2.imwrite(path, np.zeros((H, W, 3), np.uint8))
        blobs = decode_case_inputs(case)
        fn = rmpe.decode.process_multi_scale if multi else rmpe.decode.process_single_scale
        canvas, cand, sub = fn(path, FakeModel(blobs), dict(params), mparams)
        assert canvas.shape == (H, W, 3)
        assert np.array_equal(cand, decode_golden[name + "_candidate"])
        assert np.array_equal(sub, decode_golden[name + "_subset"])


def test_pad_right_down_corner(rmpe):
    img = np.random.RandomState(0).randint(0, 256, size=(45, 61, 3)).astype(np.uint8)
    out, pad = rmpe.util.padRightDownCorner(img, 8, 128)
    want, wpad = do.pad_right_down_corner(img, 8, 128)
    assert pad == wpad and np.array_equal(out, want)


def test_full_size_properties_ski_batch(rmpe):
    """BASELINE config 3 shape (674x712 frames, 84x89 blobs): per-frame results are independent
    of batch position, ids are consecutive, peaks sorted row-major per part; one frame is checked
    against the oracle."""
    H, W = 674, 712
    h, w = rmpe.synth.single_scale_grid(H, W)
    frames = []
    for i in range(8):
        paf, heat, _ = rmpe.synth.decode_blobs(500 + (i % 4), (H, W), (h, w), 3 + (i % 4))
        frames.append(dict(H=H, W=W, scales=[(paf, heat, 0, 0)]))
    res = rmpe.batch.decode_batch_host(frames)
    for i in range(4):
        assert np.array_equal(res[i]["candidate"], res[i + 4]["candidate"])
        assert np.array_equal(res[i]["subset"], res[i + 4]["subset"])
    for r in res:
        c = r["candidate"]
        assert np.array_equal(c[:, 3], np.arange(len(c)))
        off = 0
        for p in range(18):
            blk = c[off:off + r["n_peaks"][p]]
            key = blk[:, 1] * W + blk[:, 0]
            assert np.all(np.diff(key) > 0)
            off += r["n_peaks"][p]
        assert r["status"] == 0 and len(r["subset"]) >= 1
    cand, sub = do.single_scale(frames[1]["scales"][0][0], frames[1]["scales"][0][1], H, W)
    assert np.array_equal(res[1]["candidate"], cand) and np.array_equal(res[1]["subset"], sub)


def test_mixed_screened_and_materialised_frames(rmpe):
    """One batch holding frames of both heat paths: the float32 screening path and -- for frames whose
    up-sampling factor is too small for its tiles (240x320 at scale 2) -- the materialised exact path."""
    cases = [("m0", 240, 320, 3, 21, True), ("m1", 427, 640, 4, 22, True), ("m2", 96, 120, 2, 23, False),
             ("m3", 480, 640, 2, 24, True), ("m4", 240, 320, 2, 25, True)]
    frames = [frames_of(c) for c in cases]
    res = rmpe.batch.decode_batch_host(frames)
    for c, r in zip(cases, res):
        o = _oracle(c, detail=False)
        assert r["status"] == 0
        assert np.array_equal(r["candidate"], o[0]), c[0]
        assert np.array_equal(r["subset"], o[1]), c[0]


@pytest.mark.parametrize("seed", range(8))
def test_random_frame_shapes(rmpe, seed):
    """Odd frame sizes (tiles partly outside the frame, frames smaller than the Gaussian radius of a tile
    ring, reflect/replicate borders inside the screening operators), single and multi scale, against the oracle."""
    rng = np.random.RandomState(9000 + seed)
    H, W = int(rng.randint(40, 420)), int(rng.randint(40, 520))
    if seed == 0:
        H, W = 33, 47            # smaller than one screening tile in both directions
    if seed == 1:
        H, W = 127, 253          # interior width of a tile + 1
    P = int(rng.randint(1, 5))
    multi = bool(seed % 2)
    case = ("r%d" % seed, H, W, P, 500 + seed, multi)
    o = _oracle(case, detail=False)
    r = rmpe.batch.decode_batch_host([frames_of(case)])[0]
    assert r["status"] == 0
    assert np.array_equal(r["candidate"], o[0]), "%s: %d vs %d peaks" % (case, len(r["candidate"]), len(o[0]))
    assert np.array_equal(r["subset"], o[1])


def test_plan_of_many_shapes_rerun(rmpe):
    """A device plan over more frames than one chunk (64) with a different shape per frame, run twice: the second run
    asks for table reuse, which only holds for single-chunk batches (every chunk rebuilds its operator tables in the
    same workspace region) -- found with the 1000-frame COCO-val-shaped run of BASELINE configs[3]."""
    rng = np.random.RandomState(77)
    cases = [("m%d" % k, int(rng.randint(48, 200)), int(rng.randint(48, 260)), int(rng.randint(1, 4)), 31000 + k,
              bool(k % 2)) for k in range(80)]
    frames = [frames_of(c) for c in cases]
    plan = rmpe.batch.DecodeDevicePlan(frames)
    plan.run()
    res = plan.results()
    plan.run()
    plan.run()
    res2 = plan.results()
    for c, a, b in zip(cases, res, res2):
        assert a["status"] == 0 and b["status"] == 0, c
        assert np.array_equal(a["candidate"], b["candidate"]) and np.array_equal(a["subset"], b["subset"]), c
    for i in (0, 17, 63, 79):
        o = _oracle(cases[i], detail=False)
        assert np.array_equal(res2[i]["candidate"], o[0]) and np.array_equal(res2[i]["subset"], o[1]), cases[i]


def test_full_size_decode_batch_properties(rmpe):
    """BASELINE config 3 size on one device (64 ski-shaped frames in one call): a few frames equal the oracle,
    size-independent properties hold for all, and a second run is bit-identical (atomics only feed sorted lists)."""
    H, W = 674, 712
    h, w = rmpe.synth.single_scale_grid(H, W)
    frames = []
    for i in range(64):
        paf, heat, _ = rmpe.synth.decode_blobs(4000 + i, (H, W), (h, w), 3 + (i % 3))
        frames.append(dict(H=H, W=W, scales=[(paf, heat, 0, 0)]))
    plan = rmpe.batch.DecodeDevicePlan(frames)
    plan.run()
    res = plan.results()
    plan.run()
    res2 = plan.results()
    for i, (a, b) in enumerate(zip(res, res2)):
        assert a["status"] == 0
        assert np.array_equal(a["candidate"], b["candidate"]) and np.array_equal(a["subset"], b["subset"])
        c, s = a["candidate"], a["subset"]
        assert np.array_equal(c[:, 3], np.arange(len(c)))                     # ids consecutive across parts
        assert int(a["n_peaks"].sum()) == len(c)
        off = 0
        for n in a["n_peaks"]:
            part = c[off:off + n]
            assert np.array_equal(part, part[np.lexsort((part[:, 0], part[:, 1]))])
            off += n
        ids = s[:, :18]
        assert ((ids == -1) | ((ids >= 0) & (ids < len(c)))).all()
        assert (s[:, 19] >= (ids >= 0).sum(axis=1)).all()                     # count column (the reference may count a re-assigned part twice)
        assert len(s) >= 3                                                    # the synthetic persons are found
    for i in (0, 31, 63):
        f = frames[i]["scales"][0]
        cand, sub = do.single_scale(f[0], f[1], H, W)
        assert np.array_equal(res[i]["candidate"], cand) and np.array_equal(res[i]["subset"], sub)
