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
    need1 = max(int(rmpe.lib.load().rmpe_decode_workspace_bytes(1, rmpe.batch.make_frames([f])[0].ctypes.data, 128, 1024, 8))
                for f in frames)
    need_all = int(rmpe.lib.load().rmpe_decode_workspace_bytes(len(frames), plan.desc_host.ctypes.data, 128, 1024, 8))
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


def _check_against_oracle(r, o):
    assert r["status"] == 0
    assert np.array_equal(r["candidate"], o["candidate"]), "peaks"
    assert r["special_k"] == o["special_k"]
    for k in range(19):
        oc = np.array(o["limb_candidates"][k], dtype=np.float64).reshape(-1, 4)
        assert np.array_equal(r["limb_candidates"][k], oc), "limb %d candidates" % k
        if k not in o["special_k"]:
            assert np.array_equal(r["connections"][k], np.asarray(o["connection_all"][k]).reshape(-1, 5)), "limb %d" % k
    assert np.array_equal(r["subset"], o["subset"]), "persons"


@pytest.mark.parametrize("thre1,thre2", [(0.05, 0.05), (0.2, 0.05), (0.1, 0.01), (0.1, 0.1), (0.3, 0.2)])
@pytest.mark.parametrize("case", [DECODE_CASES[2], DECODE_CASES[3]], ids=["d2", "d3"])
def test_decode_thresholds(rmpe, case, thre1, thre2):
    """params['thre1'] / ['thre2'] other than the defaults: the screening bound of k_screen_plan and the exact test of
    k_peak_verify both depend on thre1, the limb criterion on thre2 (eval...:101-111, :150-157)."""
    name, H, W, P, seed, multi = case
    blobs = decode_case_inputs(case)
    paf, heat = blobs[0][0], blobs[0][1].copy()
    # weak bumps (smoothed peak heights between the thresholds under test) so that thre1 decides which of them are peaks
    h, w = heat.shape[:2]
    yy, xx = np.mgrid[0:h, 0:w]
    amps = (0.07, 0.09, 0.12, 0.16, 0.22, 0.26, 0.35, 0.45)
    for n, amp in enumerate(amps):
        cy, cx = 3 + (h - 6) * n // len(amps), (w - 4 - 5 * n) % w
        heat[:, :, (2 * n + 1) % 18] += (amp * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 2.0)).astype(np.float32)
    frame = dict(H=H, W=W, scales=[(paf, heat, 0, 0)])
    o = do.single_scale(paf, heat, H, W, thre1=thre1, thre2=thre2, detail=True)
    r = rmpe.batch.decode_batch_host([frame], thre1=thre1, thre2=thre2, want_limb_candidates=True)[0]
    _check_against_oracle(r, o)
    n_default = len(do.single_scale(paf, heat, H, W)[0])
    if thre1 < 0.1:
        assert len(o["candidate"]) > n_default      # the threshold matters here
    if thre1 > 0.1:
        assert len(o["candidate"]) < n_default


def _multi_frame(rmpe, seed, H, W, P, scale_search):
    S = rmpe.synth
    _, _, persons = S.decode_blobs(seed, (H, W), (4, 4), P)
    sc = []
    for (Hs, Ws, pd, pr, hs, ws) in do.multi_scale_feed_shapes(H, W, scale_search):
        paf, heat, _ = S.decode_blobs(seed + 1000 * len(sc), (H, W), (hs, ws), P, persons=persons, stride=8.0 * H / Hs)
        sc.append((paf, heat, pd, pr))
    return sc


@pytest.mark.parametrize("scale_search,thre1", [((0.5, 1), 0.1), ((0.5, 1, 1.5), 0.1), ((1, 2, 1.5), 0.1), ((1, 1.5), 0.2),
                                                 ((0.5, 1, 1.5, 2), 0.05)])
def test_decode_two_and_three_scales(rmpe, scale_search, thre1):
    """process_multi_scale with params['scale_search'] of 2 or 3 entries (RmpeFrameDesc.n_scales 2..4): the average
    divides by len(multiplier) (eval...:91-92)."""
    H, W = 200, 264
    sc = _multi_frame(rmpe, 31, H, W, 2, scale_search)
    o = do.multi_scale(sc, H, W, thre1=thre1, detail=True)
    r = rmpe.batch.decode_batch_host([dict(H=H, W=W, scales=sc)], thre1=thre1, want_limb_candidates=True)[0]
    _check_against_oracle(r, o)
    assert len(o["subset"]) >= 1


@pytest.mark.parametrize("multi", [False, True], ids=["single", "multi"])
def test_dense_blobs_nothing_culled(rmpe, multi):
    """Worst case of the screening: every (tile, part) pair is active, every part holds ~35 peaks, a saturated
    plateau sits in the middle -- the lists must still be the reference's, element by element."""
    H, W = (64, 80) if multi else (96, 120)    # ~30 peaks per part either way (the assembly holds at most 128 rows)
    rng = np.random.RandomState(5)

    def blob(h, w):
        heat = (0.3 + 0.15 * rng.normal(size=(h, w, 19))).astype(np.float32)
        heat[h // 4:h // 2 + 1, w // 4:w // 2 + 2, :] = 0.75      # exact plateau in the blob
        return (0.3 * rng.normal(size=(h, w, 38))).astype(np.float32), heat

    if not multi:
        paf, heat = blob(12, 15)
        sc = [(paf, heat, 0, 0)]
        o = do.single_scale(paf, heat, H, W, detail=True)
    else:
        sc = []
        for (Hs, Ws, pd, pr, hs, ws) in do.multi_scale_feed_shapes(H, W, (0.5, 1)):
            paf, heat = blob(hs, ws)
            sc.append((paf, heat, pd, pr))
        o = do.multi_scale(sc, H, W, detail=True)
    assert min(len(p) for p in o["all_peaks"]) >= 5 and not o["overflow"]
    r = rmpe.batch.decode_batch_host([dict(H=H, W=W, scales=sc)], max_peaks=256, max_cand=4096, max_persons=128,
                                     want_limb_candidates=True)[0]
    _check_against_oracle(r, o)


def test_dense_full_size_frame(rmpe):
    """A ski.jpg-sized frame of dense blobs (bench.py's worst-case leg): every tile of every part is screened, ~25 peaks
    per part, thousands of limb pairs -- peaks, limb candidates, connections and persons equal the oracle's."""
    f = rmpe.synth.dense_frame(29000, 674, 712)
    paf, heat, _, _ = f["scales"][0]
    o = do.single_scale(paf, heat, 674, 712, detail=True)
    assert min(len(p) for p in o["all_peaks"]) >= 10 and not o["overflow"]
    r = rmpe.batch.decode_batch_host([f], max_peaks=128, max_cand=1024, max_persons=128, want_limb_candidates=True)[0]
    _check_against_oracle(r, o)


def test_found_more_than_two_rows_raises_like_the_reference(rmpe):
    """eval...:192-195: a connection whose A peak and B peak sit in three subset rows makes the reference index
    subset_idx[2] -> IndexError.  Lists made by k_limbs are one-to-one and cannot get there, so the assembly kernel is fed
    crafted lists: limb 14 (Reye->Rear) hands ONE ear to two persons, limb 17 (Rsho->Rear) then finds three rows."""
    import torch
    L = rmpe.lib
    lib = L.load()
    MP, MS = 8, 16
    dev = torch.device("cuda", 0)
    n_peaks = np.zeros(18, np.int32)
    cand = np.zeros((18 * MP, 4))
    conn = np.zeros((19, MP, 5))
    n_conn = np.full(19, -1, np.int32)
    # peaks: necks 0,1,2 | Rshos 3,4,5 | noses 6,7 | Reyes 8,9 | Rear 10   (ids consecutive over parts 0..17)
    counts = {0: 2, 1: 3, 2: 3, 14: 2, 16: 1}
    ids, nxt = {}, 0
    for part in range(18):
        n_peaks[part] = counts.get(part, 0)
        ids[part] = list(range(nxt, nxt + n_peaks[part]))
        for i in ids[part]:
            cand[i] = [10 * i, 5 * i, 0.9, i]
        nxt += n_peaks[part]

    def put(k, rows):
        n_conn[k] = len(rows)
        for i, (a, b) in enumerate(rows):
            conn[k, i] = [a, b, 0.8, 0, 0]

    put(0, [(ids[1][0], ids[2][0]), (ids[1][1], ids[2][1]), (ids[1][2], ids[2][2])])   # neck -> Rsho: three persons
    put(12, [(ids[1][0], ids[0][0]), (ids[1][1], ids[0][1])])                          # neck -> nose
    put(13, [(ids[0][0], ids[14][0]), (ids[0][1], ids[14][1])])                        # nose -> Reye
    # Reye -> Rear: the same ear twice.  Person 2 takes it first; person 1's connection then finds two rows (its own through
    # the eye, person 2's through the ear), they overlap, so the FIRST row (person 1) gets the ear as well: one ear, two rows
    put(14, [(ids[14][1], ids[16][0]), (ids[14][0], ids[16][0])])
    put(17, [(ids[2][2], ids[16][0])])                                                 # Rsho of person 3 -> that ear
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_cand, d_conn, d_nc, d_np = t(cand), t(conn), t(n_conn), t(n_peaks)
    sub = torch.zeros((MS, 20), dtype=torch.float64, device=dev)
    nsub = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    L.check(lib.rmpe_debug_assemble(MP, MS, d_cand.data_ptr(), d_conn.data_ptr(), d_nc.data_ptr(), d_np.data_ptr(),
                                    sub.data_ptr(), nsub.data_ptr(), st.data_ptr(), None))
    torch.cuda.synchronize()
    assert int(st.item()) & L.ST_FOUND_GT2
    with pytest.raises(IndexError):
        rmpe.decode._raise_status(int(st.item()))
    # the oracle's restatement of the same loop flags the same connection (and keeps the first two rows, like the kernel)
    all_peaks = [[tuple(cand[i]) for i in ids[p]] for p in range(18)]
    special = [k for k in range(19) if n_conn[k] < 0]
    conn_all = [conn[k, :max(n_conn[k], 0)] for k in range(19)]
    _, osub, overflow = do.assemble(all_peaks, conn_all, special)
    assert overflow
    assert np.array_equal(sub.cpu().numpy()[:int(nsub.item())], osub)
    # without the duplicated ear nothing is flagged
    put(14, [(ids[14][0], ids[16][0])])
    L.check(lib.rmpe_debug_assemble(MP, MS, d_cand.data_ptr(), t(conn).data_ptr(), t(n_conn).data_ptr(), d_np.data_ptr(),
                                    sub.data_ptr(), nsub.data_ptr(), st.data_ptr(), None))
    torch.cuda.synchronize()
    assert int(st.item()) == 0


def test_host_wrapper_rejects_descriptors_beyond_the_blobs(rmpe):
    paf = np.zeros((12, 15, 38), np.float32)
    heat = np.zeros((12, 15, 19), np.float32)
    desc, hflat, pflat = rmpe.batch.make_frames([dict(H=96, W=120, scales=[(paf, heat, 0, 0)])])
    desc[0]["heat_offset"][0] = 19       # one cell too far
    with pytest.raises(rmpe.lib.RmpeError):
        rmpe.batch.decode_batch_host_raw(desc, hflat, pflat)


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


@pytest.mark.parametrize("backend", ["nccl", "gloo"])
def test_decode_records_gather_across_ranks(rmpe, backend):
    """The result tail of compute_keypoints (eval...:497-548) over ranks: frames i % world == rank are decoded on the rank's
    GPU, turned into COCO keypoint records and gathered (all_gather_object).  NCCL with one rank per GPU when the box has
    two GPUs (else a one-rank NCCL group); gloo with two ranks sharing GPU 0."""
    import subprocess
    import sys
    import torch
    n = min(2, torch.cuda.device_count()) if backend == "nccl" else 2
    env = dict(os.environ, RMPE_TEST_BACKEND=backend)
    port = 29600 + (os.getpid() % 300) + (0 if backend == "nccl" else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(os.path.dirname(os.path.abspath(__file__)), "_gather_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "GATHER_OK" in out.stdout
