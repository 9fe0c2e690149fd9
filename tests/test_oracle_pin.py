"""CPU: the oracle restatement is pinned (a) against the golden vectors produced by the real
reference (tests/golden/, oracle/make_golden.py), (b) against the installed OpenCV / SciPy whose
arithmetic it restates, and (c) against the live reference when /root/reference is present."""
import os
import re

import numpy as np
import pytest

from cases import GT_CASES, DECODE_CASES, FULL_LABEL_CASES, gt_case_inputs, decode_case_inputs, sha, ROOT
from oracle import gt_oracle as go
from oracle import decode_oracle as do
from oracle import ref_shim


@pytest.mark.parametrize("case", GT_CASES, ids=[c[0] for c in GT_CASES])
def test_gt_oracle_matches_reference_golden(case, gt_golden):
    name = case[0]
    s = gt_case_inputs(case)
    assert str(gt_golden[name + "_in_sha"]) == sha(s["img"]) + sha(s["mask"]) + sha(s["joints"]), "generator drift"
    flip, deg, crop, scale = s["aug"]
    M = go.affine_closed_form(flip, deg, crop, scale, s["objpos"][0], s["scale_provided"][0])
    assert np.array_equal(M, gt_golden[name + "_M"])
    assert np.array_equal(M, go.affine_chain(flip, deg, crop, scale, s["objpos"][0], s["scale_provided"][0]))
    img, mask, joints = go.transform(s["img"], s["mask"], s["joints"], M, flip)
    assert sha(img) == str(gt_golden[name + "_img_sha"])
    assert np.array_equal(img[::23], gt_golden[name + "_img_rows"])
    assert np.array_equal(np.rint(mask * 255).astype(np.uint8), gt_golden[name + "_mask46"])
    assert sha(mask) == str(gt_golden[name + "_mask_sha"])
    assert np.array_equal(go.mask46_from_src(s["mask"], M), gt_golden[name + "_mask46"])   # fused T2+T3
    assert np.array_equal(joints, gt_golden[name + "_joints"])
    labels = go.create_heatmaps(joints, mask)
    assert sha(labels) == str(gt_golden[name + "_labels_sha"])                              # bit-exact f64
    if name in FULL_LABEL_CASES:
        assert np.array_equal(labels.astype(np.float32), gt_golden[name + "_labels_f32"])


@pytest.mark.parametrize("case", DECODE_CASES, ids=[c[0] for c in DECODE_CASES])
def test_decode_oracle_matches_reference_golden(case, decode_golden):
    name, H, W, P, seed, multi = case
    blobs = decode_case_inputs(case)
    assert str(decode_golden[name + "_in_sha"]) == "".join(sha(b[0]) + sha(b[1]) for b in blobs), "generator drift"
    if multi:
        cand, sub = do.multi_scale(blobs, H, W)
    else:
        cand, sub = do.single_scale(blobs[0][0], blobs[0][1], H, W)
    assert np.array_equal(cand, decode_golden[name + "_candidate"])
    assert np.array_equal(sub, decode_golden[name + "_subset"])


def test_bicubic_table_sums_and_known_entries():
    tab = go.bicubic_tab_i16()
    assert (tab.astype(np.int64).sum(axis=(2, 3)) == 32768).all()
    # phase (0,0): the centre tap saturates at 32767 and the correction lands on tap (2,2)
    assert tab[0, 0, 1, 1] == 32767 and tab[0, 0, 2, 2] == 1


def test_warp_oracle_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(5)
    for t in range(6):
        H, W = rng.randint(40, 200), rng.randint(40, 200)
        img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
        ang = rng.uniform(-3.1, 3.1)
        s = rng.uniform(0.4, 2.5)
        M = np.array([[s * np.cos(ang), s * np.sin(ang), rng.uniform(-50, 150)],
                      [-s * np.sin(ang), s * np.cos(ang), rng.uniform(-50, 150)]])
        if t % 2:
            M[0] *= -1
        ref = cv2.warpAffine(img, M, (368, 368), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT,
                             borderValue=(127, 127, 127))
        assert np.array_equal(go.warp_affine_cubic_u8(img, M, 127), ref)


def test_mask_resize_oracle_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(6)
    for _ in range(4):
        m = rng.randint(0, 256, size=(368, 368)).astype(np.uint8)
        assert np.array_equal(go.mask_resize_46(m), cv2.resize(m, (46, 46), interpolation=cv2.INTER_CUBIC))


@pytest.mark.parametrize("shape", [(20, 23, 163, 181, 19), (84, 89, 674, 712, 19), (23, 31, 184, 248, 38),
                                   (50, 60, 123, 77, 38), (9, 9, 100, 101, 19), (92, 123, 61, 80, 19)])
def test_resize_oracle_vs_cv2(shape):
    cv2 = pytest.importorskip("cv2")
    h, w, H, W, C = shape
    src = np.random.RandomState(h * w).randn(h, w, C).astype(np.float32)
    ref = cv2.resize(src, (W, H), interpolation=cv2.INTER_CUBIC)
    assert np.array_equal(do.resize_cubic_f32(src, W, H), ref)      # bit-exact incl. row tails


def test_resize_fx8_oracle_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    src = np.random.RandomState(3).randn(23, 31, 19).astype(np.float32)
    ref = cv2.resize(src, (0, 0), fx=8, fy=8, interpolation=cv2.INTER_CUBIC)
    assert np.array_equal(do.resize_cubic_f32(src, 248, 184, 8.0, 8.0), ref)


@pytest.mark.parametrize("shape,dt", [((60, 70), np.float32), ((60, 70), np.float64), ((5, 9), np.float32),
                                      ((30, 3), np.float64), ((1, 40), np.float32)])
def test_gaussian_oracle_vs_scipy(shape, dt):
    ndi = pytest.importorskip("scipy.ndimage")
    x = np.random.RandomState(1).rand(*shape).astype(dt)
    a = ndi.gaussian_filter(x, sigma=3)
    b = do.gaussian_filter_sigma3(x)
    assert a.dtype == b.dtype and np.array_equal(a, b)


def test_gauss_constants_in_cuda_source_match_live_weights():
    """The 13 hex-float weights baked into csrc/rmpe_decode.cu are the ones numpy produces here."""
    src = open(os.path.join(ROOT, "adapting-rgb-pose-estimation-to-new-domains_b200", "csrc", "rmpe_decode.cu")).read()
    blk = src[src.index("c_gauss[13]"):]
    blk = blk[:blk.index("};")]
    vals = [float.fromhex(v) for v in re.findall(r"0x1\.[0-9a-f]+p-\d+", blk)]
    w, r = do.gaussian_weights()
    assert r == 12 and len(vals) == 13
    assert vals == [float(v) for v in w[:13]]
    assert np.array_equal(w, w[::-1])


def test_pad_right_down_corner_oracle_vs_reference_semantics():
    img = np.arange(5 * 7 * 3, dtype=np.uint8).reshape(5, 7, 3)
    out, pad = do.pad_right_down_corner(img, 8, 128)
    assert pad == [0, 0, 3, 1] and out.shape == (8, 8, 3)
    assert np.array_equal(out[:5, :7], img) and (out[5:] == 128).all() and (out[:, 7:] == 128).all()
    if ref_shim.available():
        ref = ref_shim.load()
        r, p = ref.util.padRightDownCorner(img, 8, 128)
        assert p == pad and np.array_equal(r, out)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_vs_live_reference_extra_seeds():
    """Beyond the committed fixtures: fresh seeds against the reference itself."""
    import importlib
    synth = importlib.import_module("adapting-rgb-pose-estimation-to-new-domains_b200.synth")
    ref = ref_shim.load()
    for seed in (101, 102, 103):
        s = synth.gt_sample(seed, 4, integer_joints=(seed == 103))
        flip, deg, crop, scale = s["aug"]
        aug = ref.AugmentSelection(flip, deg, crop, scale)
        meta = dict(objpos=s["objpos"], scale_provided=s["scale_provided"], joints=s["joints"].copy())
        rimg, rmask, rmeta = ref.Transformer.transform(s["img"], s["mask"], meta, aug)
        rl = ref.Heatmapper().create_heatmaps(rmeta["joints"], rmask)
        M = go.affine_closed_form(flip, deg, crop, scale, s["objpos"][0], s["scale_provided"][0])
        oimg, omask, oj = go.transform(s["img"], s["mask"], s["joints"], M, flip)
        assert np.array_equal(rimg, oimg) and np.array_equal(rmask, omask) and np.array_equal(rmeta["joints"], oj)
        assert np.array_equal(rl, go.create_heatmaps(oj, omask))


def test_keras_batch_oracle_vs_reference_ds_generators():
    """The batch-assembly restatement against the reference's own DataIteratorBase.gen (h5py stubbed: the
    reference module imports RawDataIterator, which imports h5py, absent in this image)."""
    import sys
    import types
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.load()
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_ds_generators", "/root/reference/training/ds_generators.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.RandomState(2)
    labels = rng.uniform(-1, 1, size=(3, 57, 46, 46))
    mask = rng.uniform(0, 1, size=(3, 46, 46))
    imgs = rng.randint(0, 256, size=(3, 3, 368, 368)).astype(np.uint8)

    class It(mod.DataIteratorBase):
        def __init__(self):
            mod.DataIteratorBase.__init__(self, batch_size=3)
            self.i = 0

        def _recv_arrays(self):
            i = self.i % 3
            self.i += 1
            return imgs[i], mask[i], labels[i]

    (x, x1, x2), ys = next(It().gen(n_stages=2))
    ox1, ox2, oy1, oy2 = go.keras_batch(labels, mask)
    assert np.array_equal(x, np.transpose(imgs, (0, 2, 3, 1)))
    assert np.array_equal(x1, ox1) and np.array_equal(x2, ox2)
    assert np.array_equal(ys[0], oy1) and np.array_equal(ys[1], oy2) and len(ys) == 4


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_draw_canvas_equals_the_references_drawing_tail(tmp_path):
    """D7 (eval...:417-441) stays on cv2 and is outside the GPU path; the drop-in's own implementation must still paint the
    same pixels.  The reference is run end to end on a synthetic frame; its (candidate, subset) feed draw_canvas."""
    import cv2
    import rmpe_b200
    ref = ref_shim.load().eval
    H, W = 240, 320
    h, w = rmpe_b200.synth.single_scale_grid(H, W)
    paf, heat, _ = rmpe_b200.synth.decode_blobs(3, (H, W), (h, w), 2)
    img = np.random.RandomState(0).randint(0, 256, (H, W, 3)).astype(np.uint8)
    path = str(tmp_path / "frame.png")
    cv2.imwrite(path, img)
    params = {'scale_search': [.5, 1, 1.5, 2], 'thre1': .1, 'thre2': .05}
    mparams = {'boxsize': 368, 'stride': 8, 'padValue': 128}
    canvas, cand, sub = ref.process_single_scale(path, ref_shim.FakeModel(lambda hh, ww, call: (paf, heat)), params, mparams)
    assert len(sub) >= 1
    o = do.single_scale(paf, heat, H, W, detail=True)
    n_peaks = np.array([len(p) for p in o["all_peaks"]])
    got = rmpe_b200.decode.draw_canvas(cv2.imread(path), cand, sub, n_peaks)
    assert np.array_equal(got, canvas)
