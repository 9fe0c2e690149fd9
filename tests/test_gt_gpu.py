"""GPU parity of the target-generation path, through the C ABI, against the oracle and the
reference's golden vectors.  Bars (BASELINE.json north_star): warped pixels, mask, joints and PAF
counts bit-exact; confidence maps / PAF vectors max-abs <= 1e-5."""
import numpy as np
import pytest

from cases import GT_CASES, FULL_LABEL_CASES, gt_case_inputs, sha
from oracle import gt_oracle as go

pytestmark = pytest.mark.gpu
LABEL_TOL = 1e-5


def _run_case(rmpe, s, **kw):
    flip, deg, crop, scale = s["aug"]
    M = rmpe.batch.aug_affine([flip], [deg], [crop], [scale], [s["objpos"][0]], [s["scale_provided"][0]])
    P = s["joints"].shape[0]
    r = rmpe.batch.gt_batch_host(s["img"][None], s["mask"][None], s["joints"][None], [P], M, [flip],
                                 want_count=True, **kw)
    return M[0], r


@pytest.mark.parametrize("simple", [False, True], ids=["tile", "simple"])
@pytest.mark.parametrize("case", GT_CASES, ids=[c[0] for c in GT_CASES])
def test_gt_matches_reference_golden(rmpe, gt_golden, case, simple):
    name = case[0]
    s = gt_case_inputs(case)
    M, r = _run_case(rmpe, s, f64=True, simple=simple)
    assert np.array_equal(M, gt_golden[name + "_M"])
    assert sha(r["img"][0]) == str(gt_golden[name + "_img_sha"]), \
        "warped image: %d pixels differ in the 16 stored rows" % int((r["img"][0][::23] != gt_golden[name + "_img_rows"]).sum())
    assert np.array_equal(np.rint(r["mask"][0] * 255).astype(np.uint8), gt_golden[name + "_mask46"])
    assert sha(r["mask"][0]) == str(gt_golden[name + "_mask_sha"])
    if s["joints"].shape[0]:
        assert np.array_equal(r["joints"][0], gt_golden[name + "_joints"])
    # 9 of the 12 golden cases store only the 57 plane sums of the reference's labels (a coarse check); what pins their
    # values element by element is test_gt_matches_oracle below (oracle == reference: tests/test_oracle_pin.py)
    assert np.abs(r["labels"][0].sum(axis=(1, 2)) - gt_golden[name + "_labels_sum"]).max() < 1e-2
    if name in FULL_LABEL_CASES:
        assert np.abs(r["labels"][0] - gt_golden[name + "_labels_f32"].astype(np.float64)).max() <= LABEL_TOL + 1e-7


@pytest.mark.parametrize("case", GT_CASES, ids=[c[0] for c in GT_CASES])
def test_gt_matches_oracle(rmpe, case):
    s = gt_case_inputs(case)
    M, r = _run_case(rmpe, s, f64=True)
    flip = s["aug"][0]
    oimg, omask, oj = go.transform(s["img"], s["mask"], s["joints"], M, flip)
    olab, ocnt = go.create_heatmaps(oj, omask, return_count=True)
    assert np.array_equal(r["img"][0], oimg)
    assert np.array_equal(r["mask"][0], omask)
    assert np.array_equal(r["joints"][0], oj)
    assert np.array_equal(r["count"][0], ocnt)                      # PAF counts bit-exact
    assert np.array_equal(r["labels"][0] != 0, olab != 0) or np.abs(r["labels"][0] - olab).max() <= LABEL_TOL
    assert np.abs(r["labels"][0] - olab).max() <= LABEL_TOL
    assert r["status"][0] == 0


def test_gt_f32_chw_outputs(rmpe):
    s = gt_case_inputs(GT_CASES[1])
    M, r64 = _run_case(rmpe, s, f64=True)
    _, r32 = _run_case(rmpe, s, f64=False, chw=True)
    assert r32["labels"].dtype == np.float32 and r32["mask"].dtype == np.float32
    assert np.array_equal(r32["img"][0], np.transpose(r64["img"][0], (2, 0, 1)))
    assert np.array_equal(r32["mask"][0], r64["mask"][0].astype(np.float32))
    assert np.abs(r32["labels"][0] - r64["labels"][0]).max() <= 2e-7
    assert np.array_equal(r32["count"], r64["count"])


def test_background_is_one_minus_max_where_unmasked(rmpe):
    """The reference's own notebook assertion (check_dataset_iterator.ipynb cell 7)."""
    s = gt_case_inputs(GT_CASES[2])
    _, r = _run_case(rmpe, s, f64=False)
    lab, mask = r["labels"][0], r["mask"][0]
    on = mask == 1
    assert on.any()
    assert np.abs(lab[56][on] - (1 - lab[38:56].max(axis=0))[on]).max() <= 1e-6


def test_drop_in_classes_match_golden(rmpe, gt_golden):
    case = GT_CASES[3]
    name = case[0]
    s = gt_case_inputs(case)
    flip, deg, crop, scale = s["aug"]
    aug = rmpe.transformer.AugmentSelection(flip, deg, crop, scale)
    meta = dict(objpos=s["objpos"], scale_provided=s["scale_provided"], joints=s["joints"].copy())
    img, mask, meta2 = rmpe.transformer.Transformer.transform(s["img"], s["mask"], meta, aug)
    assert meta2 is meta and img.shape == (368, 368, 3) and mask.shape == (46, 46) and mask.dtype == np.float64
    assert sha(img) == str(gt_golden[name + "_img_sha"])
    assert sha(mask) == str(gt_golden[name + "_mask_sha"])
    assert np.array_equal(meta["joints"], gt_golden[name + "_joints"])          # mutated in place
    labels = rmpe.heatmapper.Heatmapper().create_heatmaps(meta["joints"], mask)
    assert labels.shape == (57, 46, 46) and labels.dtype == np.float64
    assert np.abs(labels.sum(axis=(1, 2)) - gt_golden[name + "_labels_sum"]).max() < 1e-2
    olab = go.create_heatmaps(meta["joints"], mask)
    assert np.abs(labels - olab).max() <= LABEL_TOL
    # iterator glue: one fused call gives the same sample
    it = rmpe.data_iterator.RawDataIterator(None, shuffle=False, augment=False)
    meta3 = dict(objpos=s["objpos"], scale_provided=s["scale_provided"], joints=s["joints"].copy())
    i2, m2, _, l2 = it.transform_data(s["img"], s["mask"], meta3)
    assert i2.shape == (368, 368, 3) and l2.shape == (57, 46, 46)


def test_edge_cases(rmpe):
    rng = np.random.RandomState(11)
    # (a) source far larger than the staged footprint (inverse scale 4): falls back to global taps
    img = rng.randint(0, 256, size=(900, 1200, 3)).astype(np.uint8)
    mask = rng.randint(0, 2, size=(900, 1200)).astype(np.uint8) * 255
    joints = rng.uniform(0, 900, size=(1, 2, 18, 3))
    joints[..., 2] = 1
    M = rmpe.batch.aug_affine([0], [25.0], [(10, -5)], [1.0], [(600., 450.)], [2.4])
    r = rmpe.batch.gt_batch_host(img[None], mask[None], joints, [2], M, [0], f64=True)
    oimg, omask, oj = go.transform(img, mask, joints[0], M[0], False)
    assert np.array_equal(r["img"][0], oimg) and np.array_equal(r["mask"][0], omask)
    # (b) tiny source, heavy zoom, odd pitch
    img = rng.randint(0, 256, size=(37, 53, 3)).astype(np.uint8)
    mask = rng.randint(0, 256, size=(37, 53)).astype(np.uint8)
    M = rmpe.batch.aug_affine([1], [-38.0], [(3, 2)], [1.0], [(26., 18.)], [0.08])
    r = rmpe.batch.gt_batch_host(img[None], mask[None], np.zeros((1, 0, 18, 3)), [0], M, [1], f64=True)
    oimg, omask, _ = go.transform(img, mask, np.zeros((0, 18, 3)), M[0], True)
    assert np.array_equal(r["img"][0], oimg) and np.array_equal(r["mask"][0], omask)
    assert (r["labels"][0][:56] == 0).all()
    # (c) zero-length limb is skipped and flagged; absent joints (vis 2) are not drawn
    j = np.zeros((1, 1, 18, 3))
    j[0, 0, :, 0] = 100
    j[0, 0, :, 1] = 120
    j[0, 0, 5:, 2] = 2
    m1 = np.ones((1, 46, 46))
    h = rmpe.batch.heatmaps_host(j, [1], m1, f64=True, want_count=True)
    olab, ocnt = go.create_heatmaps(j[0], m1[0], return_count=True)
    assert h["status"][0] & 1
    assert np.array_equal(h["count"][0], ocnt) and np.abs(h["labels"][0] - olab).max() <= LABEL_TOL
    # (d) singular matrix is flagged, output is all border like cv2
    M0 = np.zeros((1, 2, 3))
    r = rmpe.batch.gt_batch_host(img[None], mask[None], np.zeros((1, 0, 18, 3)), [0], M0, [0], f64=True)
    assert r["status"][0] & 0x20
    oimg, omask, _ = go.transform(img, mask, np.zeros((0, 18, 3)), M0[0], False)
    assert np.array_equal(r["img"][0], oimg) and np.array_equal(r["mask"][0], omask)


def test_tie_prone_integer_joints_counts_exact(rmpe):
    """Axis-aligned limbs with integer joints put pixels exactly on dist == 8.0 (fp64 decision)."""
    rng = np.random.RandomState(3)
    j = np.zeros((1, 6, 18, 3))
    j[..., 0] = rng.randint(0, 46, size=(1, 6, 18)) * 8
    j[..., 1] = rng.randint(0, 46, size=(1, 6, 18)) * 8
    j[..., 2] = (rng.uniform(size=(1, 6, 18)) < 0.1) * 2.0
    m = np.ones((1, 46, 46))
    h = rmpe.batch.heatmaps_host(j, [6], m, f64=True, want_count=True)
    olab, ocnt = go.create_heatmaps(j[0], m[0], return_count=True)
    assert np.array_equal(h["count"][0], ocnt)
    assert np.array_equal(h["labels"][0][:38] != 0, olab[:38] != 0)
    assert np.abs(h["labels"][0] - olab).max() <= LABEL_TOL


def test_full_size_batch_properties(rmpe):
    """BASELINE config 2 size (batch 256, 3 persons) on device-resident buffers: the staged
    (TMA bulk copy + dp4a) warp equals the straight-line kernel bit for bit, a handful of samples
    equal the oracle, and size-independent properties hold for all 256."""
    import torch
    B, P = 256, 3
    b = rmpe.synth.gt_batch(B, n_persons=P, seed0=1000)
    flip = np.array([a[0] for a in b["augs"]], np.uint8)
    M = rmpe.batch.aug_affine(flip, [a[1] for a in b["augs"]], [a[2] for a in b["augs"]],
                              [a[3] for a in b["augs"]], b["centers"], b["scale_self"])
    plan = rmpe.batch.GtDevicePlan(B, P, want_count=True)
    plan.upload(b["imgs"], b["masks"], b["joints"], b["n_persons"], M, flip)
    plan.run(simple=True)
    torch.cuda.synchronize()
    img_simple = plan.out_img.cpu().numpy().copy()
    lab_simple = plan.out_labels.cpu().numpy().copy()
    plan.out_img.zero_()
    plan.run(simple=False)
    torch.cuda.synchronize()
    img_tile = plan.out_img.cpu().numpy()
    assert np.array_equal(img_tile, img_simple)
    assert np.array_equal(plan.out_labels.cpu().numpy(), lab_simple)
    assert (plan.status.cpu().numpy() == 0).all()
    lab = plan.out_labels.cpu().numpy()
    mask = plan.out_mask.cpu().numpy()
    cnt = plan.out_count.cpu().numpy()
    # properties: bkg = (1-max heat)*mask; PAF vectors are unit (or zero) where mask==1; count>0 <=> PAF set
    assert np.abs(lab[:, 56] - (1 - lab[:, 38:56].max(axis=1) / np.where(mask > 0, mask, 1)) * mask).max() <= 2e-6
    full = mask == 1
    n2 = lab[:, 0:38:2] ** 2 + lab[:, 1:38:2] ** 2
    sel = np.broadcast_to(full[:, None], n2.shape)
    assert np.all((np.abs(n2[sel] - 1) < 1e-5) | (n2[sel] == 0))
    assert np.array_equal((cnt > 0)[sel], (n2 > 0)[sel])
    for i in (0, 17, 101, 255):
        oimg, omask, oj = go.transform(b["imgs"][i], b["masks"][i], b["joints"][i], M[i], bool(flip[i]))
        assert np.array_equal(img_tile[i], oimg)
        assert np.array_equal(mask[i], omask.astype(np.float32))
        olab, ocnt = go.create_heatmaps(oj, omask, return_count=True)
        assert np.array_equal(cnt[i], ocnt)
        assert np.abs(lab[i] - olab).max() <= LABEL_TOL


def test_crowded_20_persons(rmpe):
    s = rmpe.synth.gt_sample(77, 20)
    M, r = _run_case(rmpe, s, f64=True)
    oimg, omask, oj = go.transform(s["img"], s["mask"], s["joints"], M, s["aug"][0])
    olab, ocnt = go.create_heatmaps(oj, omask, return_count=True)
    assert np.array_equal(r["img"][0], oimg) and np.array_equal(r["count"][0], ocnt)
    assert np.abs(r["labels"][0] - olab).max() <= LABEL_TOL


@pytest.mark.parametrize("P", [5, 12, 20, 33, 64])
@pytest.mark.parametrize("f64", [True, False], ids=["f64", "f32"])
def test_crowded_rasteriser_person_counts(rmpe, P, f64):
    """The crowded rasteriser (k_raster_blocks) over its regimes: all 18 parts' distance tables resident (P <= 27),
    parts in batches beyond that (33, 64 persons); person masks of up to 64 bits in the limb scatter."""
    s = rmpe.synth.gt_sample(500 + P, P)
    M, r = _run_case(rmpe, s, f64=f64)
    _, omask, oj = go.transform(s["img"], s["mask"], s["joints"], M, s["aug"][0])
    olab, ocnt = go.create_heatmaps(oj, omask, return_count=True)
    assert r["status"][0] == 0
    assert np.array_equal(r["count"][0], ocnt)
    assert np.array_equal(r["joints"][0], oj)
    assert np.abs(r["labels"][0] - olab).max() <= LABEL_TOL + (0 if f64 else 1e-7)


@pytest.mark.parametrize("thre", [8., 5.5])
def test_crowded_limbs_far_outside_the_grid(rmpe, thre):
    """k_raster_blocks decides the band test in float only where the float numerator is further from the limit than its
    rounding bound: long limbs whose joints lie thousands of pixels outside the crop (large products, small margins
    relative to them) and limbs that graze cell corners must still give the reference's counts bit for bit."""
    rng = np.random.RandomState(23)
    P = 9
    j = np.zeros((P, 18, 3))
    j[..., 0] = rng.uniform(-3000, 3300, size=(P, 18))
    j[..., 1] = rng.uniform(-3000, 3300, size=(P, 18))
    j[:3, :, :2] = rng.randint(-40, 90, size=(3, 18, 2)) * 8.0             # cell-corner lattice: dd lands exactly on the limit
    j[3, :, :2] = rng.uniform(0, 368, size=(18, 2))
    j[..., 2] = rng.choice([0., 1., 2.], size=(P, 18), p=[.2, .7, .1])
    mask = np.ones((46, 46))
    lab, cnt = rmpe.heatmapper.Heatmapper(7., thre).create_heatmaps(j, mask, return_count=True)
    olab, ocnt = go.create_heatmaps(j, mask, 7., thre, return_count=True)
    assert ocnt.sum() > 1000
    assert np.array_equal(cnt, ocnt)
    assert np.abs(lab - olab).max() <= LABEL_TOL


@pytest.mark.parametrize("sigma,thre", [(5., 8.), (7., 6.5), (9.5, 12.), (3., 4.)])
def test_heatmapper_sigma_thre(rmpe, sigma, thre):
    """Heatmapper(sigma, thre) (py_rmpe_heatmapper.py:10-14) with non-default values: a non-power-of-two thre takes the
    division in the band test, a power-of-two one the exact product form."""
    s = rmpe.synth.gt_sample(61, 3, augment=False)
    mask = np.ones((46, 46))
    lab, cnt = rmpe.heatmapper.Heatmapper(sigma, thre).create_heatmaps(s["joints"], mask, return_count=True)
    olab, ocnt = go.create_heatmaps(s["joints"], mask, sigma, thre, return_count=True)
    assert np.array_equal(cnt, ocnt)
    assert np.abs(lab - olab).max() <= LABEL_TOL
    # tie-prone integer joints at this thre
    rng = np.random.RandomState(9)
    j = np.zeros((5, 18, 3))
    j[..., 0] = rng.randint(0, 46, size=(5, 18)) * 8
    j[..., 1] = rng.randint(0, 46, size=(5, 18)) * 8
    lab, cnt = rmpe.heatmapper.Heatmapper(sigma, thre).create_heatmaps(j, mask, return_count=True)
    olab, ocnt = go.create_heatmaps(j, mask, sigma, thre, return_count=True)
    assert np.array_equal(cnt, ocnt) and np.abs(lab - olab).max() <= LABEL_TOL


def test_heatmapper_building_blocks(rmpe):
    """put_joints / put_gaussian_maps / put_limbs / put_vector_maps (py_rmpe_heatmapper.py:47-138), the reference's
    in-place building blocks: composed like create_heatmaps (:32-44) they give create_heatmaps; on a stack that already
    holds values the Gaussian layers merge by max and the PAF layers are overwritten only inside a band."""
    G = rmpe.config.RmpeGlobalConfig
    s = rmpe.synth.gt_sample(88, 6, augment=False)
    joints = s["joints"]
    mask = np.random.RandomState(4).randint(0, 256, size=(46, 46)) / 255.0
    hm = rmpe.heatmapper.Heatmapper()
    want = hm.create_heatmaps(joints, mask)
    got = np.zeros(G.parts_shape)
    hm.put_joints(got, joints)
    sl = slice(G.heat_start, G.heat_start + G.heat_layers)
    got[G.bkg_start] = 1. - np.amax(got[sl], axis=0)
    hm.put_limbs(got, joints)
    got *= mask
    # identical except the background, which the kernel forms as 1 - max in float32 and this composition in float64
    assert np.array_equal(got[:G.bkg_start], want[:G.bkg_start]) and np.abs(got[G.bkg_start] - want[G.bkg_start]).max() <= 1e-7
    olab = go.create_heatmaps(joints, mask)
    assert np.abs(got - olab).max() <= LABEL_TOL
    # layer by layer, on a pre-filled stack
    pre = np.full(G.parts_shape, 0.25)
    hm.put_gaussian_maps(pre, 3, joints[joints[:, 3, 2] < 2][:, 3, 0:2])
    ref = np.maximum(0.25, go.create_heatmaps(joints, np.ones((46, 46)))[G.heat_start + 3])
    assert np.abs(pre[G.heat_start + 3] - ref).max() <= LABEL_TOL and np.all(pre[:G.heat_start] == 0.25)
    fr, to = G.limbs_conn[4]
    vis = (joints[:, fr, 2] < 2) & (joints[:, to, 2] < 2)
    pre = np.full(G.parts_shape, 0.25)
    hm.put_vector_maps(pre, 10, 11, joints[vis, fr, 0:2], joints[vis, to, 0:2])
    full, cnt = go.create_heatmaps(joints, np.ones((46, 46)), return_count=True)
    hit = cnt[4] > 0
    assert hit.any() and np.all(pre[10][~hit] == 0.25) and np.all(pre[11][~hit] == 0.25)
    assert np.abs(pre[10][hit] - full[G.paf_start + 8][hit]).max() <= LABEL_TOL
    assert np.abs(pre[11][hit] - full[G.paf_start + 9][hit]).max() <= LABEL_TOL
    # more pairs than one rasteriser call holds (64): later pairs still win
    rng = np.random.RandomState(6)
    jf = rng.uniform(0, 368, size=(70, 2)); jt = jf + rng.uniform(-60, 60, size=(70, 2))
    a = np.zeros(G.parts_shape); hm.put_vector_maps(a, 0, 1, jf, jt)
    persons = np.zeros((70, 18, 3)); persons[:, :, 2] = 2.0
    f0, t0 = G.limbs_conn[0]
    persons[:, f0, 0:2] = jf; persons[:, t0, 0:2] = jt; persons[:, f0, 2] = persons[:, t0, 2] = 1.0
    ref70 = go.create_heatmaps(persons, np.ones((46, 46)))
    assert np.abs(a[0] - ref70[0]).max() <= LABEL_TOL and np.abs(a[1] - ref70[1]).max() <= LABEL_TOL


def test_paf_average_variant(rmpe):
    """Non-default flag: the averaging the reference keeps commented out (py_rmpe_heatmapper.py:119-126)."""
    s = rmpe.synth.gt_sample(77, 20, augment=False)
    s["joints"][:, :, :2] += np.random.RandomState(1).normal(0, 15, size=(20, 18, 2))   # limbs of different directions
    mask = np.ones((46, 46))
    lab, cnt = rmpe.heatmapper.Heatmapper().create_heatmaps(s["joints"], mask, return_count=True, paf_average=True)
    olab, ocnt = go.create_heatmaps(s["joints"], mask, return_count=True, paf_average=True)
    assert (ocnt > 1).any()                     # crowded enough for bands to overlap
    assert np.array_equal(cnt, ocnt) and np.abs(lab - olab).max() <= LABEL_TOL
    ref = go.create_heatmaps(s["joints"], mask)
    assert np.abs(ref - olab).max() > 1e-3      # and it is a different result from the reference's overwrite


@pytest.mark.parametrize("P,f64", [(3, False), (3, True), (20, False), (20, True), (30, False)])
def test_keras_tensors_from_the_rasteriser(rmpe, P, f64):
    """RMPE NHWC outputs (training/ds_generators.py:52-63) written by the rasteriser itself == the planar labels
    re-laid-out by the oracle; P = 30 takes the two-pass route (tables of 30 persons + a band do not fit one CTA)."""
    b = rmpe.synth.gt_batch(3, n_persons=P, seed0=810)
    flip = np.array([a[0] for a in b["augs"]], np.uint8)
    M = rmpe.batch.aug_affine(flip, [a[1] for a in b["augs"]], [a[2] for a in b["augs"]], [a[3] for a in b["augs"]],
                              b["centers"], b["scale_self"])
    r = rmpe.batch.gt_batch_host(b["imgs"], b["masks"], b["joints"], b["n_persons"], M, flip, f64=f64, keras=True,
                                 want_count=True)
    r2 = rmpe.batch.gt_batch_host(b["imgs"], b["masks"], b["joints"], b["n_persons"], M, flip, f64=f64, keras=True,
                                  want_labels=(P > 24), keras_weights=False)
    ft = np.float64 if f64 else np.float32
    for i in range(3):
        _, omask, oj = go.transform(b["imgs"][i], b["masks"][i], b["joints"][i], M[i], bool(flip[i]))
        olab, ocnt = go.create_heatmaps(oj, omask, return_count=True)
        assert np.array_equal(r["count"][i], ocnt)
        assert np.abs(r["labels"][i] - olab).max() <= LABEL_TOL + (0 if f64 else 1e-7)
    x1, x2, y1, y2 = go.keras_batch(r["labels"], r["mask"])       # the same values, NHWC
    for k, want in (("x1", x1), ("x2", x2), ("y1", y1), ("y2", y2)):
        assert r[k].dtype == ft and np.array_equal(r[k], want), k
    assert np.array_equal(r2["y1"], y1) and np.array_equal(r2["y2"], y2) and r2["x1"] is None


def test_n_persons_beyond_the_person_stride_is_clamped_and_flagged(rmpe):
    """Device entry: n_persons is caller data; a count above max_persons must not read the next sample's joints."""
    import ctypes as C
    import torch
    L = rmpe.lib
    plan = rmpe.batch.GtDevicePlan(2, 3)
    b = rmpe.synth.gt_batch(2, n_persons=3, seed0=40)
    flip = np.array([a[0] for a in b["augs"]], np.uint8)
    M = rmpe.batch.aug_affine(flip, [a[1] for a in b["augs"]], [a[2] for a in b["augs"]], [a[3] for a in b["augs"]],
                              b["centers"], b["scale_self"])
    plan.upload(b["imgs"], b["masks"], b["joints"], np.array([7, -2], np.int32), M, flip)
    plan.run()
    torch.cuda.synchronize()
    st = plan.status.cpu().numpy()
    assert (st & L.ST_PERSONS_CLAMPED).all()
    lab = plan.out_labels.cpu().numpy()
    for i, n in enumerate((3, 0)):
        _, omask, oj = go.transform(b["imgs"][i], b["masks"][i], b["joints"][i][:n], M[i], bool(flip[i]))
        assert np.abs(lab[i] - go.create_heatmaps(oj, omask)).max() <= LABEL_TOL + 1e-7
    # host entry: the same mistake is an argument error
    with pytest.raises(L.RmpeError):
        rmpe.batch.gt_batch_host(b["imgs"], b["masks"], b["joints"], [4, 1], M, flip)


def test_host_calls_from_threads_do_not_serialise_or_mix(rmpe):
    """Keras worker threads (ds_generators.py:209 under fit_generator): every thread has its own arena and streams."""
    import threading
    b = rmpe.synth.gt_batch(6, n_persons=2, seed0=70)
    flip = np.array([a[0] for a in b["augs"]], np.uint8)
    M = rmpe.batch.aug_affine(flip, [a[1] for a in b["augs"]], [a[2] for a in b["augs"]], [a[3] for a in b["augs"]],
                              b["centers"], b["scale_self"])
    want = rmpe.batch.gt_batch_host(b["imgs"], b["masks"], b["joints"], b["n_persons"], M, flip)
    got, errs = {}, []

    def work(t):
        try:
            for _ in range(5):
                sl = slice(t, t + 3)
                got[t] = rmpe.batch.gt_batch_host(b["imgs"][sl], b["masks"][sl], b["joints"][sl], b["n_persons"][sl],
                                                  M[sl], flip[sl])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for t in range(4):
        for k in ("img", "mask", "labels", "joints"):
            assert np.array_equal(got[t][k], want[k][t:t + 3]), (t, k)


@pytest.mark.parametrize("seed", range(10))
def test_random_geometry_sweep(rmpe, seed):
    """Fused warp+mask kernel against the oracle over odd source sizes (unaligned row pitches),
    zoom factors on both sides of the staging capacity, rotations beyond the augmentation range
    and centres near / outside the frame (all-border tiles)."""
    rng = np.random.RandomState(4000 + seed)
    H, W = int(rng.randint(40, 700)), int(rng.randint(40, 700))
    img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    mask = (rng.randint(0, 4, size=(H, W)) > 0).astype(np.uint8) * 255
    if seed % 2:
        mask = rng.randint(0, 256, size=(H, W)).astype(np.uint8)
    flip = int(rng.randint(0, 2))
    deg = float(rng.uniform(-180, 180)) if seed % 3 == 0 else float(rng.uniform(-40, 40))
    scale_self = float(np.exp(rng.uniform(np.log(0.12), np.log(2.5))))
    center = (float(rng.uniform(-0.2, 1.2) * W), float(rng.uniform(-0.2, 1.2) * H))
    crop = (int(rng.randint(-40, 41)), int(rng.randint(-40, 41)))
    M = rmpe.batch.aug_affine([flip], [deg], [crop], [1.0], [center], [scale_self])
    for chw in (False, True):
        r = rmpe.batch.gt_batch_host(img[None], mask[None], np.zeros((1, 0, 18, 3)), [0], M, [flip], f64=True, chw=chw)
        oimg, omask, _ = go.transform(img, mask, np.zeros((0, 18, 3)), M[0], bool(flip))
        got = r["img"][0] if not chw else np.transpose(r["img"][0], (1, 2, 0))
        assert np.array_equal(got, oimg), "%d warped pixels differ" % int((got != oimg).sum())
        assert np.array_equal(r["mask"][0], omask)


@pytest.mark.parametrize("f64", [True, False], ids=["f64", "f32"])
def test_keras_batch_assembly(rmpe, f64):
    """DataIteratorBase.gen's NHWC tensors (training/ds_generators.py:47-77): pure data movement, bit-exact."""
    rng = np.random.RandomState(5)
    ft = np.float64 if f64 else np.float32
    labels = rng.uniform(-1, 1, size=(5, 57, 46, 46)).astype(ft)
    mask = rng.uniform(0, 1, size=(5, 46, 46)).astype(ft)
    kb = rmpe.batch.keras_batch_host(labels, mask)
    x1, x2, y1, y2 = go.keras_batch(labels, mask)
    for got, want in ((kb["x1"], x1), (kb["x2"], x2), (kb["y1"], y1), (kb["y2"], y2)):
        assert got.dtype == ft and got.shape == want.shape and np.array_equal(got, want)


def test_fused_data_iterator_batches(rmpe):
    """FusedDataIterator: one GT call + one assembly call per batch == the reference's per-sample route."""
    import random
    samples = [rmpe.synth.gt_sample(300 + i, 1 + (i % 3)) for i in range(4)]
    src = [(s["img"], s["mask"], dict(objpos=s["objpos"], scale_provided=s["scale_provided"], joints=s["joints"].copy()))
           for s in samples]
    random.seed(7)
    it = rmpe.ds_generators.FusedDataIterator(iter(src), augment=True, batch_size=4)
    (x, x1, x2), ys = next(it.gen(n_stages=2))
    assert x.shape == (4, 368, 368, 3) and x1.shape == (4, 46, 46, 38) and x2.shape == (4, 46, 46, 19) and len(ys) == 4
    random.seed(7)
    augs = [rmpe.transformer.AugmentSelection.random() for _ in samples]
    for i, (s, aug) in enumerate(zip(samples, augs)):
        M = go.affine_closed_form(aug.flip, aug.degree, aug.crop, aug.scale, s["objpos"][0], s["scale_provided"][0])
        oimg, omask, oj = go.transform(s["img"], s["mask"], s["joints"], M, aug.flip)
        olab = go.create_heatmaps(oj, omask)
        assert np.array_equal(x[i], oimg)
        assert np.array_equal(x1[i], np.repeat(omask[:, :, None], 38, axis=2))
        assert np.abs(ys[0][i] - np.transpose(olab[:38], (1, 2, 0))).max() <= LABEL_TOL
        assert np.abs(ys[1][i] - np.transpose(olab[38:], (1, 2, 0))).max() <= LABEL_TOL
        assert np.array_equal(it.keypoints[i], oj)          # transformed joints handed to the accuracy code


class _FakeEntry:
    """One HDF5 dataset of generate_hdf5_coco2014.py: a (6,H,W) u8 datum + a JSON `meta` attribute."""

    def __init__(self, data, meta):
        import json
        self._data = data
        self.attrs = {'meta': json.dumps(meta)}

    def __getitem__(self, key):
        return self._data


def _fake_datum(shapes, seed0=900):
    datum = {}
    for i, (H, W) in enumerate(shapes):
        rng = np.random.RandomState(seed0 + i)
        data = rng.randint(0, 256, (6, H, W)).astype(np.uint8)
        data[4] = 255
        data[4, H // 4:H // 2, W // 3:W // 2] = 0            # a miss-mask rectangle
        P = 1 + i % 3
        joints = np.zeros((P, 17, 3))
        joints[:, :, 0] = rng.uniform(0, W, (P, 17))
        joints[:, :, 1] = rng.uniform(0, H, (P, 17))
        joints[:, :, 2] = rng.choice([0., 1., 2.], (P, 17), p=[.2, .7, .1])
        meta = dict(joints=joints.tolist(), objpos=[[W / 2. + 3 * i, H / 2. - 2 * i]], scale_provided=[0.45 + 0.1 * i])
        datum["%07d" % i] = _FakeEntry(data, meta)
    return datum


def test_raw_iterator_batched_equals_per_sample(rmpe):
    """RawDataIterator.gen_batched (one C-ABI call per batch, ragged sources padded with the border constants) ==
    RawDataIterator.gen (one call per sample), bit for bit; one sample is also checked against the oracle."""
    import random
    shapes = [(240, 320), (368, 368), (300, 200), (427, 640), (180, 500), (333, 500), (96, 128)]
    it = rmpe.data_iterator.RawDataIterator(None, shuffle=False, augment=True)
    it.datum = _fake_datum(shapes)
    assert it.num_keys() == len(shapes)
    random.seed(11)
    one = [tuple(np.array(a, copy=True) for a in tpl) for tpl in it.gen()]
    it.datum = _fake_datum(shapes)
    random.seed(11)
    many = [tuple(np.array(a, copy=True) for a in tpl) for tpl in it.gen_batched(4)]
    assert len(one) == len(many) == len(shapes)
    for a, b in zip(one, many):
        assert a[0].shape == (3, 368, 368) and a[1].shape == (46, 46) and a[2].shape == (57, 46, 46)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and np.array_equal(x, y)
    # oracle on the first sample (same RNG draw order as AugmentSelection.random)
    random.seed(11)
    aug = rmpe.transformer.AugmentSelection.random()
    img, mask, meta = it.read_data("0000000")
    M = go.affine_closed_form(aug.flip, aug.degree, aug.crop, aug.scale, meta['objpos'][0], meta['scale_provided'][0])
    oimg, omask, oj = go.transform(np.ascontiguousarray(img), np.ascontiguousarray(mask), meta['joints'], M, aug.flip)
    olab = go.create_heatmaps(oj, omask)
    assert np.array_equal(many[0][0], np.transpose(oimg, (2, 0, 1)))
    assert np.array_equal(many[0][1], omask) and np.array_equal(many[0][3], oj)
    assert np.abs(many[0][2] - olab).max() <= LABEL_TOL


def test_raw_data_iterator_from_an_hdf5_file(rmpe):
    """RawDataIterator(h5file).gen() on the committed fixture (the reference's on-disk format, read without h5py): every
    sample through the C ABI against the oracle run on what read_data returned (py_rmpe_data_iterator.py:22-41)."""
    import os
    import random
    from cases import ROOT
    path = os.path.join(ROOT, "tests", "golden", "datum_3samples.h5")
    it = rmpe.data_iterator.RawDataIterator(path, shuffle=False, augment=True)
    assert it.num_keys() == 3
    random.seed(5)
    got = [tuple(np.array(a, copy=True) for a in tpl) for tpl in it.gen()]
    random.seed(5)
    batched = [tuple(np.array(a, copy=True) for a in tpl) for tpl in it.gen_batched(2)]
    random.seed(5)
    for key, tpl, tplb in zip(sorted(it.datum.keys()), got, batched):
        aug = rmpe.transformer.AugmentSelection.random()
        img, mask, meta = it.read_data(key)
        M = go.affine_closed_form(aug.flip, aug.degree, aug.crop, aug.scale, meta['objpos'][0], meta['scale_provided'][0])
        oimg, omask, oj = go.transform(np.ascontiguousarray(img), np.ascontiguousarray(mask), meta['joints'], M, aug.flip)
        olab = go.create_heatmaps(oj, omask)
        assert np.array_equal(tpl[0], np.transpose(oimg, (2, 0, 1)))
        assert np.array_equal(tpl[1], omask) and np.array_equal(tpl[3], oj)
        assert np.abs(tpl[2] - olab).max() <= LABEL_TOL
        for a, b in zip(tpl, tplb):
            assert np.array_equal(a, b)


_VARIANT_SCRIPT = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
import rmpe_b200
rmpe_b200.lib.ensure_init(0)
h = hashlib.sha256()
for seed0, P, hw in ((700, 3, (368, 368)), (720, 2, (240, 320)), (740, 4, (427, 640)), (760, 6, (368, 368))):
    b = rmpe_b200.synth.gt_batch(6, n_persons=P, seed0=seed0, src_hw=hw)
    flip = np.array([a[0] for a in b["augs"]], np.uint8)
    # person scales from 0.3 to 1.5 of the crop: up-scaling, down-scaling past the staging buffer (generic taps), borders
    ss = b["scale_self"] * np.random.RandomState(seed0).uniform(0.5, 2.5, 6)
    M = rmpe_b200.batch.aug_affine(flip, [a[1] for a in b["augs"]], [a[2] for a in b["augs"]], [a[3] for a in b["augs"]],
                                   b["centers"], ss)
    import os
    r = rmpe_b200.batch.gt_batch_host(b["imgs"], b["masks"], b["joints"], b["n_persons"], M, flip, want_count=True,
                                      simple=bool(os.environ.get("RMPE_GT_SIMPLE")))
    for k in ("img", "mask", "labels", "joints", "count", "status"):
        h.update(np.ascontiguousarray(r[k]).tobytes())
print("SHA", h.hexdigest())
"""


@pytest.mark.parametrize("env", [{"RMPE_WARP_GROUPS": "4"}, {"RMPE_WARP_GROUPS": "6"}, {"RMPE_WARP_GROUPS": "8"},
                                 {"RMPE_WARP_PREFETCH": "1"}, {"RMPE_RASTER_GROUPS": "2"}, {"RMPE_RASTER_GROUPS": "4"},
                                 {"RMPE_GT_SIMPLE": "1"}],
                         ids=["groups4", "groups6", "groups8", "prefetch", "raster2", "raster4", "simple"])
def test_kernel_variants_are_bit_identical(rmpe, env):
    """The A/B variants kept in the library (tile groups per SM, L2 prefetch, rasteriser plane groups, straight-line kernels)
    read their switch once per process: each runs in its own interpreter and must reproduce the default's outputs bit for
    bit."""
    import os
    import subprocess
    import sys
    from cases import ROOT

    def run(extra):
        e = dict(os.environ)
        for k in ("RMPE_WARP_GROUPS", "RMPE_WARP_PREFETCH", "RMPE_RASTER_GROUPS", "RMPE_GT_SIMPLE"):
            e.pop(k, None)
        e.update(extra)
        out = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT % ROOT], env=e, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        return [l for l in out.stdout.splitlines() if l.startswith("SHA")][0]

    if not hasattr(test_kernel_variants_are_bit_identical, "_default"):
        test_kernel_variants_are_bit_identical._default = run({})
    assert run(env) == test_kernel_variants_are_bit_identical._default
