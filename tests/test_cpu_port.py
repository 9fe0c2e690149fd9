"""CPU: the cv2/scipy-backed CPU baseline port (oracle/cpu_port.py) reproduces the reference's
golden vectors, so the CPU number bench.py reports is the reference's own arithmetic."""
import numpy as np
import pytest

from cases import GT_CASES, DECODE_CASES, gt_case_inputs, decode_case_inputs, sha
from oracle import gt_oracle as go
from oracle import cpu_port

pytest.importorskip("cv2")
pytest.importorskip("scipy")


@pytest.mark.parametrize("case", [GT_CASES[0], GT_CASES[6], GT_CASES[7]], ids=["g0", "g6", "g7"])
def test_gt_port(case, gt_golden):
    name = case[0]
    s = gt_case_inputs(case)
    flip, deg, crop, scale = s["aug"]
    M = go.affine_closed_form(flip, deg, crop, scale, s["objpos"][0], s["scale_provided"][0])
    img, mask, joints, labels = cpu_port.gt_sample(s["img"], s["mask"], s["joints"], M, flip)
    assert sha(img) == str(gt_golden[name + "_img_sha"])
    assert sha(mask) == str(gt_golden[name + "_mask_sha"])
    assert np.array_equal(joints, gt_golden[name + "_joints"])
    assert sha(labels) == str(gt_golden[name + "_labels_sha"])


@pytest.mark.parametrize("case", [DECODE_CASES[2], DECODE_CASES[4]], ids=["d2", "d4"])
def test_decode_port(case, decode_golden):
    name, H, W, P, seed, multi = case
    blobs = decode_case_inputs(case)
    cand, sub = cpu_port.decode_frame(blobs, H, W)
    assert np.array_equal(cand, decode_golden[name + "_candidate"])
    assert np.array_equal(sub, decode_golden[name + "_subset"])
