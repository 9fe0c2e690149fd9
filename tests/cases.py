"""Shared case tables: the golden cases of oracle/make_golden.py rebuilt from seeds."""
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
mg = importlib.import_module("oracle.make_golden")
GT_CASES = mg.GT_CASES
DECODE_CASES = mg.DECODE_CASES
FULL_LABEL_CASES = mg.FULL_LABEL_CASES
gt_case_inputs = mg.gt_case_inputs
decode_case_inputs = mg.decode_case_inputs


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frames_of(case):
    name, H, W, P, seed, multi = case
    blobs = decode_case_inputs(case)
    return dict(H=H, W=W, scales=[(b[0], b[1], b[2], b[3]) for b in blobs])
