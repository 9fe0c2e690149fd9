"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference from /root/reference.

Used by oracle/make_golden.py (fixture generation) and by the CPU tests that pin the
oracle restatement against the real reference when /root/reference is present (it is
NOT present on the GPU box, so nothing on the `-m gpu` / smoke / bench path may import
this file).

Shims:
  * numpy>=1.24 removed np.float / np.int which the reference still uses
    (py_rmpe_heatmapper.py:34,71; py_rmpe_transformer.py:95; py_rmpe_config.py:73)
  * eval/eval_coco2014_multi_modes.py imports keras model / configobj / matplotlib /
    pycocotools / skimage at module import; none is touched by process_*_scale.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "py_rmpe_server"))


_cache = {}


def load():
    """Returns a namespace with the reference's GT classes and the eval module."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from py_rmpe_server.py_rmpe_config import RmpeGlobalConfig, TransformationParams, RmpeCocoConfig
    from py_rmpe_server.py_rmpe_transformer import Transformer, AugmentSelection
    from py_rmpe_server.py_rmpe_heatmapper import Heatmapper

    def _stub(name, **attrs):
        if name in sys.modules:
            return
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    for n, a in [("model", dict(get_testing_model=lambda *a, **k: None)),
                 ("config_reader", dict(config_reader=lambda: None)),
                 ("matplotlib", {}), ("matplotlib.pyplot", {}), ("pylab", dict(rcParams={})),
                 ("pycocotools", {}), ("pycocotools.coco", dict(COCO=object)),
                 ("pycocotools.cocoeval", dict(COCOeval=object)),
                 ("skimage", {}), ("skimage.io", {}), ("IPython", {}),
                 ("IPython.display", dict(Image=object, display=lambda *a: None))]:
        _stub(n, **a)
    import util as ref_util
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = importlib.util.spec_from_file_location(
            "ref_eval", os.path.join(REF_ROOT, "eval", "eval_coco2014_multi_modes.py"))
        ref_eval = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_eval)

    ns = types.SimpleNamespace(
        RmpeGlobalConfig=RmpeGlobalConfig, TransformationParams=TransformationParams,
        RmpeCocoConfig=RmpeCocoConfig, Transformer=Transformer, AugmentSelection=AugmentSelection,
        Heatmapper=Heatmapper, util=ref_util, eval=ref_eval)
    _cache["ns"] = ns
    return ns


class FakeModel:
    """model.predict stand-in: returns caller-provided blobs keyed by input shape (the Keras
    net is upstream of the path)."""

    def __init__(self, blob_fn):
        self.blob_fn = blob_fn
        self.calls = []

    def predict(self, x):
        self.calls.append(x.shape)
        paf, heat = self.blob_fn(x.shape[1], x.shape[2], len(self.calls) - 1)
        return [paf[None].astype(np.float32), heat[None].astype(np.float32)]
