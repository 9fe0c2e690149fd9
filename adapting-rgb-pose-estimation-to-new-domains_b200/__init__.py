"""rmpe_b200 -- B200-native (sm_100a) OpenPose target generation and inference decode, a drop-in
for the py_rmpe_server / eval decode path of GuruMulay/Adapting-RGB-Pose-Estimation-to-New-Domains.

Layout (mirrors the reference's module paths for the path that is replaced):
  py_rmpe_server/py_rmpe_config.py        RmpeGlobalConfig, TransformationParams, RmpeCocoConfig
  py_rmpe_server/py_rmpe_transformer.py   AugmentSelection, Transformer
  py_rmpe_server/py_rmpe_heatmapper.py    Heatmapper
  py_rmpe_server/py_rmpe_data_iterator.py RawDataIterator
  eval/eval_coco2014_multi_modes.py       process_single_scale, process_multi_scale
  util.py                                 padRightDownCorner
  batch.py                                batched host / device-resident entry points
  csrc/ + librmpe_b200.so                 the CUDA kernels and the C ABI (include/rmpe_b200.h)
"""
from . import _lib  # noqa: F401
from .py_rmpe_server.py_rmpe_config import RmpeGlobalConfig, TransformationParams, RmpeCocoConfig  # noqa: F401

__all__ = ["RmpeGlobalConfig", "TransformationParams", "RmpeCocoConfig"]
