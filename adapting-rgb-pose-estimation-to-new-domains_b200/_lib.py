"""ctypes binding of include/rmpe_b200.h (librmpe_b200.so, built in-tree by
__graft_entry__.build()).  There is no CPU fallback: if the library is missing or no sm_100
device is present, every entry point raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librmpe_b200.so")

OK = 0
GT_IMG_CHW = 0x1
GT_LABELS_F64 = 0x2
GT_NO_TRANSFORM = 0x4
GT_NO_WARP = 0x8
GT_SIMPLE_KERNELS = 0x10
GT_WARP_ONLY = 0x20
GT_PAF_AVERAGE = 0x40

ST_ZERO_LIMB = 0x1
ST_PEAK_OVERFLOW = 0x2
ST_CAND_OVERFLOW = 0x4
ST_PERSON_OVERFLOW = 0x8
ST_FOUND_GT2 = 0x10
ST_SINGULAR = 0x20
ST_PERSONS_CLAMPED = 0x40
ABI_VERSION = 2

MAX_SCALES = 4
DECODE_REUSE_TABLES = 0x1

EXPORTS = [
    "rmpe_init", "rmpe_shutdown", "rmpe_last_error", "rmpe_abi_version", "rmpe_device",
    "rmpe_aug_affine", "rmpe_aug_random", "rmpe_gt_batch", "rmpe_gt_batch_host",
    "rmpe_decode_workspace_bytes", "rmpe_decode_batch", "rmpe_decode_batch_host",
    "rmpe_debug_heat_maps", "rmpe_debug_paf_points", "rmpe_pad_right_down_corner",
    "rmpe_launch_count", "rmpe_profile_enable", "rmpe_profile_reset", "rmpe_profile_count",
    "rmpe_profile_get", "rmpe_keras_batch", "rmpe_keras_batch_host", "rmpe_debug_assemble",
]

_vp = C.c_void_p


class SrcDesc(C.Structure):
    _fields_ = [("img_offset", C.c_int64), ("mask_offset", C.c_int64), ("height", C.c_int32),
                ("width", C.c_int32), ("img_pitch", C.c_int32), ("mask_pitch", C.c_int32)]


SRC_DESC_DTYPE = np.dtype([("img_offset", "<i8"), ("mask_offset", "<i8"), ("height", "<i4"),
                           ("width", "<i4"), ("img_pitch", "<i4"), ("mask_pitch", "<i4")])


class GtBatch(C.Structure):
    _fields_ = [("batch", C.c_int32), ("max_persons", C.c_int32), ("flags", C.c_int32),
                ("reserved", C.c_int32),
                ("src_img", _vp), ("src_mask", _vp), ("src_desc", _vp), ("joints", _vp),
                ("n_persons", _vp), ("M", _vp), ("flip", _vp),
                ("out_img", _vp), ("out_mask", _vp), ("out_labels", _vp), ("out_joints", _vp),
                ("out_count", _vp), ("status", _vp),
                ("sigma", C.c_double), ("thre", C.c_double),
                ("out_vec_label", _vp), ("out_heat_label", _vp), ("out_vec_weights", _vp), ("out_heat_weights", _vp)]


class GtBatchHost(C.Structure):
    _fields_ = [("batch", C.c_int32), ("max_persons", C.c_int32), ("flags", C.c_int32),
                ("src_height", C.c_int32), ("src_width", C.c_int32), ("reserved", C.c_int32),
                ("src_img", _vp), ("src_mask", _vp), ("joints", _vp), ("n_persons", _vp),
                ("M", _vp), ("flip", _vp),
                ("out_img", _vp), ("out_mask", _vp), ("out_labels", _vp), ("out_joints", _vp),
                ("out_count", _vp), ("status", _vp),
                ("sigma", C.c_double), ("thre", C.c_double),
                ("out_vec_label", _vp), ("out_heat_label", _vp), ("out_vec_weights", _vp), ("out_heat_weights", _vp)]


class KerasBatch(C.Structure):
    _fields_ = [("batch", C.c_int32), ("flags", C.c_int32), ("labels", _vp), ("mask", _vp),
                ("vec_weights", _vp), ("heat_weights", _vp), ("vec_label", _vp), ("heat_label", _vp)]


class FrameDesc(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("n_scales", C.c_int32),
                ("reserved", C.c_int32),
                ("grid_h", C.c_int32 * MAX_SCALES), ("grid_w", C.c_int32 * MAX_SCALES),
                ("pad_down", C.c_int32 * MAX_SCALES), ("pad_right", C.c_int32 * MAX_SCALES),
                ("heat_offset", C.c_int64 * MAX_SCALES), ("paf_offset", C.c_int64 * MAX_SCALES)]


FRAME_DESC_DTYPE = np.dtype([("height", "<i4"), ("width", "<i4"), ("n_scales", "<i4"), ("reserved", "<i4"),
                             ("grid_h", "<i4", (MAX_SCALES,)), ("grid_w", "<i4", (MAX_SCALES,)),
                             ("pad_down", "<i4", (MAX_SCALES,)), ("pad_right", "<i4", (MAX_SCALES,)),
                             ("heat_offset", "<i8", (MAX_SCALES,)), ("paf_offset", "<i8", (MAX_SCALES,))])
assert FRAME_DESC_DTYPE.itemsize == C.sizeof(FrameDesc) == 144
assert SRC_DESC_DTYPE.itemsize == C.sizeof(SrcDesc) == 32


class DecodeBatch(C.Structure):
    _fields_ = [("batch", C.c_int32), ("max_peaks", C.c_int32), ("max_cand", C.c_int32),
                ("max_persons", C.c_int32), ("stride", C.c_int32), ("flags", C.c_int32),
                ("thre1", C.c_double), ("thre2", C.c_double),
                ("heat", _vp), ("paf", _vp), ("frames", _vp), ("frames_host", _vp),
                ("candidate", _vp), ("n_peaks", _vp), ("connections", _vp), ("n_conn", _vp),
                ("limb_cand", _vp), ("n_limb_cand", _vp), ("subset", _vp), ("n_subset", _vp),
                ("status", _vp), ("workspace", _vp), ("workspace_bytes", C.c_size_t)]


class DecodeBatchHost(C.Structure):
    _fields_ = [("batch", C.c_int32), ("max_peaks", C.c_int32), ("max_cand", C.c_int32),
                ("max_persons", C.c_int32), ("stride", C.c_int32), ("flags", C.c_int32),
                ("thre1", C.c_double), ("thre2", C.c_double),
                ("heat", _vp), ("paf", _vp), ("heat_elems", C.c_size_t), ("paf_elems", C.c_size_t),
                ("frames", _vp),
                ("candidate", _vp), ("n_peaks", _vp), ("connections", _vp), ("n_conn", _vp),
                ("limb_cand", _vp), ("n_limb_cand", _vp), ("subset", _vp), ("n_subset", _vp),
                ("status", _vp)]


_lib = None
_inited_device = None


def load():
    """dlopen the library (works without a GPU: the CUDA runtime is linked statically and is
    only touched by rmpe_init)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "librmpe_b200.so is not built (%s missing): run `python -c 'import __graft_entry__ as g; "
            "g.build()'` -- this package has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.rmpe_init.argtypes = [C.c_int]
    lib.rmpe_init.restype = C.c_int
    lib.rmpe_shutdown.restype = None
    lib.rmpe_last_error.restype = C.c_char_p
    lib.rmpe_abi_version.restype = C.c_int
    lib.rmpe_device.restype = C.c_int
    lib.rmpe_launch_count.restype = C.c_int64
    lib.rmpe_aug_affine.argtypes = [C.c_int] + [_vp] * 7
    lib.rmpe_aug_affine.restype = C.c_int
    lib.rmpe_aug_random.argtypes = [C.c_int] + [_vp] * 5
    lib.rmpe_aug_random.restype = C.c_int
    lib.rmpe_gt_batch.argtypes = [C.POINTER(GtBatch), _vp]
    lib.rmpe_gt_batch.restype = C.c_int
    lib.rmpe_gt_batch_host.argtypes = [C.POINTER(GtBatchHost)]
    lib.rmpe_gt_batch_host.restype = C.c_int
    lib.rmpe_decode_workspace_bytes.argtypes = [C.c_int, _vp, C.c_int, C.c_int, C.c_int]
    lib.rmpe_decode_workspace_bytes.restype = C.c_size_t
    lib.rmpe_decode_batch.argtypes = [C.POINTER(DecodeBatch), _vp]
    lib.rmpe_decode_batch.restype = C.c_int
    lib.rmpe_decode_batch_host.argtypes = [C.POINTER(DecodeBatchHost)]
    lib.rmpe_decode_batch_host.restype = C.c_int
    lib.rmpe_debug_heat_maps.argtypes = [_vp, _vp, _vp, _vp, _vp]
    lib.rmpe_debug_heat_maps.restype = C.c_int
    lib.rmpe_debug_paf_points.argtypes = [_vp, _vp, C.c_int, _vp, _vp, _vp]
    lib.rmpe_debug_paf_points.restype = C.c_int
    lib.rmpe_pad_right_down_corner.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]
    lib.rmpe_pad_right_down_corner.restype = C.c_int
    lib.rmpe_profile_enable.argtypes = [C.c_int]
    lib.rmpe_profile_enable.restype = C.c_int
    lib.rmpe_profile_reset.restype = C.c_int
    lib.rmpe_profile_count.restype = C.c_int
    lib.rmpe_profile_get.argtypes = [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.rmpe_profile_get.restype = C.c_int
    lib.rmpe_keras_batch.argtypes = [C.POINTER(KerasBatch), _vp]
    lib.rmpe_keras_batch.restype = C.c_int
    lib.rmpe_keras_batch_host.argtypes = [C.POINTER(KerasBatch)]
    lib.rmpe_keras_batch_host.restype = C.c_int
    lib.rmpe_debug_assemble.argtypes = [C.c_int, C.c_int] + [_vp] * 8
    lib.rmpe_debug_assemble.restype = C.c_int
    lib.rmpe_debug_bicubic_table.argtypes = [_vp]
    lib.rmpe_debug_bicubic_table.restype = C.c_int
    _lib = lib
    return lib


class RmpeError(RuntimeError):
    pass


def check(rc):
    if rc != OK:
        raise RmpeError("rmpe_b200 error %d: %s" % (rc, load().rmpe_last_error().decode("utf-8", "replace")))


def ensure_init(device=None):
    """Initialise the library on `device` (default: $LOCAL_RANK or 0).  Raises when no B200 is
    visible -- the product path never computes on the CPU."""
    global _inited_device
    lib = load()
    if _inited_device is not None:
        return lib
    if device is None:
        device = int(os.environ.get("RMPE_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    check(lib.rmpe_init(int(device)))
    _inited_device = int(device)
    return lib


def profile_enable(on=True, reset=True):
    lib = load()
    if reset:
        check(lib.rmpe_profile_reset())
    check(lib.rmpe_profile_enable(1 if on else 0))


def profile_read():
    """{kernel name: (total device ms, launches)} accumulated since the last reset."""
    lib = load()
    out = {}
    for i in range(lib.rmpe_profile_count()):
        name = C.create_string_buffer(64)
        ms = C.c_double()
        n = C.c_int64()
        check(lib.rmpe_profile_get(i, name, 64, C.byref(ms), C.byref(n)))
        out[name.value.decode()] = (ms.value, n.value)
    return out


def ptr(a):
    """Address of a numpy array's (or torch tensor's) first element."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())


def c_contig(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a
