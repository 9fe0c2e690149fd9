"""Drop-in for the path-relevant part of the reference's util.py: padRightDownCorner (:57-77).
The notebook display / colour-map helpers of that file are outside the accelerated path."""
import ctypes as C

import numpy as np

from . import _lib as L


def padRightDownCorner(img, stride, padValue):
    """Pads bottom/right with `padValue` up to multiples of `stride`; returns
    (img_padded, [0, 0, pad_down, pad_right]); dtype preserved (u8 runs on the device)."""
    img = np.asarray(img)
    h, w = img.shape[0], img.shape[1]
    pad = [0, 0, 0 if (h % stride == 0) else stride - (h % stride),
           0 if (w % stride == 0) else stride - (w % stride)]
    if img.dtype != np.uint8 or img.ndim != 3:
        raise TypeError("padRightDownCorner: expected an (H,W,C) uint8 image")
    import torch
    lib = L.ensure_init()
    dev = torch.device("cuda", L._inited_device)
    src = torch.from_numpy(np.ascontiguousarray(img)).to(dev)
    dst = torch.empty((h + pad[2], w + pad[3], img.shape[2]), dtype=torch.uint8, device=dev)
    pad4 = (C.c_int * 4)()
    st = torch.cuda.current_stream(dev).cuda_stream
    L.check(lib.rmpe_pad_right_down_corner(L.ptr(src), h, w, img.shape[2], int(stride), int(padValue),
                                           L.ptr(dst), C.cast(pad4, C.c_void_p), C.c_void_p(st)))
    assert list(pad4) == pad
    return dst.cpu().numpy(), pad
