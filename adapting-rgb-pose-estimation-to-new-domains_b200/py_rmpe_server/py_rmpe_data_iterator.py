"""Drop-in for py_rmpe_server/py_rmpe_data_iterator.py (RawDataIterator :11-79): per-sample glue
around Transformer.transform + Heatmapper.create_heatmaps, plus a batched transform_batch that
runs the whole batch through one C-ABI call.  h5py is imported lazily (absent in this image)."""
import json
import random

import numpy as np

from .. import batch as _batch
from .py_rmpe_config import RmpeGlobalConfig, RmpeCocoConfig
from .py_rmpe_transformer import Transformer, AugmentSelection
from .py_rmpe_heatmapper import Heatmapper


class RawDataIterator:

    def __init__(self, h5file, shuffle=True, augment=True):
        self.h5file = h5file
        self.h5 = None
        self.datum = None
        if h5file is not None:
            import h5py  # noqa: deferred, optional dependency
            self.h5 = h5py.File(self.h5file, "r")
            self.datum = self.h5['datum']
        self.heatmapper = Heatmapper()
        self.augment = augment
        self.shuffle = shuffle

    def gen(self, dbg=False):
        keys = list(self.datum.keys())
        if self.shuffle:
            random.shuffle(keys)
        for key in keys:
            image, mask, meta = self.read_data(key)
            image, mask, meta, labels = self.transform_data(image, mask, meta)
            image = np.transpose(image, (2, 0, 1))
            yield image, mask, labels, meta['joints']

    def num_keys(self):
        return len(list(self.datum.keys()))

    def read_data(self, key):
        entry = self.datum[key]
        assert 'meta' in entry.attrs, "No 'meta' attribute in .h5 file. Did you generate .h5 with new code?"
        meta = json.loads(entry.attrs['meta'])
        meta['joints'] = RmpeCocoConfig.convert(np.array(meta['joints']))
        data = entry[()]
        if data.shape[0] <= 6:
            data = data.transpose([1, 2, 0])
        return data[:, :, 0:3], data[:, :, 4], meta

    def transform_data(self, img, mask, meta):
        """One sample through one fused call (warp + mask + joints + labels)."""
        aug = AugmentSelection.random() if self.augment else AugmentSelection.unrandom()
        M = aug.affine(meta['objpos'][0], meta['scale_provided'][0])
        joints = np.asarray(meta['joints'], dtype=np.float64)
        P = joints.shape[0]
        res = _batch.gt_batch_host(np.asarray(img)[None], np.asarray(mask)[None], joints[None], [P], M[None],
                                   [1 if aug.flip else 0], f64=True)
        if P:
            meta['joints'][:, :, :] = res["joints"][0]
        return res["img"][0], res["mask"][0], meta, res["labels"][0]

    def __del__(self):
        if getattr(self, "h5", None) is not None:
            self.h5.close()
