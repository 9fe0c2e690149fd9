"""Drop-in for py_rmpe_server/py_rmpe_data_iterator.py (RawDataIterator :11-79): per-sample glue
around Transformer.transform + Heatmapper.create_heatmaps, plus a batched transform_batch that
runs the whole batch through one C-ABI call.  The HDF5 file is opened with h5py when it is installed and with the
package's own reader (h5lite.py: contiguous u8 datasets + the 'meta' string attribute, which is all the reference
writes, training/generate_hdf5_coco2014.py:356-366) when it is not."""
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
            try:
                import h5py as _h5  # noqa: optional dependency
            except ImportError:
                from . import h5lite as _h5
            self.h5 = _h5.File(self.h5file, "r")
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

    def gen_batched(self, batch, dbg=False):
        """Same stream of (image, mask, labels, joints) tuples as gen(), produced `batch` samples per C-ABI call.
        Sources of different sizes share a call after padding to the largest with the border constants of the two
        warps (127 / 255, py_rmpe_transformer.py:90-91): cv2 substitutes exactly those values for every tap outside
        the source, so the padded warp is bit-identical to the unpadded one."""
        keys = list(self.datum.keys())
        if self.shuffle:
            random.shuffle(keys)
        for k0 in range(0, len(keys), batch):
            raw = [self.read_data(key) for key in keys[k0:k0 + batch]]
            for tpl in self.transform_batch(raw):
                yield tpl

    def transform_batch(self, raw):
        """raw: list of (img HxWx3 u8, mask HxW u8, meta) -> list of (image (3,368,368), mask, labels, joints)."""
        n = len(raw)
        H = max(r[0].shape[0] for r in raw)
        W = max(r[0].shape[1] for r in raw)
        imgs = np.full((n, H, W, 3), 127, np.uint8)
        masks = np.full((n, H, W), 255, np.uint8)
        P = max(max(np.asarray(r[2]['joints']).shape[0] for r in raw), 1)
        joints = np.zeros((n, P, 18, 3))
        joints[:, :, :, 2] = 2.0                      # padding persons are "absent"
        n_persons = np.zeros(n, np.int32)
        augs = [AugmentSelection.random() if self.augment else AugmentSelection.unrandom() for _ in raw]
        for i, (img, mask, meta) in enumerate(raw):
            imgs[i, :img.shape[0], :img.shape[1]] = img
            masks[i, :mask.shape[0], :mask.shape[1]] = mask
            j = np.asarray(meta['joints'], dtype=np.float64)
            joints[i, :j.shape[0]] = j
            n_persons[i] = j.shape[0]
        M = _batch.aug_affine([a.flip for a in augs], [a.degree for a in augs], [a.crop for a in augs],
                              [a.scale for a in augs], [r[2]['objpos'][0] for r in raw],
                              [r[2]['scale_provided'][0] for r in raw])
        res = _batch.gt_batch_host(imgs, masks, joints, n_persons, M, [1 if a.flip else 0 for a in augs], f64=True, chw=True)
        out = []
        for i, (_, _, meta) in enumerate(raw):
            if n_persons[i]:
                meta['joints'][:, :, :] = res["joints"][i, :n_persons[i]]
            out.append((res["img"][i], res["mask"][i], res["labels"][i], meta['joints']))
        return out

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
