"""Drop-in for the batching half of the reference's training/ds_generators.py (DataIteratorBase.gen
:31-106, DataIterator :189-218): the step right after the GT path.

`gen(n_stages)` yields `[x (B,368,368,3) u8, x1 (B,46,46,38), x2 (B,46,46,19)], [y1 (B,46,46,38),
y2 (B,46,46,19)] * n_stages` like the reference.  The per-sample transposes / repeats / concatenations
of the reference (:47-77) run as ONE pass of k_keras_batch over the whole batch (rmpe_keras_batch);
`FusedDataIterator` runs warp + mask + labels for the whole batch in ONE C call whose rasteriser writes the
NHWC tensors itself (RmpeGtBatchHost.out_vec_label ...): one host-to-device and one device-to-host trip per batch.  `DataGeneratorClient` (:109-186) is the ZMQ PULL side of py_rmpe_server/rmpe_server.py;
the wire format lives in rmpe_server.send_arrays / recv_arrays."""
import numpy as np

from .. import batch as _batch
from ..py_rmpe_server.py_rmpe_data_iterator import RawDataIterator
from ..py_rmpe_server.py_rmpe_transformer import AugmentSelection
from ..py_rmpe_server.rmpe_server import recv_arrays as _recv_arrays_wire


class DataIteratorBase:

    def __init__(self, batch_size=10):
        self.batch_size = batch_size
        self.split_point = 38
        self.vec_num = 38
        self.heat_num = 19
        self.keypoints = [None] * self.batch_size  # not passed to the NN; read by accuracy code

    def _recv_arrays(self):
        raise NotImplementedError

    def gen_raw(self):
        while True:
            try:
                arrays = self._recv_arrays()
            except StopIteration:      # `stop` header / limit reached: end the stream (a bare StopIteration inside a
                return                 # generator is a RuntimeError since PEP 479)
            yield tuple(arrays)

    def gen(self, n_stages):
        imgs, masks, labels = [], [], []
        for foo in self.gen_raw():
            if len(foo) == 4:
                data_img, mask_img, label, kpts = foo
            else:
                data_img, mask_img, label = foo
                kpts = None
            imgs.append(np.transpose(data_img, (1, 2, 0)))
            masks.append(mask_img)
            labels.append(label)
            self.keypoints[len(imgs) - 1] = kpts
            if len(imgs) == self.batch_size:
                kb = _batch.keras_batch_host(np.stack(labels), np.stack(masks))
                batch_x = np.stack(imgs)
                imgs, masks, labels = [], [], []
                yield [batch_x, kb["x1"], kb["x2"]], [kb["y1"], kb["y2"]] * n_stages
                self.keypoints = [None] * self.batch_size


class DataGeneratorClient(DataIteratorBase):
    """ZMQ client of rmpe_server.Server (reference :109-186): same constructor, same `stop` / `limit` behaviour."""

    def __init__(self, host, port, hwm=20, batch_size=10, limit=None):
        super(DataGeneratorClient, self).__init__(batch_size)
        import zmq
        self.limit = limit
        self.records = 0
        self.host = host
        self.port = port
        self.hwm = hwm
        context = zmq.Context()
        self.socket = context.socket(zmq.PULL)
        self.socket.set_hwm(self.hwm)
        self.socket.connect("tcp://{}:{}".format(self.host, self.port))

    def _recv_arrays(self):
        if self.limit is not None and self.records > self.limit:
            raise StopIteration
        arrays = _recv_arrays_wire(self.socket)
        self.records += 1
        return arrays


class DataIterator(DataIteratorBase):
    """In-process iterator over an HDF5 file, sample by sample like the reference (:189-218)."""

    def __init__(self, file, shuffle=True, augment=True, batch_size=10, limit=None):
        super(DataIterator, self).__init__(batch_size)
        self.limit = limit
        self.records = 0
        self.raw_data_iterator = RawDataIterator(file, shuffle=shuffle, augment=augment)
        self.generator = self.raw_data_iterator.gen()

    def _recv_arrays(self):
        while True:
            if self.limit is not None and self.records > self.limit:
                raise StopIteration
            tpl = next(self.generator, None)
            if tpl is not None:
                self.records += 1
                return tpl
            if self.limit is None or self.records < self.limit:
                print("Staring next generator loop cycle")
                self.generator = self.raw_data_iterator.gen()
            else:
                raise StopIteration


class FusedDataIterator(DataIteratorBase):
    """Batched superset: `source` yields raw (img HxWx3 u8, mask HxW u8, meta) triples of one
    geometry (what RawDataIterator.read_data returns); a whole batch goes through ONE
    rmpe_gt_batch_host call: the planar (57,46,46) labels never exist, the rasteriser emits y1 / y2 / x1 / x2."""

    def __init__(self, source, augment=True, batch_size=10):
        super(FusedDataIterator, self).__init__(batch_size)
        self.source = source
        self.augment = augment

    def gen(self, n_stages):
        buf = []
        for img, mask, meta in self.source:
            buf.append((np.asarray(img), np.asarray(mask), meta))
            if len(buf) < self.batch_size:
                continue
            augs = [AugmentSelection.random() if self.augment else AugmentSelection.unrandom() for _ in buf]
            P = max(np.asarray(m['joints']).shape[0] for _, _, m in buf)
            joints = np.zeros((len(buf), max(P, 1), 18, 3))
            joints[:, :, :, 2] = 2.0                     # padding persons are "absent"
            n_persons = np.zeros(len(buf), np.int32)
            for i, (_, _, m) in enumerate(buf):
                j = np.asarray(m['joints'], dtype=np.float64)
                joints[i, :j.shape[0]] = j
                n_persons[i] = j.shape[0]
            M = _batch.aug_affine([a.flip for a in augs], [a.degree for a in augs], [a.crop for a in augs],
                                  [a.scale for a in augs], [m['objpos'][0] for _, _, m in buf],
                                  [m['scale_provided'][0] for _, _, m in buf])
            r = _batch.gt_batch_host(np.stack([b[0] for b in buf]), np.stack([b[1] for b in buf]), joints, n_persons, M,
                                     [1 if a.flip else 0 for a in augs], f64=True, want_labels=False, keras=True)
            for i, (_, _, m) in enumerate(buf):
                if n_persons[i]:
                    m['joints'][:, :, :] = r["joints"][i, :n_persons[i]]
                self.keypoints[i] = m['joints']
            buf = []
            yield [r["img"], r["x1"], r["x2"]], [r["y1"], r["y2"]] * n_stages
            self.keypoints = [None] * self.batch_size
