"""Drop-in for py_rmpe_server/rmpe_server.py (Server :13-89): a forked process that pushes augmented
samples to training clients over ZeroMQ PUSH/PULL, now fed by the GPU path.

Wire format (reference rmpe_server.py:69-89, read back by training/ds_generators.py:148-186): per
sample FIVE frames -- one JSON list of four `{"descr", "shape", "fortran_order"}` headers, then the raw
C-contiguous bytes of image (3,368,368) u8, mask (46,46) f64, labels (57,46,46) f64 and keypoints
(P,18,3) f64.  `send_arrays` / `recv_arrays` are that format as two functions so that the server, the
client (training/ds_generators.DataGeneratorClient) and the tests share one implementation; existing
clients of the reference server keep working against this one.

`Server(..., batch=N)` is the B200 superset: the child process reads N raw samples and sends them through
ONE rmpe_gt_batch_host call (RawDataIterator.gen_batched) instead of one call per sample; what goes on
the wire is unchanged (one 5-frame message per sample)."""
from ast import literal_eval as make_tuple
from multiprocessing import Process
from time import time

import numpy as np


def produce_headers(arrays):
    """rmpe_server.py:79-89: dtype string, shape and C order of every array of a sample."""
    return [{"descr": a.dtype.str, "shape": a.shape, "fortran_order": False} for a in arrays]


def send_arrays(socket, arrays):
    """One sample: JSON headers, then one frame per array (rmpe_server.py:67-73)."""
    arrays = [np.ascontiguousarray(a) for a in arrays]
    socket.send_json(produce_headers(arrays))
    for a in arrays:
        socket.send(a)


def recv_arrays(socket):
    """Inverse of send_arrays (ds_generators.py:148-186).  Raises StopIteration on a `stop` header; accepts the
    shape as a string (the C++ server of the reference sends it that way) and fortran_order."""
    headers = socket.recv_json()
    if 'stop' in headers:
        raise StopIteration
    arrays = []
    for header in headers:
        data = socket.recv()
        array = np.frombuffer(memoryview(data), dtype=np.dtype(header['descr']))
        shape = make_tuple(header['shape']) if isinstance(header['shape'], str) else header['shape']
        if header['fortran_order']:
            array = array.reshape(tuple(shape)[::-1]).transpose()
        else:
            array = array.reshape(tuple(shape))
        arrays.append(array)
    return arrays


class Server:

    # these methods all called in parent process

    def __init__(self, h5file, port, name, shuffle, augment, batch=1, hwm=1, iterator_factory=None):
        self.name = name
        self.port = port
        self.h5file = h5file
        self.shuffle = shuffle
        self.augment = augment
        self.batch = batch
        self.hwm = hwm
        self.iterator_factory = iterator_factory
        self.process = Process(target=Server.loop, args=(self,))
        self.process.daemon = True
        self.process.start()

    def join(self):
        return self.process.join(10)

    # these methods all called in child process (the CUDA context is created there, after the fork)

    def init(self):
        import zmq
        self.context = zmq.Context()
        self.socket = self.context.socket(zmq.PUSH)
        self.socket.set_hwm(self.hwm)
        self.socket.bind("tcp://*:%s" % self.port)

    def make_iterator(self):
        if self.iterator_factory is not None:
            return self.iterator_factory()
        from .py_rmpe_data_iterator import RawDataIterator
        return RawDataIterator(self.h5file, shuffle=self.shuffle, augment=self.augment)

    @staticmethod
    def loop(self):
        print("%s: Child process init... " % self.name)
        self.init()
        iterator = self.make_iterator()
        print("%s: Loop started... " % self.name)
        num = 0
        generation = 0
        while True:
            keys = iterator.num_keys()
            print("%s: generation %s, %d images " % (self.name, generation, keys))
            start = time()
            gen = iterator.gen_batched(self.batch) if self.batch > 1 and hasattr(iterator, "gen_batched") else iterator.gen()
            for (image, mask, labels, keypoints) in gen:
                augment_time = time() - start
                send_arrays(self.socket, (image, mask, labels, keypoints))
                num += 1
                print("%s [%d/%d] aug %0.2f ms (%0.2f im/s), send %0.2f s" % (
                    self.name, num, keys, augment_time * 1000, 1. / max(augment_time, 1e-9), time() - start - augment_time))
                start = time()
            generation += 1

    def produce_headers(self, img, mask, labels, keypoints):
        return produce_headers((img, mask, labels, keypoints))


def main():
    train = Server("../dataset/train_dataset.h5", 5555, "Train", shuffle=False, augment=True)
    val = Server("../dataset/val_dataset.h5", 5556, "Val", shuffle=False, augment=False)
    processes = [val, train]
    while None in [p.process.exitcode for p in processes]:
        print("exitcodes", [p.process.exitcode for p in processes])
        for p in processes:
            if p.process.exitcode is None:
                p.join()


if __name__ == "__main__":
    np.set_printoptions(precision=1, linewidth=100 * 3, suppress=True, threshold=100000)
    main()
