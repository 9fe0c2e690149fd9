"""ZMQ wire format of the data server (reference py_rmpe_server/rmpe_server.py:65-89) and its client
(training/ds_generators.py:109-186): CPU-only, loopback on 127.0.0.1."""
import json
import socket as pysocket
import time

import numpy as np
import pytest

zmq = pytest.importorskip("zmq")


def _free_port():
    s = pysocket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _sample(seed, P=2):
    rng = np.random.RandomState(seed)
    return (rng.randint(0, 256, (3, 368, 368)).astype(np.uint8), rng.rand(46, 46), rng.rand(57, 46, 46),
            rng.rand(P, 18, 3) * 368)


def _pair(port):
    ctx = zmq.Context.instance()
    push = ctx.socket(zmq.PUSH)
    push.bind("tcp://127.0.0.1:%d" % port)
    pull = ctx.socket(zmq.PULL)
    pull.setsockopt(zmq.RCVTIMEO, 20000)
    pull.connect("tcp://127.0.0.1:%d" % port)
    return push, pull


def test_headers_are_the_references(built_lib):
    srv = built_lib.sub("py_rmpe_server.rmpe_server")
    img, mask, labels, kp = _sample(0)
    # rmpe_server.py:79-89, after a JSON round trip (tuples become lists)
    expected = [{"descr": "|u1", "shape": [3, 368, 368], "fortran_order": False},
                {"descr": "<f8", "shape": [46, 46], "fortran_order": False},
                {"descr": "<f8", "shape": [57, 46, 46], "fortran_order": False},
                {"descr": "<f8", "shape": [2, 18, 3], "fortran_order": False}]
    assert json.loads(json.dumps(srv.produce_headers((img, mask, labels, kp)))) == expected


def test_send_recv_round_trip(built_lib):
    srv = built_lib.sub("py_rmpe_server.rmpe_server")
    push, pull = _pair(_free_port())
    try:
        for seed in range(3):
            s = _sample(seed, P=seed)           # P = 0: an empty keypoint array still travels
            srv.send_arrays(push, s)
            r = srv.recv_arrays(pull)
            assert len(r) == 4
            for a, b in zip(s, r):
                assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
        # non-contiguous inputs go out C-contiguous, like np.ascontiguousarray in the reference (:69-73)
        hwc = np.transpose(_sample(5)[0], (1, 2, 0))
        srv.send_arrays(push, (hwc,))
        assert np.array_equal(srv.recv_arrays(pull)[0], hwc)
    finally:
        push.close(0)
        pull.close(0)


def test_client_parses_cpp_server_headers_and_stop(built_lib):
    """The reference's C++ server sends the shape as a string; fortran_order transposes (ds_generators.py:172-180)."""
    dsg = built_lib.sub("training.ds_generators")
    port = _free_port()
    ctx = zmq.Context.instance()
    push = ctx.socket(zmq.PUSH)
    push.bind("tcp://127.0.0.1:%d" % port)
    client = dsg.DataGeneratorClient("127.0.0.1", port, hwm=4, batch_size=2)
    client.socket.setsockopt(zmq.RCVTIMEO, 20000)
    try:
        a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
        f = np.asfortranarray(np.arange(6, dtype=np.int32).reshape(2, 3))
        push.send_json([{"descr": "<f4", "shape": "(2, 3, 4)", "fortran_order": False},
                        {"descr": "<i4", "shape": [2, 3], "fortran_order": True}])
        push.send(a)
        push.send(f.tobytes(order="F"))
        got = client._recv_arrays()
        assert np.array_equal(got[0], a) and got[0].shape == (2, 3, 4)
        assert np.array_equal(got[1], f)
        push.send_json({"stop": True})
        with pytest.raises(StopIteration):
            client._recv_arrays()
        push.send_json({"stop": True})
        assert list(client.gen_raw()) == []        # the stream ends cleanly
    finally:
        push.close(0)
        client.socket.close(0)


class _FakeIterator:
    """Stands in for RawDataIterator in the forked server: three fixed samples per generation."""

    def num_keys(self):
        return 3

    def gen(self):
        for seed in range(3):
            yield _sample(100 + seed)


def test_server_process_feeds_client_batches(built_lib):
    """Server (forked child, PUSH) -> DataGeneratorClient (PULL) -> DataIteratorBase.gen batches; the batch assembly
    itself needs the GPU, so only the raw stream is checked here."""
    srv = built_lib.sub("py_rmpe_server.rmpe_server")
    dsg = built_lib.sub("training.ds_generators")
    port = _free_port()
    server = srv.Server(None, port, "Test", shuffle=False, augment=False, hwm=8, iterator_factory=_FakeIterator)
    client = dsg.DataGeneratorClient("127.0.0.1", port, hwm=8, batch_size=3)
    client.socket.setsockopt(zmq.RCVTIMEO, 30000)
    try:
        raw = client.gen_raw()
        for i in range(6):                      # two generations of the fake iterator
            got = next(raw)
            want = _sample(100 + i % 3)
            assert len(got) == 4
            for a, b in zip(want, got):
                assert a.dtype == b.dtype and np.array_equal(a, b)
        assert client.records == 6
    finally:
        server.process.terminate()
        server.process.join(5)
        client.socket.close(0)


def test_comparator_metrics_are_the_references():
    """rmpe_server_comparator's L1 / L2 / exact-match metrics (reference :19-111) on two synthetic sample streams, and
    the weights.tsv layout (6 + 57 * 3 columns)."""
    import rmpe_b200
    cmpm = rmpe_b200.sub("py_rmpe_server.rmpe_server_comparator")
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, size=(3, 368, 368)).astype(np.uint8)
    mask = rng.randint(0, 256, size=(46, 46)) / 255.0
    lab = rng.uniform(-1, 1, size=(57, 46, 46))
    img2 = img.copy(); img2[0, 0, :4] ^= 1
    mask2 = mask.copy(); mask2[1, 1] += 1 / 255.0
    lab2 = lab.copy(); lab2[5] += 2 / 255.0
    rows = cmpm.step(0, {"a": (img, mask, lab), "b": (img2, mask2, lab2), "c": (img, mask, lab)})
    assert len(rows) == 3 and all(len(r) == 6 + 57 * 3 for r in rows)
    ab, ac, bc = rows
    assert ac[:6] == [0.0, 0.0, 1.0, 0.0, 0.0, 1.0] and all(v == 1.0 for v in ac[8::3])
    n = img.size
    assert abs(ab[0] - 4.0 / n) < 1e-12 and abs(ab[2] - (1 - 4.0 / n)) < 1e-12            # four pixels differ by one
    assert abs(ab[3] - 1.0 / 2116) < 1e-9 and abs(ab[5] - (1 - 1.0 / 2116)) < 1e-12        # one mask cell by 1/255
    cols = cmpm.columns()
    assert abs(ab[cols.index("Layer5L1")] - 2.0) < 1e-9 and ab[cols.index("Layer5AC")] == 0.0
    assert ab[cols.index("Layer6L1")] == 0.0 and ab == bc
    import tempfile, os
    p = os.path.join(tempfile.mkdtemp(), "weights.tsv")
    cmpm.write_tsv(rows, p)
    lines = open(p).read().splitlines()
    assert len(lines) == 4 and lines[0].split("\t")[1:] == cols and len(lines[1].split("\t")) == 1 + len(cols)
