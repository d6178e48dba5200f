""".params byte format (reference src/ndarray/ndarray.cc:1583-1811): golden bytes assembled by hand, V1 / legacy loaders, native vs python codec,
Block.save_parameters / load_parameters, save_checkpoint naming, trainer state pickles."""
import struct

import numpy as np
import pytest
import torch

import geomx_b200 as mx
from geomx_b200 import runtime
from geomx_b200.ndarray import utils as ndu


def golden(arrs, names, v=2):
    out = struct.pack("<QQQ", 0x112, 0, len(arrs))
    for a in arrs:
        if v == 2:
            out += struct.pack("<Ii", 0xF993FAC9, 0)
            out += struct.pack("<I", a.ndim) + struct.pack("<%dq" % a.ndim, *a.shape)
        elif v == 1:
            out += struct.pack("<I", 0xF993FAC8) + struct.pack("<I", a.ndim) + struct.pack("<%dq" % a.ndim, *a.shape)
        else:
            out += struct.pack("<I", a.ndim) + struct.pack("<%dI" % a.ndim, *a.shape)
        out += struct.pack("<ii", 1, 0)
        out += struct.pack("<i", {"float32": 0, "float64": 1, "float16": 2, "uint8": 3, "int32": 4, "int8": 5, "int64": 6}[a.dtype.name])
        out += a.tobytes()
    out += struct.pack("<Q", len(names))
    for n in names:
        out += struct.pack("<Q", len(n)) + n.encode()
    return out


def test_writer_is_byte_exact():
    a = np.arange(6, dtype=np.float32).reshape(2, 3); b = np.array([1, 2, 3], dtype=np.int32)
    blob = ndu.save_bytes({"arg:w": mx.nd.array(a), "aux:b": mx.nd.array(b, dtype="int32")})
    assert blob == golden([a, b], ["arg:w", "aux:b"])
    assert ndu.save_bytes([mx.nd.array(a)]) == golden([a], [])


@pytest.mark.parametrize("v", [2, 1, 0])
def test_loader_versions(v):
    a = np.random.rand(3, 4).astype(np.float32); h = np.random.rand(5).astype(np.float16)
    d = ndu.load_bytes(golden([a, h], ["x", "y"], v))
    assert np.array_equal(d["x"].asnumpy(), a) and np.array_equal(d["y"].asnumpy(), h) and d["y"].dtype == np.float16


def test_row_sparse_is_densified():
    shape, rows = (4, 2), np.array([1, 3], dtype=np.int64)
    data = np.array([[1, 2], [3, 4]], dtype=np.float32)
    blob = struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<Ii", 0xF993FAC9, 1)
    blob += struct.pack("<I", 2) + struct.pack("<2q", 2, 2)            # storage shape
    blob += struct.pack("<I", 2) + struct.pack("<2q", *shape)
    blob += struct.pack("<ii", 1, 0) + struct.pack("<i", 0)
    blob += struct.pack("<i", 6) + struct.pack("<I", 1) + struct.pack("<q", 2)   # aux type int64 + shape
    blob += data.tobytes() + rows.tobytes() + struct.pack("<Q", 0)
    out = ndu.load_bytes(blob)[0].asnumpy()
    assert np.array_equal(out, np.array([[0, 0], [1, 2], [0, 0], [3, 4]], dtype=np.float32))
    if runtime.available():
        arrays, names = runtime.C().params_load(blob)
        assert np.array_equal(np.frombuffer(arrays[0][2], dtype=np.float32).reshape(shape), out)


@pytest.mark.skipif(not runtime.available(), reason="native runtime not built")
def test_native_codec_matches_python():
    C = runtime.C()
    a = np.random.rand(7, 3).astype(np.float32); b = np.arange(4, dtype=np.int64)
    blob = golden([a, b], ["p", "q"])
    arrays, names = C.params_load(blob)
    assert names == ["p", "q"] and arrays[0][0] == 0 and list(arrays[0][1]) == [7, 3] and arrays[1][0] == 6
    assert np.array_equal(np.frombuffer(arrays[0][2], dtype=np.float32).reshape(7, 3), a)
    assert C.params_save([(0, [7, 3], a.tobytes(), 1, 0), (6, [4], b.tobytes(), 1, 0)], ["p", "q"]) == blob
    with pytest.raises(RuntimeError):
        C.params_load(b"\x00" * 24)


def test_block_save_load_and_checkpoint(tmp_path):
    net = mx.models.build_cnn(); net.initialize(init=mx.init.Xavier())
    net(mx.nd.zeros((2, 1, 28, 28)))
    f = str(tmp_path / "cnn.params")
    net.save_parameters(f)
    loaded = mx.nd.load(f)
    assert set(loaded) == {"%d.%s" % (i, n) for i in (0, 2, 4, 5, 6) for n in ("weight", "bias")}      # structural names
    net2 = mx.models.build_cnn(); net2.initialize()
    net2(mx.nd.zeros((2, 1, 28, 28)))
    net2.load_parameters(f)
    for p, q in zip(net.collect_params().values(), net2.collect_params().values()):
        assert np.array_equal(p.data().asnumpy(), q.data().asnumpy())
    args = {p.name: p.data() for p in net.collect_params().values()}
    mx.model.save_checkpoint(str(tmp_path / "ck"), 3, "{}", args, {})
    sym, arg2, aux2 = mx.model.load_checkpoint(str(tmp_path / "ck"), 3)
    assert (tmp_path / "ck-0003.params").exists() and (tmp_path / "ck-symbol.json").exists()
    assert set(arg2) == set(args) and aux2 == {}
    with pytest.raises(mx.base.MXNetError):
        ndu.load_bytes(b"garbage!" * 4)


def test_sparse_arrays_roundtrip_in_params_format(tmp_path):
    """row_sparse and csr arrays are written as V2 sparse records (stype 1 / 2 with storage shape + aux arrays, ndarray.cc:1583-1651) and read
    back as sparse objects; csr supports row slices and dot(csr, dense)."""
    import numpy as np
    import geomx_b200 as mx
    dense = np.zeros((6, 4), dtype=np.float32); dense[1] = [1, 0, 2, 0]; dense[4] = [0, 3, 0, 4]
    rs = mx.nd.array(dense).tostype("row_sparse"); csr = mx.nd.array(dense).tostype("csr")
    assert csr.indptr.asnumpy().tolist() == [0, 0, 2, 2, 2, 4, 4] and csr.indices.asnumpy().tolist() == [0, 2, 1, 3]
    f = str(tmp_path / "sparse.params")
    mx.nd.save(f, {"rs": rs, "csr": csr, "dense": mx.nd.array(dense)})
    back = mx.nd.load(f)
    assert back["rs"].stype == "row_sparse" and back["rs"].indices.asnumpy().tolist() == [1, 4] and np.array_equal(back["rs"].asnumpy(), dense)
    assert back["csr"].stype == "csr" and np.array_equal(back["csr"].asnumpy(), dense) and np.array_equal(back["dense"].asnumpy(), dense)
    assert np.array_equal(csr[1:5].asnumpy(), dense[1:5])
    w = np.arange(8, dtype=np.float32).reshape(4, 2)
    assert np.allclose(mx.nd.sparse.dot(csr, mx.nd.array(w)).asnumpy(), dense @ w)
    assert np.allclose(mx.nd.sparse.dot(csr, mx.nd.array(np.ones((6, 3), dtype=np.float32)), transpose_a=True).asnumpy(), dense.T @ np.ones((6, 3)))
