"""Property-based tests (hypothesis) of the wire codecs: what a node reads from a socket is untrusted, so ANY byte string must either decode
or be rejected with an exception; valid messages round-trip exactly; the gradient codecs keep their contracts on arbitrary inputs."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from geomx_b200 import runtime
from geomx_b200.kvstore import compression as gc

pytestmark = pytest.mark.skipif(not runtime.available(), reason="native runtime not built")
i32 = st.integers(-2**31, 2**31 - 1)


@settings(max_examples=400, deadline=None)
@given(st.binary(max_size=600))
def test_unpack_meta_never_crashes_on_garbage(blob):
    ok, again = runtime.C().fuzz_unpack_meta(blob)
    if ok:                                   # whatever decoded re-encodes canonically (checked natively) and decodes again
        assert runtime.C().fuzz_unpack_meta(again)[0]


@settings(max_examples=200, deadline=None)
@given(st.binary(min_size=4, max_size=200), st.integers(0, 199), st.integers(0, 255))
def test_unpack_meta_survives_bit_flips_of_valid_messages(body, pos, val):
    C = runtime.C()
    good, _ = C.pack_meta_fields(1, 2, 3, 4, 9, 101, True, True, False, body.decode("latin1"), -3, 7, [0.5, 2.0], 2, [(1, 100, "10.0.0.1", 9000)])
    bad = bytearray(good)
    bad[pos % len(bad)] = val
    C.fuzz_unpack_meta(bytes(bad))           # must return (True, ...) or (False, ...) — not raise, not crash
    huge = bytearray(good)                   # a length field blown up to 4 GiB must be rejected without allocating
    for off in range(0, len(huge) - 4, 4):
        probe = bytearray(good); probe[off:off + 4] = b"\xff\xff\xff\xff"
        C.fuzz_unpack_meta(bytes(probe))


@settings(max_examples=200, deadline=None)
@given(i32, i32, i32, i32, i32, i32, st.booleans(), st.booleans(), st.booleans(), st.text(max_size=64), i32, i32,
       st.lists(st.floats(width=32, allow_nan=False), max_size=4), st.sampled_from([0, 1, 2, 3, 4, 5]),
       st.lists(st.tuples(st.integers(0, 7), i32, st.text(alphabet="abc.0123456789", max_size=24), st.integers(0, 65535)), max_size=4))
def test_meta_roundtrip_is_exact(head, app, cust, ts, snd, rcv, req, push, simple, body, prio, key, compr, cmd, nodes):
    _, back = runtime.C().pack_meta_fields(head, app, cust, ts, snd, rcv, req, push, simple, body, prio, key, compr, cmd, nodes)
    want_nodes = nodes if cmd != 0 else []   # nodes only travel with a control command
    assert tuple(back[:12]) == (head, app, cust, ts, snd, rcv, req, push, simple, body, prio, key)
    assert list(back[12]) == pytest.approx(compr) and back[13] == cmd and [tuple(n) for n in back[14]] == [tuple(n) for n in want_nodes]


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 300), st.floats(0.05, 4.0), st.integers(0, 2**31 - 1))
def test_2bit_codec_contract(n, thr, seed):
    g = torch.from_numpy(np.random.RandomState(seed).randn(n).astype(np.float32) * 2)
    r = torch.zeros(n)
    q = gc.quantize_2bit(g.clone(), r, thr)
    d = gc.dequantize_2bit(q, n, thr)
    assert set(np.unique(d.numpy())).issubset({-np.float32(thr), 0.0, np.float32(thr)})
    assert torch.allclose(d + r, g, atol=1e-5)                        # nothing is lost: sent value + residual = gradient
    assert float(r.abs().max()) <= max(float(g.abs().max()) - thr, thr) + 1e-5


@settings(max_examples=60, deadline=None)
@given(st.integers(64, 4000), st.floats(0.005, 0.2), st.integers(0, 2**31 - 1))
def test_bsc_codec_contract(n, ratio, seed):
    rs = np.random.RandomState(seed)
    g = torch.from_numpy(rs.randn(n).astype(np.float32))
    u, v = torch.zeros(n), torch.zeros(n)
    out = gc.bsc_compress(g, u, v, ratio)
    k = out.numel() // 2
    if k == 0:                               # int(n * ratio) == 0 (the reference's truncation): nothing travels, the gradient stays in the accumulators
        assert int(n * ratio) == 0 and torch.allclose(v, g) and torch.allclose(u, g)
        return
    vals, idx = out[:k], out[k:]
    live = idx >= 0
    ids = idx[live].long()
    assert bool((ids[1:] > ids[:-1]).all()) and bool((ids < n).all())        # index-ordered, in range, sentinel-padded
    dense = gc.bsc_decompress(out, n)
    assert torch.equal(dense[ids], vals[live]) and float(dense.abs().sum()) == pytest.approx(float(vals[live].abs().sum()), rel=1e-5)
    assert float(v[ids].abs().max()) == 0.0 if ids.numel() else True                      # error feedback cleared exactly where values were sent


@settings(max_examples=500, deadline=None)
@given(i32, i32, i32, i32, st.sampled_from([0, 4, 32]), st.integers(0, 4), st.integers(0, 70000), st.integers(0, 3))
def test_dgt_reassembly_rejects_inconsistent_blocks(seq, seq_end, total, val_bytes, bits, ncompr, payload, nparts):
    """A DGT block header comes off the wire (UDP): offsets / sizes that do not fit the tensor they claim must be dropped, not written."""
    done = runtime.C().fuzz_dgt_block(seq, seq_end, total, val_bytes, bits, ncompr, payload, nparts)
    if done:                                          # a tensor was completed: then the block was the last one and fitted
        assert seq == seq_end and 0 < total and 0 <= val_bytes <= total and nparts >= 2


def test_dgt_reassembly_accepts_a_valid_tensor():
    C = runtime.C()
    # block size = DGT_BLOCK_SIZE (4096 bytes by default): a 2-block tensor of 6000 bytes
    assert C.fuzz_dgt_block(0, 1, 6000, 4096, 32, 0, 4096, 2, reset=True) is False      # reset: the property test above leaves partial tensors behind
    assert C.fuzz_dgt_block(1, 1, 6000, 1904, 32, 0, 1904, 2) is True
    assert C.fuzz_dgt_block(1, 1, 6000, 4096, 32, 0, 4096, 2) is False      # would run past the end of the tensor


@settings(max_examples=150, deadline=None)
@given(st.binary(max_size=400), st.integers(0, 500))
def test_native_record_reader_on_corrupt_files(tmp_path_factory, blob, off):
    """A damaged / hostile .rec file: scanning and reading either work or raise — lengths in chunk headers are checked against the file size."""
    from geomx_b200 import recordio
    import struct
    d = tmp_path_factory.mktemp("rec")
    p = str(d / "x.rec")
    good = struct.pack("<II", 0xced7230a, 5) + b"hello\x00\x00\x00"
    with open(p, "wb") as f:
        f.write(good + blob)
    r = recordio.RecordReader(p)
    try:
        offs = r.offsets
    except Exception:
        offs = [0]
    assert r.read(0) == b"hello"
    for o in list(offs)[:4] + [off]:
        try:
            r.read(o)
        except Exception:
            pass


_PRED_JSON = None


def _pred_fixture():
    global _PRED_JSON
    if _PRED_JSON is None:
        import geomx_b200 as mx
        s = mx.sym
        d = s.Variable("data")
        c = s.Convolution(d, kernel=(3, 3), num_filter=4, pad=(1, 1), name="c")
        p = s.Pooling(s.Activation(s.BatchNorm(c, name="bn"), "relu"), kernel=(2, 2), pool_type="avg")
        net = s.SoftmaxOutput(s.FullyConnected(s.Concat(s.Flatten(p), s.Flatten(p) * 2.0, dim=1), num_hidden=5, name="fc"), name="sm")
        shapes = {"data": (2, 3, 8, 8)}
        arg_shapes, _, aux_shapes = net.infer_shape(**shapes)
        import numpy as np
        import os
        import tempfile
        save = {"arg:" + n: mx.nd.array(np.full(sh, 0.1, np.float32)) for n, sh in zip(net.list_arguments(), arg_shapes) if n not in ("data", "sm_label")}
        save.update({"aux:" + n: mx.nd.array(np.ones(sh, np.float32)) for n, sh in zip(net.list_auxiliary_states(), aux_shapes)})
        f = os.path.join(tempfile.mkdtemp(), "f.params")
        mx.nd.save(f, save)
        _PRED_JSON = (net.tojson(nnvm=True), net.tojson(), open(f, "rb").read(), shapes)
    return _PRED_JSON


@settings(max_examples=300, deadline=None)
@given(st.data())
def test_native_predictor_on_hostile_graphs_and_params(data):
    """Symbol JSON and parameter files are user-supplied deployment artefacts: mutated graphs (attribute values, input references, operator
    names), truncated / bit-flipped JSON and parameter blobs either load and run or raise MXNetError — no crash, no hang, no huge allocation."""
    import json
    import numpy as np
    from geomx_b200 import predictor
    from geomx_b200.base import MXNetError
    nn, ours, blob, shapes = _pred_fixture()
    mode = data.draw(st.sampled_from(["attr", "ref", "bytes", "params", "op"]))
    text, pb = data.draw(st.sampled_from([nn, ours])), blob
    if mode in ("attr", "ref", "op"):
        d = json.loads(text)
        n = data.draw(st.sampled_from([x for x in d["nodes"] if x["op"] != "null"]))
        if mode == "attr":
            at = n.setdefault("attrs", {})
            key = data.draw(st.sampled_from(sorted(at) + ["kernel", "stride", "pad", "axis", "dim", "num_hidden", "shape"]))
            at[key] = data.draw(st.one_of(st.integers(-5, 70000), st.sampled_from(["(0, 0)", "(-1, 3)", "None", "nan", "", "(99999, 99999)", "True", [0, 0], [-2, 1 << 40], None, 1e30]),
                                          st.text(max_size=8)))
        elif mode == "ref":
            if n["inputs"]:
                i = data.draw(st.integers(0, len(n["inputs"]) - 1))
                v = data.draw(st.integers(-3, len(d["nodes"]) + 3))
                n["inputs"][i] = [v, data.draw(st.integers(0, 2)), 0] if isinstance(n["inputs"][i], list) else v
        else:
            n["op"] = data.draw(st.sampled_from(["Pooling", "Convolution", "transpose", "Reshape", "Concat", "BatchNorm", "Embedding", "softmax", "_group", "_item", "x"]))
        text = json.dumps(d)
    elif mode == "bytes":
        b = bytearray(text.encode())
        for _ in range(data.draw(st.integers(1, 4))):
            b[data.draw(st.integers(0, len(b) - 1))] = data.draw(st.integers(32, 126))
        text = bytes(b[:data.draw(st.integers(1, len(b)))]).decode("latin-1") if data.draw(st.booleans()) else b.decode("latin-1")
    else:
        b = bytearray(pb)
        for _ in range(data.draw(st.integers(1, 4))):
            b[data.draw(st.integers(0, min(len(b) - 1, 400)))] = data.draw(st.integers(0, 255))
        pb = bytes(b[:data.draw(st.integers(0, len(b)))]) if data.draw(st.booleans()) else bytes(b)
    try:
        p = predictor.Predictor(text, pb, shapes)
        p.forward(data=np.ones(shapes["data"], np.float32))
        for i in range(p.num_outputs):
            if int(np.prod(p.get_output_shape(i))) < (1 << 24):
                p.get_output(i)
    except MXNetError:
        pass


@settings(max_examples=300, deadline=None)
@given(st.data())
def test_native_params_parser_on_corrupt_files(data):
    """The C++ `.params` parser (C API GXNDArrayLoad / GXNDList*, predictor parameters): dense, row_sparse and csr records with flipped bytes
    or cut short either parse or raise — ranks, extents, dtype flags and sparse indices are validated before any allocation or copy."""
    import numpy as np
    import geomx_b200 as mx
    from geomx_b200 import predictor
    from geomx_b200.base import MXNetError
    from geomx_b200.ndarray.utils import save_bytes
    global _SPARSE_BLOB
    try:
        blob = _SPARSE_BLOB
    except NameError:
        dense = mx.nd.array(np.arange(12, dtype=np.float32).reshape(3, 4))
        rsp = mx.nd.sparse.row_sparse_array((np.ones((2, 4), np.float32), np.array([0, 2])), shape=(3, 4))
        csr = mx.nd.sparse.csr_matrix((np.array([1., 2., 3.], np.float32), np.array([0, 2, 1]), np.array([0, 1, 2, 3])), shape=(3, 4))
        blob = _SPARSE_BLOB = save_bytes({"d": dense, "r": rsp, "c": csr})
        got = predictor.load_ndarray_file(blob)
        assert got["r"].tolist() == [[1] * 4, [0] * 4, [1] * 4] and got["c"].tolist() == [[1, 0, 0, 0], [0, 0, 2, 0], [0, 3, 0, 0]]
    b = bytearray(blob)
    for _ in range(data.draw(st.integers(1, 6))):
        b[data.draw(st.integers(0, len(b) - 1))] = data.draw(st.integers(0, 255))
    if data.draw(st.booleans()):
        b = b[:data.draw(st.integers(0, len(b)))]
    try:
        predictor.load_ndarray_file(bytes(b))
    except MXNetError:
        pass
