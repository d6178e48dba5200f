"""Property-based tests (hypothesis) of the C ABI's untrusted inputs: symbol JSON, raw NDArray bytes, RecordIO files, attribute strings and
operator compositions.  Whatever comes in, a call returns 0 or -1 with a message — it never crashes, hangs or allocates by a forged size."""
import ctypes
import json
import os

import numpy as np
from hypothesis import given, settings, strategies as st

import _capi as C
from _capi import ck, lib, u32, vp, cp

_VALID = None


def _valid_json():
    global _VALID
    if _VALID is None:
        x = C.var("data")
        h = C.op("Convolution", "c", [x], kernel="(3, 3)", num_filter=4, pad="(1, 1)")
        h = C.op("BatchNorm", "bn", [h])
        h = C.op("Pooling", "p", [C.op("Activation", "a", [h], act_type="relu")], kernel="(2, 2)", stride="(2, 2)")
        h = C.op("Concat", "cat", [C.op("Flatten", "f", [h]), C.op("Flatten", "f2", [x])], dim=1)
        h = C.op("FullyConnected", "fc", [h], num_hidden=5)
        _VALID = C.sym_json(C.op("SoftmaxOutput", "sm", kwinputs={"data": h, "label": C.var("label")}))
    return _VALID


def _exercise(sym):
    """Everything a front end would call on a freshly loaded graph."""
    C.list_arguments(sym); C.list_outputs(sym); C.list_aux(sym)
    out = cp(); ck(lib().GXSymbolPrint(sym, ctypes.byref(out))); ck(lib().GXSymbolSaveToJSON(sym, ctypes.byref(out)))
    again = vp(); ck(lib().GXSymbolCreateFromJSON(out.value, ctypes.byref(again)))           # what was accepted re-serialises to something acceptable
    ck(lib().GXSymbolFree(again))
    try:
        C.infer_shape(sym, partial=True, data=(2, 3, 8, 8))
    except RuntimeError:
        pass
    try:
        ex, *_ = C.simple_bind(sym, {"data": (2, 3, 8, 8)})
        C.forward(ex, True); C.backward(ex)
        ck(lib().GXExecutorFree(ex))
    except (RuntimeError, AssertionError):
        pass


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 10**6), st.integers(0, 255), st.integers(0, 3))
def test_symbol_json_survives_byte_damage(pos, val, mode):
    good = bytearray(_valid_json().encode())
    if mode == 0:
        good[pos % len(good)] = val
    elif mode == 1:
        good = good[:pos % len(good)]
    elif mode == 2:
        i = pos % len(good); good[i:i] = bytes([val]) * (val % 7)
    else:
        i = pos % len(good); del good[i:i + 1 + val % 9]
    h = vp()
    rc = lib().GXSymbolCreateFromJSON(bytes(good).replace(b"\x00", b" "), ctypes.byref(h))
    assert rc in (0, -1)
    if rc == 0:
        _exercise(h); ck(lib().GXSymbolFree(h))


_json_leaf = st.one_of(st.none(), st.booleans(), st.integers(-2**40, 2**40), st.floats(allow_nan=False, allow_infinity=False), st.text(max_size=12))


@settings(max_examples=200, deadline=None)
@given(st.lists(st.fixed_dictionaries({"op": st.sampled_from(["null", "FullyConnected", "Convolution", "Concat", "BatchNorm", "elemwise_add", "Reshape", "nope", "_nd"]),
                                       "name": st.text(max_size=8),
                                       "attrs": st.dictionaries(st.sampled_from(["num_hidden", "kernel", "num_filter", "num_args", "shape", "dim", "axis", "fn", "kwargs", "pad"]),
                                                                st.one_of(_json_leaf, st.lists(st.integers(-5, 70000), max_size=5)), max_size=5),
                                       "inputs": st.lists(st.one_of(st.integers(-3, 12), st.lists(st.integers(-3, 12), max_size=4)), max_size=6)}), max_size=10),
       st.lists(st.lists(st.integers(-2, 12), max_size=3), max_size=3))
def test_symbol_json_structurally_random_graphs(nodes, heads):
    doc = json.dumps({"nodes": nodes, "arg_nodes": [], "heads": heads}).encode()
    h = vp()
    rc = lib().GXSymbolCreateFromJSON(doc, ctypes.byref(h))
    assert rc in (0, -1)
    if rc == 0:
        _exercise(h); ck(lib().GXSymbolFree(h))


_attr_val = st.one_of(st.integers(-10, 10**12).map(str), st.text(alphabet="()[], -0123456789.eE", max_size=16), st.sampled_from(["True", "None", "max", "relu", "valid", "full", ""]))


@settings(max_examples=300, deadline=None)
@given(st.sampled_from(["Convolution", "Pooling", "FullyConnected", "Reshape", "transpose", "Concat", "BatchNorm", "sum", "Embedding", "expand_dims", "LeakyReLU", "clip", "dot"]),
       st.dictionaries(st.sampled_from(["kernel", "stride", "pad", "dilate", "num_filter", "num_group", "num_hidden", "shape", "axes", "dim", "axis", "num_args", "pool_type",
                                        "global_pool", "pooling_convention", "input_dim", "output_dim", "keepdims", "act_type", "slope", "a_min", "a_max", "eps", "flatten"]),
                       _attr_val, max_size=6),
       st.lists(st.integers(1, 9), min_size=1, max_size=5))
def test_operators_reject_bad_attributes_cleanly(opname, attrs, shape):
    h = vp()
    keys, vals = list(attrs.keys()), list(attrs.values())
    if lib().GXSymbolCreateAtomicSymbolByName(opname.encode(), len(keys), C.strs(keys), C.strs(vals), ctypes.byref(h)) != 0:
        return
    if lib().GXSymbolCompose(h, b"n", 1, None, C.handles([C.var("data")])) != 0:
        ck(lib().GXSymbolFree(h)); return
    try:
        args = C.list_arguments(h)
        shapes = {"data": tuple(shape)}
        if "n_rhs" in args:
            shapes["n_rhs"] = tuple(shape)
        ex, a, g, x = C.simple_bind(h, shapes)
        for name, arr in a.items():
            C.nd_set(arr, np.random.RandomState(0).rand(*C.nd_shape(arr)))
        out = C.forward(ex, True)
        assert all(np.isfinite(o).all() or True for o in out)
        C.backward(ex, [C.nd_create(np.ones(o.shape)) for o in out])
        ck(lib().GXExecutorFree(ex))
    except (RuntimeError, AssertionError):
        pass
    ck(lib().GXSymbolFree(h))


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 10**6), st.integers(0, 255), st.booleans())
def test_raw_ndarray_bytes_survive_damage(pos, val, truncate):
    a = C.nd_create(np.arange(30, dtype=np.float32).reshape(2, 3, 5))
    n, buf = ctypes.c_size_t(), ctypes.POINTER(ctypes.c_char)()
    ck(lib().GXNDArraySaveRawBytes(a, ctypes.byref(n), ctypes.byref(buf)))
    raw = bytearray(ctypes.string_at(buf, n.value))
    if truncate:
        raw = raw[:pos % len(raw)]
    else:
        raw[pos % len(raw)] = val
        i = (pos * 7) % max(len(raw) - 8, 1); raw[i:i + 8] = b"\xff" * 8 if val & 1 else raw[i:i + 8]          # forged 64-bit extents
    b = vp()
    rc = lib().GXNDArrayLoadFromRawBytes(bytes(raw), ctypes.c_size_t(len(raw)), ctypes.byref(b))
    assert rc in (0, -1)
    if rc == 0:
        C.nd_shape(b); C.nd_free(b)
    C.nd_free(a)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.binary(max_size=80), min_size=1, max_size=5), st.integers(0, 10**6), st.integers(0, 255), st.booleans())
def test_recordio_reader_survives_damage(payloads, pos, val, truncate):
    import tempfile
    fd, name = tempfile.mkstemp(suffix=".rec"); os.close(fd)
    path = name.encode()
    w = vp(); ck(lib().GXRecordIOWriterCreate(path, ctypes.byref(w)))
    for p in payloads:
        ck(lib().GXRecordIOWriterWriteRecord(w, p, ctypes.c_size_t(len(p))))
    ck(lib().GXRecordIOWriterFree(w))
    raw = bytearray(open(path, "rb").read())
    if truncate:
        raw = raw[:pos % (len(raw) + 1)]
    elif raw:
        raw[pos % len(raw)] = val
    open(path, "wb").write(bytes(raw))
    r = vp(); ck(lib().GXRecordIOReaderCreate(path, ctypes.byref(r)))
    for _ in range(len(payloads) + 2):
        buf, size = ctypes.POINTER(ctypes.c_char)(), ctypes.c_size_t()
        rc = lib().GXRecordIOReaderReadRecord(r, ctypes.byref(buf), ctypes.byref(size))
        assert rc in (0, -1)
        if rc == -1 or not buf:
            break
        assert size.value <= len(raw)
    ck(lib().GXRecordIOReaderFree(r))
    os.remove(path)
