"""Operator namespaces and sub-module import paths of the reference that user code relies on: mx.nd.{linalg,image,op,_internal},
mx.sym.{linalg,random,sparse,image,op,contrib}, mxnet.{module,io,image,symbol}.<submodule>, optimizer.contrib, gluon.nn.activations,
gluon.model_zoo.model_store."""
import hashlib
import importlib
import os

import numpy as np
import pytest

import geomx_b200 as mx


def test_reference_import_paths_exist():
    for p in ["gluon.nn.activations", "gluon.model_zoo.model_store", "ndarray.image", "ndarray.linalg", "ndarray._internal", "ndarray.op",
              "symbol.linalg", "symbol.contrib", "symbol.random", "symbol.sparse", "symbol.image", "symbol.op", "symbol.symbol", "symbol._internal",
              "symbol.register", "module.base_module", "module.bucketing_module", "module.module", "module.python_module", "module.sequential_module",
              "module.executor_group", "io.io", "io.utils", "image.detection", "image.image", "optimizer.contrib", "optimizer.optimizer", "rnn.rnn_cell", "rnn.rnn", "rnn.io"]:
        importlib.import_module("geomx_b200." + p)
    from geomx_b200.image.detection import ImageDetIter
    from geomx_b200.image.image import ImageIter
    from geomx_b200.io.io import DataIter, NDArrayIter
    from geomx_b200.module.base_module import BaseModule
    from geomx_b200.module.module import Module
    from geomx_b200.symbol.symbol import Symbol
    assert issubclass(Module, BaseModule) and issubclass(NDArrayIter, DataIter) and issubclass(ImageDetIter, ImageIter) and Symbol is mx.sym.Symbol
    assert mx.mod.Module is Module and mx.io.NDArrayIter is NDArrayIter


def test_nd_linalg_namespace():
    rs = np.random.RandomState(0)
    a = rs.rand(2, 4, 4).astype(np.float32)
    spd = a @ a.transpose(0, 2, 1) + 4 * np.eye(4, dtype=np.float32)
    A = mx.nd.array(spd)
    L = mx.nd.linalg.potrf(A)
    assert np.allclose(mx.nd.linalg.gemm2(L, L, transpose_b=True).asnumpy(), spd, atol=1e-4)
    assert np.allclose(mx.nd.linalg.sumlogdiag(L).asnumpy() * 2, np.linalg.slogdet(spd)[1], atol=1e-4)
    sign, logdet = mx.nd.linalg.slogdet(A)
    assert np.allclose(sign.asnumpy(), 1.0) and np.allclose(logdet.asnumpy(), np.linalg.slogdet(spd)[1], atol=1e-4)
    tri = mx.nd.linalg.extracttrian(A)
    assert tri.shape == (2, 10) and np.allclose(mx.nd.linalg.maketrian(tri).asnumpy(), np.tril(spd))
    up = mx.nd.linalg.extracttrian(A, offset=1)
    assert up.shape == (2, 6) and np.allclose(mx.nd.linalg.maketrian(up, offset=1).asnumpy(), np.triu(spd, 1))
    assert np.allclose(mx.nd.linalg.inverse(A).asnumpy() @ spd, np.broadcast_to(np.eye(4), (2, 4, 4)), atol=1e-3)


def test_nd_image_namespace():
    rs = np.random.RandomState(1)
    img = mx.nd.array(rs.randint(0, 256, (8, 6, 3)).astype(np.uint8))
    t = mx.nd.image.to_tensor(img)
    assert t.shape == (3, 8, 6) and t.dtype == np.float32 and float(t.max().asscalar()) <= 1.0
    n = mx.nd.image.normalize(t, mean=(0.5, 0.4, 0.3), std=(0.2, 0.2, 0.2))
    assert np.allclose(n.asnumpy()[1], (t.asnumpy()[1] - 0.4) / 0.2, atol=1e-6)
    assert np.array_equal(mx.nd.image.flip_left_right(img).asnumpy(), img.asnumpy()[:, ::-1]) and np.array_equal(mx.nd.image.flip_top_bottom(img).asnumpy(), img.asnumpy()[::-1])
    assert mx.nd.image.resize(img, (12, 16)).shape == (16, 12, 3) and mx.nd.image.resize(img, 12, keep_ratio=True).shape == (16, 12, 3)
    assert mx.nd.image.crop(img, 1, 2, 3, 4).shape == (4, 3, 3)
    for fn in (lambda x: mx.nd.image.random_brightness(x, 0.5, 1.5), lambda x: mx.nd.image.random_contrast(x, 0.5, 1.5),
               lambda x: mx.nd.image.random_saturation(x, 0.5, 1.5), lambda x: mx.nd.image.random_hue(x, -0.1, 0.1),
               lambda x: mx.nd.image.random_color_jitter(x, 0.2, 0.2, 0.2, 0.05), lambda x: mx.nd.image.random_lighting(x, 0.1)):
        o = fn(img)
        assert o.shape == img.shape and o.dtype == np.uint8
    assert np.array_equal(mx.nd.image.random_saturation(img, 1.0, 1.0).asnumpy(), img.asnumpy())       # factor 1 is the identity
    assert np.array_equal(mx.nd.image.adjust_lighting(img, (0.0, 0.0, 0.0)).asnumpy(), img.asnumpy())


def test_nd_op_and_internal_namespaces():
    a = mx.nd.array([[1.0, -2.0], [3.0, 4.0]])
    assert np.array_equal(mx.nd.op.relu(a).asnumpy(), [[1, 0], [3, 4]]) and mx.nd.op.zeros((2, 3)).shape == (2, 3)
    I = mx.nd._internal
    assert np.array_equal(I._plus_scalar(a, 1).asnumpy(), a.asnumpy() + 1) and np.array_equal(I._rminus_scalar(a, 1).asnumpy(), 1 - a.asnumpy())
    assert np.array_equal(I._rdiv_scalar(a, 2).asnumpy(), 2 / a.asnumpy()) and np.array_equal(I._greater_scalar(a, 0).asnumpy(), (a.asnumpy() > 0))
    assert np.array_equal(I._mul(a, a).asnumpy(), a.asnumpy() ** 2) and np.array_equal(I._lesser_equal(a, a).asnumpy(), np.ones((2, 2)))
    out = mx.nd.zeros((2, 2)); I._set_value(7.0, out=out)
    assert np.array_equal(out.asnumpy(), np.full((2, 2), 7.0)) and I._zeros((3,)).shape == (3,) and I._full((2,), 5).asnumpy().tolist() == [5, 5]
    assert I._random_uniform(0, 1, shape=(4,)).shape == (4,) and np.array_equal(I._copyto(a, out=out).asnumpy(), a.asnumpy())
    with pytest.raises(AttributeError):
        I._no_such_operator


def test_sym_namespaces():
    a = mx.sym.Variable("a")
    g = mx.sym.linalg.gemm2(a, a, transpose_b=True)
    x = np.arange(6, dtype=np.float32).reshape(2, 3)
    assert np.allclose(g.bind(mx.cpu(), {"a": mx.nd.array(x)}).forward()[0].asnumpy(), x @ x.T)
    assert mx.sym.random.uniform(low=0, high=1, shape=(2, 3)).bind(mx.cpu(), {}).forward()[0].shape == (2, 3)
    assert np.array_equal(mx.sym.op.relu(a).bind(mx.cpu(), {"a": mx.nd.array(x - 2)}).forward()[0].asnumpy(), np.maximum(x - 2, 0))
    img = mx.sym.image.to_tensor(a)
    assert img.bind(mx.cpu(), {"a": mx.nd.array(np.zeros((4, 5, 3), dtype=np.uint8))}).forward()[0].shape == (3, 4, 5)
    assert np.array_equal(mx.sym._internal._plus_scalar(a, scalar=2.0).bind(mx.cpu(), {"a": mx.nd.array(x)}).forward()[0].asnumpy(), x + 2)
    with pytest.raises(AttributeError):
        mx.sym.linalg.no_such_op


def test_group_adagrad_and_model_store(tmp_path):
    o = mx.optimizer.create("groupadagrad", learning_rate=0.1)
    assert isinstance(o, mx.optimizer.contrib.GroupAdaGrad)
    w, g = mx.nd.ones((4, 3)), mx.nd.array(np.arange(12, dtype=np.float32).reshape(4, 3))
    st = o.create_state(0, w)
    o.update(0, w, g, st)
    h = (g.asnumpy() ** 2).mean(1, keepdims=True)
    assert np.allclose(st.asnumpy(), h) and np.allclose(w.asnumpy(), 1 - 0.1 * g.asnumpy() / np.sqrt(h + 1e-5), atol=1e-6)
    with pytest.raises(Exception):
        mx.optimizer.contrib.GroupAdaGrad(wd=0.1).update(0, w, g, st)
    # model store: resolves files that are already there (hash-tagged names are verified), explains itself otherwise
    store = mx.gluon.model_zoo.model_store
    net = mx.gluon.model_zoo.vision.get_model("squeezenet1.1", classes=10)
    net.initialize(); net(mx.nd.zeros((1, 3, 64, 64)))
    f = str(tmp_path / "squeezenet1.1.params")
    net.save_parameters(f)
    assert store.get_model_file("squeezenet1.1", root=str(tmp_path)) == f
    tag = hashlib.sha1(open(f, "rb").read()).hexdigest()[:8]
    tagged = str(tmp_path / ("squeezenet1.1-%s.params" % tag))
    os.rename(f, tagged)
    assert store.get_model_file("squeezenet1.1", root=str(tmp_path)) == tagged and store.short_hash("squeezenet1.1", str(tmp_path)) == tag
    net2 = mx.gluon.model_zoo.vision.get_model("squeezenet1.1", pretrained=True, root=str(tmp_path), classes=10)
    x = mx.nd.array(np.random.RandomState(0).rand(1, 3, 64, 64).astype(np.float32))
    assert np.allclose(net2(x).asnumpy(), net(x).asnumpy(), atol=1e-5)
    with open(tagged, "ab") as fh:
        fh.write(b"x")
    with pytest.raises(mx.MXNetError, match="does not match the hash"):
        store.get_model_file("squeezenet1.1", root=str(tmp_path))
    with pytest.raises(mx.MXNetError, match="no parameter file"):
        store.get_model_file("resnet18_v1", root=str(tmp_path))
    store.purge(str(tmp_path))
    assert not os.listdir(str(tmp_path))


def test_class_level_members_of_the_reference_front_end():
    """Members the reference's classes have and user code touches: sparse array utilities, DataDesc helpers, ImageIter pipeline hooks,
    BaseModule interface, Context.default_ctx, Initializer.set_verbosity, RecurrentCell.reset, Symbol.split."""
    csr = mx.nd.array(np.array([[0, 1, 0], [2, 0, 3], [0, 0, 0]], dtype=np.float32)).tostype("csr")
    csr.check_format()
    assert (csr.size, csr.ndim) == (9, 2) and csr.astype("float64").dtype == np.float64
    assert np.array_equal(csr.asscipy().toarray(), csr.asnumpy()) and np.array_equal(csr.copyto(mx.nd.zeros((3, 3))).asnumpy(), csr.asnumpy())
    assert isinstance(csr, mx.nd.sparse.BaseSparseNDArray) and isinstance(csr.copyto(mx.cpu()), mx.nd.sparse.CSRNDArray)
    rs = mx.nd.sparse.row_sparse_array((np.ones((2, 2), dtype=np.float32), [3, 1]), shape=(4, 2))
    rs.check_format()
    with pytest.raises(mx.MXNetError):
        mx.nd.sparse.RowSparseNDArray(mx.nd.ones((2, 2)), mx.nd.array([1, 1], dtype="int64"), (4, 2)).check_format()
    with pytest.raises(NotImplementedError):
        rs.reshape((2, 4))
    d = mx.io.DataDesc("data", [2, 3])
    assert d.shape == (2, 3) and d.dtype == "float32" and mx.io.DataDesc.get_batch_axis("TNC") == 1 and mx.io.DataDesc.get_batch_axis(None) == 0
    assert mx.io.DataDesc.get_list([("a", (1, 2))], [("a", "float16")])[0].dtype == "float16"
    for m in ("hard_reset", "next_sample", "read_image", "imdecode", "check_valid_image", "check_data_shape", "augmentation_transform", "postprocess_data"):
        assert callable(getattr(mx.image.ImageIter, m))
    with pytest.raises(NotImplementedError):
        mx.mod.BaseModule().forward(None)
    assert issubclass(mx.io.MNISTIter, mx.io.MXDataIter) and not issubclass(mx.io.NDArrayIter, mx.io.MXDataIter)
    with mx.cpu(0):
        assert str(mx.cpu(3).default_ctx) == "cpu(0)"
    a = mx.sym.Variable("a")
    halves = a.split(num_outputs=2, axis=1)
    x = np.arange(8, dtype=np.float32).reshape(2, 4)
    assert np.array_equal(halves[1].bind(mx.cpu(), {"a": mx.nd.array(x)}).forward()[0].asnumpy(), x[:, 2:])
    assert np.array_equal((1.0 - a).bind(mx.cpu(), {"a": mx.nd.array(x)}).forward()[0].asnumpy(), 1 - x)
    assert np.allclose((2.0 / (a + 1)).bind(mx.cpu(), {"a": mx.nd.array(x)}).forward()[0].asnumpy(), 2 / (x + 1))
    w = mx.nd.zeros((3, 3))
    mx.init.Xavier().set_verbosity(True)(mx.init.InitDesc("w_weight"), w)
    assert float(w.abs().sum().asscalar()) > 0
    c = mx.gluon.rnn.LSTMCell(4, input_size=3)
    c.initialize(); c(mx.nd.ones((2, 3)), c.begin_state(2)); c.reset()
    assert c._counter == -1


def test_top_level_aliases_and_torch_bridge():
    """mx.th / mx.torch (torch functions on NDArrays, zero copy), mx.rnd, mx.mon, mx.ndarray_doc / mx.symbol_doc."""
    a = mx.nd.array([[1.0, 2.0], [3.0, 4.0]])
    assert np.array_equal(mx.th.mm(a, a).asnumpy(), a.asnumpy() @ a.asnumpy()) and isinstance(mx.th.svd(a)[0], mx.nd.NDArray)
    t = mx.th.to_torch(a)
    t[0, 0] = 9.0
    assert float(a[0, 0].asscalar()) == 9.0 and mx.th.from_torch(t)._t is t and mx.torch is mx.th
    with pytest.raises(AttributeError):
        mx.th.no_such_function
    assert mx.rnd is mx.random and mx.mon is mx.monitor
    doc = mx.ndarray_doc._build_doc("relu", "Rectifier.", ["data", "num_args"], ["NDArray", "int"], ["the   input\n array", "count"], key_var_num_args="num_args",
                                    ret_type="NDArray")
    assert "Parameters\n----------\ndata : NDArray\n    the input array" in doc and "num_args" not in doc and "Returns" in doc
    fc = mx.sym.FullyConnected(mx.sym.Variable("data"), num_hidden=3, name="fc")
    assert mx.symbol_doc.SymbolDoc.get_output_shape(fc, data=(2, 5)) == {"fc_output": (2, 3)}


def test_data_parallel_executor_group():
    from geomx_b200.module.executor_group import DataParallelExecutorGroup
    net = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.Variable("data"), num_hidden=3, name="fc"), name="softmax")
    g = DataParallelExecutorGroup(net, [mx.cpu(0), mx.cpu(1)], [1, 1], [("data", (8, 4))], [("softmax_label", (8,))], ["fc_weight", "fc_bias"], True, False)
    w = np.random.RandomState(0).rand(3, 4).astype(np.float32)
    g.set_params({"fc_weight": mx.nd.array(w), "fc_bias": mx.nd.zeros((3,))}, {})
    x = np.random.RandomState(1).rand(8, 4).astype(np.float32)
    g.forward(mx.io.DataBatch([mx.nd.array(x)], [mx.nd.array(np.arange(8) % 3)]), is_train=True)
    g.backward()
    out = g.get_outputs()[0].asnumpy()
    e = np.exp(x @ w.T); ref = e / e.sum(1, keepdims=True)
    assert out.shape == (8, 3) and np.allclose(out, ref, atol=1e-5)
    assert len(g.execs) == 2 and [(s.start, s.stop) for s in g.slices] == [(0, 4), (4, 8)] and g.grad_arrays[0][0].shape == (3, 4)
    arg, aux = {}, {}
    g.get_params(arg, aux)
    assert np.allclose(arg["fc_weight"].asnumpy(), w)
    m = mx.metric.Accuracy(); g.update_metric(m, [mx.nd.array(np.arange(8) % 3)])
    assert m.num_inst == 8
