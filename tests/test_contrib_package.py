"""``mx.contrib``: text, quantization, SVRG, tensorboard logging, legacy autograd, DataLoaderIter, and the generic nd→sym operator bridge."""
import json
import os

import numpy as np
import pytest
import torch

import geomx_b200 as mx
from geomx_b200 import contrib


def test_text_vocabulary_and_embeddings(tmp_path):
    text = contrib.text
    c = text.utils.count_tokens_from_str("a b b c\nc c d", to_lower=True)
    assert c == {"a": 1, "b": 2, "c": 3, "d": 1}
    v = text.Vocabulary(c, most_freq_count=3, min_freq=1, reserved_tokens=["<pad>"])
    assert v.idx_to_token == ["<unk>", "<pad>", "c", "b", "a"] and len(v) == 5
    assert v.to_indices(["c", "zzz", "b"]) == [2, 0, 3] and v.to_tokens([0, 2]) == ["<unk>", "c"]
    with pytest.raises(ValueError):
        v.to_tokens(99)
    p = tmp_path / "emb.txt"
    p.write_text("a 1 2\nb 3 4\nc 5 6\n")
    e = text.embedding.create("customembedding", pretrained_file_path=str(p))
    assert len(e) == 4 and e.vec_len == 2
    np.testing.assert_allclose(e.get_vecs_by_tokens(["b", "x"]).asnumpy(), [[3, 4], [0, 0]])
    np.testing.assert_allclose(e.get_vecs_by_tokens("B", lower_case_backup=True).asnumpy(), [3, 4])
    ce = text.embedding.CompositeEmbedding(v, [e, e])
    assert ce.idx_to_vec.shape == (5, 4)
    np.testing.assert_allclose(ce.get_vecs_by_tokens("c").asnumpy(), [5, 6, 5, 6])
    e.update_token_vectors("a", mx.nd.array([9, 9]))
    np.testing.assert_allclose(e.get_vecs_by_tokens("a").asnumpy(), [9, 9])
    with pytest.raises(ValueError):
        e.update_token_vectors("nope", mx.nd.array([1, 1]))
    assert "glove.6B.50d.txt" in text.embedding.get_pretrained_file_names("glove")
    with pytest.raises(ValueError):                              # never downloads: a missing local file is an error
        text.embedding.create("glove", pretrained_file_name="glove.6B.50d.txt", embedding_root=str(tmp_path))


def _mlp_symbol():
    d = mx.sym.Variable("data")
    h = mx.sym.Activation(mx.sym.FullyConnected(d, num_hidden=16, name="fc1"), act_type="relu")
    return mx.sym.SoftmaxOutput(mx.sym.FullyConnected(h, num_hidden=3, name="fc2"), name="softmax")


def _blobs(n=240, seed=0):
    rs = np.random.RandomState(seed)
    centers = np.array([[2, 0, 0, 0], [0, 2, 0, 0], [0, 0, 2, 0]], dtype=np.float32)
    y = rs.randint(0, 3, n)
    x = centers[y] + 0.3 * rs.randn(n, 4).astype(np.float32)
    return x.astype(np.float32), y.astype(np.float32)


def test_quantize_model_and_net_keep_accuracy():
    x, y = _blobs()
    it = mx.io.NDArrayIter(x, y, batch_size=40, shuffle=False)
    mod = mx.mod.Module(_mlp_symbol())
    mod.fit(it, num_epoch=15, optimizer="adam", optimizer_params={"learning_rate": 0.02}, initializer=mx.init.Xavier())
    base = dict(mod.score(it, "acc"))["accuracy"]
    assert base > 0.95
    arg, aux = mod.get_params()
    for mode, dt in (("naive", "int8"), ("entropy", "int8"), ("none", "fp8")):
        qsym, qarg, qaux = contrib.quantization.quantize_model(mod.symbol, arg, aux, calib_mode=mode, calib_data=it if mode != "none" else None,
                                                               num_calib_examples=120, quantized_dtype=dt)
        w, qw = arg["fc1_weight"].asnumpy(), qarg["fc1_weight"].asnumpy()
        assert not np.array_equal(w, qw) and np.abs(w - qw).max() <= np.abs(w).max() / (127 if dt == "int8" else 14) + 1e-6
        if dt == "int8":
            levels = np.unique(np.round(qw / (np.abs(w).max() / 127.0)))
            assert len(levels) <= 255 and np.allclose(levels, np.round(levels))
        assert float(qarg["fc1_weight_max"].asnumpy()[0]) == pytest.approx(np.abs(w).max())
        node = [s for s in qsym._topo() if s.name == "fc2"][0]
        assert node.attrs["__quantized__"]["dtype"] == dt and (node.attrs["__quantized__"]["act_threshold"] is not None) == (mode != "none")
        qsym = mx.sym.load_json(qsym.tojson())                       # the annotation survives serialisation
        qmod = mx.mod.Module(qsym)
        qmod.bind(it.provide_data, it.provide_label, for_training=False)
        qmod.init_params(arg_params=qarg, aux_params=qaux, allow_extra=True)
        assert dict(qmod.score(it, "acc"))["accuracy"] > base - 0.05
    with pytest.raises(ValueError):
        contrib.quantization.quantize_model(mod.symbol, arg, aux, calib_mode="entropy")
    # gluon
    net = mx.gluon.nn.Sequential()
    net.add(mx.gluon.nn.Dense(16, activation="relu"), mx.gluon.nn.Dense(3))
    net.initialize(mx.init.Xavier())
    tr = mx.gluon.Trainer(net.collect_params(), "adam", {"learning_rate": 0.02})
    L = mx.gluon.loss.SoftmaxCrossEntropyLoss()
    X, Y = mx.nd.array(x), mx.nd.array(y)
    for _ in range(60):
        with mx.autograd.record():
            l = L(net(X), Y).mean()
        l.backward(); tr.step(1)
    acc = float((net(X).argmax(axis=1).asnumpy() == y).mean())
    contrib.quantization.quantize_net(net, calib_data=[(X[:120], Y[:120])], calib_mode="naive")
    qacc = float((net(X).argmax(axis=1).asnumpy() == y).mean())
    assert acc > 0.95 and qacc > acc - 0.05
    assert all(getattr(b, "_quantized", None) for b in net._children.values())


def test_entropy_threshold_clips_outliers():
    rs = np.random.RandomState(0)
    a = rs.randn(20000).astype(np.float32)
    a[:3] = [40.0, -55.0, 60.0]
    thr = contrib.quantization.calib_thresholds({"x": [a]}, "entropy")["x"]
    naive = contrib.quantization.calib_thresholds({"x": [a]}, "naive")["x"]
    assert naive == pytest.approx(60.0) and 2.0 < thr < 20.0


def test_svrg_module_trains_and_combines_gradients():
    x, y = _blobs(seed=1)
    it = mx.io.NDArrayIter(x, y, batch_size=40, shuffle=False)
    with pytest.raises(ValueError):
        contrib.svrg_optimization.SVRGModule(_mlp_symbol(), update_freq=0)
    mod = contrib.svrg_optimization.SVRGModule(_mlp_symbol(), update_freq=2)
    mod.bind(it.provide_data, it.provide_label)
    mod.init_params(mx.init.Xavier())
    mod.init_optimizer(optimizer="sgd", optimizer_params={"learning_rate": 0.1})
    mod.update_full_grads(it)
    # at the snapshot point g(w) == g(w~), so the SVRG gradient of any batch equals the full gradient
    batch = next(iter(it))
    mod.forward_backward(batch)
    mod._svrg_grads_update_rule()
    for n in mod._param_names:
        np.testing.assert_allclose(mod._execs[0].grad_dict[n].asnumpy(), mod._full_grads[n].asnumpy(), rtol=1e-5, atol=1e-6)
    it.reset()
    mod2 = contrib.svrg_optimization.SVRGModule(_mlp_symbol(), update_freq=2)
    mod2.fit(it, num_epoch=12, optimizer="sgd", optimizer_params={"learning_rate": 0.3}, initializer=mx.init.Xavier())
    assert dict(mod2.score(it, "acc"))["accuracy"] > 0.95


def test_tensorboard_callback_autograd_and_dataloader_iter(tmp_path):
    cb = contrib.tensorboard.LogMetricsCallback(str(tmp_path / "tb"), prefix="train")
    m = mx.metric.Accuracy(); m.update([mx.nd.array([1, 0])], [mx.nd.array([[0.1, 0.9], [0.8, 0.2]])])
    cb(mx.model.BatchEndParam(0, 0, m, None)); cb(mx.model.BatchEndParam(0, 1, m, None))
    if isinstance(cb.summary_writer, contrib.tensorboard.JsonlSummaryWriter):
        rows = [json.loads(l) for l in open(os.path.join(str(tmp_path / "tb"), "scalars.jsonl"))]
        assert [r["step"] for r in rows] == [1, 2] and rows[0]["tag"] == "train-accuracy" and rows[0]["value"] == 1.0
    else:
        assert os.listdir(str(tmp_path / "tb"))

    ag = contrib.autograd
    f = ag.grad_and_loss(lambda a, b: (a * a * b).sum(), argnum=0)
    g, loss = f(mx.nd.array([1.0, 2.0]), mx.nd.array([3.0, 4.0]))
    np.testing.assert_allclose(g[0].asnumpy(), [6.0, 16.0]); assert float(loss.asscalar()) == 19.0
    g2 = ag.grad(lambda a: (a * 3).sum())(mx.nd.array([1.0, 1.0]))
    np.testing.assert_allclose(g2[0].asnumpy(), [3.0, 3.0])
    with ag.train_section():
        assert mx.autograd.is_training() and mx.autograd.is_recording()
        with ag.test_section():
            assert not mx.autograd.is_training()
    assert not mx.autograd.is_recording()

    ds = mx.gluon.data.ArrayDataset(np.arange(20, dtype=np.float32).reshape(10, 2), np.arange(10, dtype=np.float32))
    it = contrib.io.DataLoaderIter(mx.gluon.data.DataLoader(ds, batch_size=4))
    assert it.provide_data[0].shape == (4, 2) and it.provide_label[0].name == "softmax_label"
    batches = list(it)
    assert [b.pad for b in batches] == [0, 0, 2] and batches[-1].data[0].shape == (4, 2)
    it.reset()
    assert len(list(it)) == 3
    with pytest.raises(ImportError):
        contrib.onnx.import_model("x.onnx")


def test_generic_symbol_bridge_matches_imperative():
    x = mx.sym.Variable("x"); y = mx.sym.Variable("y")
    z = mx.sym.sum(mx.sym.broadcast_add(mx.sym.exp(x), y), axis=1)
    z = mx.sym.load_json(z.tojson())
    assert z.infer_shape(x=(4, 3), y=(1, 3))[1] == [(4,)]
    ex = z.simple_bind(mx.cpu(), x=(2, 3), y=(1, 3))
    xv = np.random.RandomState(0).randn(2, 3).astype(np.float32)
    ex.arg_dict["x"][:] = mx.nd.array(xv); ex.arg_dict["y"][:] = mx.nd.array([[1, 2, 3]])
    out = ex.forward(is_train=True)[0].asnumpy()
    np.testing.assert_allclose(out, (np.exp(xv) + [[1, 2, 3]]).sum(1), rtol=1e-5)
    ex.backward()
    np.testing.assert_allclose(ex.grad_dict["x"].asnumpy(), np.exp(xv), rtol=1e-5)
    np.testing.assert_allclose(ex.grad_dict["y"].asnumpy(), [[2, 2, 2]])
    r = mx.sym.contrib.ROIAlign(mx.sym.Variable("d"), mx.sym.Variable("r"), pooled_size=(2, 2), spatial_scale=1.0)
    assert r.infer_shape(d=(1, 2, 8, 8), r=(3, 5))[1] == [(3, 2, 2, 2)]
    assert contrib.symbol.box_iou(mx.sym.Variable("a"), mx.sym.Variable("b")).infer_shape(a=(2, 4), b=(3, 4))[1] == [(2, 3)]
    with pytest.raises(AttributeError):
        mx.sym.no_such_operator


def test_symbol_bridge_creation_and_multi_output_ops():
    x = mx.sym.Variable("x")
    assert (mx.sym.zeros(shape=(2, 3)) + x).infer_shape(x=(2, 3))[1] == [(2, 3)]
    out = mx.sym.maximum(x, mx.sym.ones(shape=(2, 3))).eval(x=mx.nd.array([[0.0, 2.0, 0.5]] * 2))[0].asnumpy()
    assert out.tolist() == [[1, 2, 1]] * 2
    rng = mx.sym.load_json((mx.sym.arange(start=0, stop=6).reshape((2, 3)) * x).tojson()).eval(x=mx.nd.ones((2, 3)))[0].asnumpy()
    assert rng.tolist() == [[0, 1, 2], [3, 4, 5]]
    parts = mx.sym.split(x, num_outputs=2, axis=1)
    y = mx.sym.load_json((parts[0] * 2 + parts[1]).tojson())
    ex = y.simple_bind(mx.cpu(), x=(2, 4))
    ex.arg_dict["x"][:] = mx.nd.array(np.arange(8, dtype=np.float32).reshape(2, 4))
    assert ex.forward(is_train=True)[0].asnumpy().tolist() == [[2, 5], [14, 17]]
    ex.backward()
    assert ex.grad_dict["x"].asnumpy().tolist() == [[2, 2, 1, 1]] * 2


def test_symbol_bridge_auto_parameters_train_an_embedding_model():
    d = mx.sym.Variable("data")
    n = mx.sym.LayerNorm(mx.sym.Embedding(d, input_dim=10, output_dim=4, name="emb"), name="ln")
    net = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.mean(n, axis=1), num_hidden=2, name="fc"), name="softmax")
    assert net.list_arguments() == ["data", "emb_weight", "ln_gamma", "ln_beta", "fc_weight", "fc_bias", "softmax_label"]
    assert net.infer_shape(data=(8, 5))[0][1:4] == [(10, 4), (4,), (4,)]
    rs = np.random.RandomState(0)
    y = rs.randint(0, 2, 64)
    x = np.where(rs.rand(64, 5) < 0.8, y[:, None] * 5 + rs.randint(0, 5, (64, 5)), rs.randint(0, 10, (64, 5)))
    it = mx.io.NDArrayIter(x.astype(np.float32), y.astype(np.float32), batch_size=8)
    mod = mx.mod.Module(net)
    mod.fit(it, num_epoch=8, optimizer="adam", optimizer_params={"learning_rate": 0.05}, initializer=mx.init.Xavier())
    assert dict(mod.score(it, "acc"))["accuracy"] > 0.9
    dc = mx.sym.Deconvolution(mx.sym.Variable("x"), kernel=(2, 2), stride=(2, 2), num_filter=3, name="up")
    assert dc.list_arguments() == ["x", "up_weight"] and dc.infer_shape(x=(1, 4, 5, 5))[1] == [(1, 3, 10, 10)]
