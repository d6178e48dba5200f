"""Second sweep over the reference's Python surface: fluent NDArray / Symbol methods, autograd.Function, SymbolBlock + export/imports, legacy
operator classes and FeedForward, extra metrics / initializers / augmenters / datasets, Module conveniences, sparse helpers."""
import os

import numpy as np
import pytest

import geomx_b200 as mx


def test_fluent_ndarray_dlpack_and_legacy_helpers():
    x = mx.nd.array([[1.0, 4.0], [9.0, 16.0]])
    assert x.sqrt().asnumpy().tolist() == [[1, 2], [3, 4]] and x.slice_axis(axis=1, begin=0, end=1).shape == (2, 1)
    assert x.flip(axis=0).asnumpy().tolist() == [[9, 16], [1, 4]] and x.topk(k=1).asnumpy().tolist() == [[1], [1]]
    view = mx.nd.from_dlpack(x.to_dlpack_for_write())
    view[:] = 0
    assert float(x.asnumpy().sum()) == 0                                # zero copy
    assert mx.nd.concatenate([x, x]).shape == (4, 2)
    assert mx.nd.onehot_encode(mx.nd.array([1, 0]), mx.nd.zeros((2, 3))).asnumpy().tolist() == [[0, 1, 0], [1, 0, 0]]
    assert (7 % mx.nd.array([2.0, 4.0])).asnumpy().tolist() == [1, 3]
    a = mx.nd.array([5.0, 7.0]); a %= 4
    assert a.asnumpy().tolist() == [1, 3]
    src = np.arange(4, dtype=np.float32)
    b = mx.nd.array(src); b += 1
    assert src.tolist() == [0, 1, 2, 3]                                  # nd.array never aliases its source
    buf = mx.nd.save_bytes({"w": mx.nd.ones((2,))})
    assert mx.nd.load_frombuffer(buf)["w"].asnumpy().tolist() == [1, 1]
    import scipy.sparse as sp
    m = mx.nd.sparse.array(sp.random(4, 5, 0.5, format="coo", dtype=np.float32, random_state=0))
    assert m.stype == "csr" and isinstance(m, mx.nd.sparse.BaseSparseNDArray) and not isinstance(x, mx.nd.sparse.BaseSparseNDArray)
    assert mx.nd.sparse.multiply(m, m).stype == "csr" and mx.nd.sparse.subtract(m, x.reshape((4,))[:1] if False else m).stype == "csr"
    assert mx.nd.sparse.empty("row_sparse", (3, 2)).shape == (3, 2)


def test_symbol_fluent_eval_executor_reshape_and_monitor():
    x = mx.sym.Variable("x")
    y = (x.exp().sum(axis=1) ** 2).sqrt()
    np.testing.assert_allclose(y.eval(x=mx.nd.array([[0.0, 0.0], [1.0, 1.0]]))[0].asnumpy(), [2, 2 * np.e], rtol=1e-5)
    np.testing.assert_allclose(mx.sym.pow(2.0, x).eval(x=mx.nd.array([1.0, 3.0]))[0].asnumpy(), [2, 8])
    assert x.reshape((2, -1)).infer_shape(x=(4, 3))[1] == [(2, 6)] and x.get_children() is None and len(y.get_children().inputs) == 1
    assert x.infer_shape_partial()[1] == [None] or x.infer_shape_partial()[0] == [None]
    fc = mx.sym.FullyConnected(x, num_hidden=2, name="fc")
    ex = fc.simple_bind(mx.cpu(), x=(3, 4))
    mon = mx.monitor.Monitor(1, pattern=".*output.*"); mon.install(ex)
    mon.tic(); ex.forward(); stats = mon.toc()
    assert any(k == "fc_output" for _, k, _ in stats) and list(ex.output_dict) == ["fc_output"]
    ex2 = ex.reshape(x=(5, 4))
    assert ex2.arg_dict["fc_weight"] is ex.arg_dict["fc_weight"] and ex2.forward()[0].shape == (5, 2)
    assert "Op:FullyConnected" in fc.debug_str()
    with pytest.raises(NotImplementedError):
        fc.asnumpy()


def test_autograd_function_symbolblock_export_and_legacy_ops(tmp_path):
    class Sigmoid(mx.autograd.Function):
        def forward(self, x):
            y = 1 / (1 + mx.nd.exp(-x)); self.save_for_backward(y); return y

        def backward(self, dy):
            y, = self.saved_tensors; return dy * y * (1 - y)
    a = mx.nd.array([0.0, 1.0]); a.attach_grad()
    f = Sigmoid()
    with mx.autograd.record():
        c = (f(a) * 2).sum()
    c.backward()
    s = 1 / (1 + np.exp(-np.array([0.0, 1.0])))
    np.testing.assert_allclose(a.grad.asnumpy(), 2 * s * (1 - s), rtol=1e-5)
    with pytest.raises(AssertionError):
        f(a)                                                             # one use per instance

    d = mx.sym.Variable("data")
    net = mx.sym.FullyConnected(mx.sym.Activation(mx.sym.BatchNorm(mx.sym.FullyConnected(d, num_hidden=8, name="fc1"), name="bn"), act_type="relu"), num_hidden=3, name="fc2")
    mx.random.seed(3)
    blk = mx.gluon.SymbolBlock(net, d); blk.initialize(mx.init.Xavier())
    x = mx.nd.array(np.random.RandomState(0).randn(5, 4).astype(np.float32))
    tr = mx.gluon.Trainer(blk.collect_params(), "sgd", {"learning_rate": 0.005})
    with mx.autograd.record():
        l0 = (blk(x) ** 2).sum()
    l0.backward(); tr.step(1)
    with mx.autograd.record():
        l1 = (blk(x) ** 2).sum()
    assert float(l1.asscalar()) < float(l0.asscalar())                  # the imported graph is trainable
    assert blk.collect_params()["bn_moving_mean"].grad_req == "null"
    blk.export(str(tmp_path / "m"), 3)
    assert sorted(os.listdir(str(tmp_path))) == ["m-0003.params", "m-symbol.json"]
    b2 = mx.gluon.SymbolBlock.imports(str(tmp_path / "m-symbol.json"), ["data"], str(tmp_path / "m-0003.params"))
    np.testing.assert_allclose(b2(x).asnumpy(), blk(x).asnumpy(), atol=1e-6)
    arg, aux = mx.model.load_params(str(tmp_path / "m"), 3)              # same file layout as save_checkpoint
    assert "fc1_weight" in arg and "bn_moving_var" in aux
    assert "fc1" in mx.viz.plot_network(net, shape={"data": (5, 4)}).source

    class Sq(mx.operator.NumpyOp):
        def forward(self, in_data, out_data): out_data[0][:] = in_data[0] ** 2
        def backward(self, out_grad, in_data, out_data, in_grad): in_grad[0][:] = 2 * in_data[0] * out_grad[0]
    op = Sq(); v = mx.nd.array([1.0, 2.0, 3.0]); v.attach_grad()
    with mx.autograd.record():
        b = op(v).sum()
    b.backward()
    assert v.grad.asnumpy().tolist() == [2, 4, 6]
    assert op.get_symbol(mx.sym.Variable("v")).eval(v=mx.nd.array([3.0]))[0].asnumpy().tolist() == [9]

    X = np.random.RandomState(0).randn(120, 4).astype(np.float32); Y = (X[:, 0] > 0).astype(np.float32)
    ff = mx.model.FeedForward(mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.Variable("data"), num_hidden=2, name="o"), name="softmax"),
                              num_epoch=20, optimizer="adam", learning_rate=0.05, initializer=mx.init.Xavier(), numpy_batch_size=40)
    ff.fit(X, Y)
    assert ff.score(mx.io.NDArrayIter(X, Y, batch_size=40)) > 0.9 and ff.predict(X).shape == (120, 2)
    ff.save(str(tmp_path / "ff"), 20)
    ff2 = mx.model.FeedForward.load(str(tmp_path / "ff"), 20)
    np.testing.assert_allclose(ff2.predict(X), ff.predict(X), atol=1e-6)


def test_metrics_initializers_augmenters_datasets(tmp_path):
    m = mx.metric.create("mcc")
    m.update([mx.nd.array([1, 0, 1, 1])], [mx.nd.array([[0.1, 0.9], [0.8, 0.2], [0.7, 0.3], [0.4, 0.6]])])
    assert m.get()[1] == pytest.approx(0.57735, abs=1e-4)
    n = mx.metric.NegativeLogLikelihood(); n.update([mx.nd.array([1, 0])], [mx.nd.array([[0.1, 0.9], [0.8, 0.2]])])
    assert n.get()[1] == pytest.approx(-(np.log(0.9) + np.log(0.8)) / 2, abs=1e-5)
    cfg = mx.metric.Accuracy().get_config()
    assert cfg["metric"] == "Accuracy" and isinstance(mx.metric.create("caffe"), mx.metric.Loss)
    acc = mx.metric.Accuracy(output_names=["o"], label_names=["l"])
    acc.update_dict({"l": mx.nd.array([1]), "zz": mx.nd.array([0])}, {"o": mx.nd.array([[0.1, 0.9]]), "other": mx.nd.array([[1.0, 0.0]])})
    assert acc.get()[1] == 1.0
    with pytest.raises(ValueError):
        mx.metric.check_label_shapes([1, 2], [1])
    b = mx.nd.zeros((8,)); mx.init.LSTMBias(2.0)(mx.init.InitDesc("l_bias"), b)
    assert b.asnumpy().tolist() == [0, 0, 2, 2, 0, 0, 0, 0]
    h, L, inp = 3, 2, 5
    tot = 4 * h * (inp + h) + 4 * h * (h + h) + L * 2 * 4 * h
    f = mx.nd.zeros((tot,)); mx.init.FusedRNN(mx.init.Uniform(0.1), h, L, "lstm")(mx.init.InitDesc("rnn_parameters"), f)
    fv = f.asnumpy(); nb = L * 2 * 4 * h
    assert (fv[:-nb] != 0).all() and fv[-nb:].reshape(-1, 4, h)[:, 1].tolist() == [[1.0] * h] * (L * 2) and fv[-nb:].sum() == L * 2 * h

    img = mx.nd.array(np.random.RandomState(0).randint(0, 255, (20, 30, 3)).astype(np.uint8))
    g = mx.image.RandomGrayAug(1.0)(img).asnumpy()
    assert np.allclose(g[..., 0], g[..., 1])
    assert np.abs(mx.image.HueJitterAug(0.0)(img).asnumpy() - img.asnumpy()).max() < 1.0
    assert mx.image.RandomSizedCropAug((8, 8), 0.3, (0.75, 1.33))(img).shape == (8, 8, 3) and mx.image.scale_down((640, 480), (720, 120)) == (640, 106)
    augs = mx.image.CreateAugmenter((3, 8, 8), rand_crop=True, rand_resize=True, brightness=0.1, hue=0.1, pca_noise=0.1, rand_gray=0.1, mean=True, std=True)
    out = img
    for a in augs:
        out = a(out)
    assert out.shape == (8, 8, 3) and len(augs) == 7
    T = mx.gluon.data.vision.transforms
    for t in (T.RandomSaturation(0.3), T.RandomHue(0.3), T.RandomColorJitter(0.1, 0.1, 0.1, 0.1), T.RandomLighting(0.1)):
        assert t(img).shape == (20, 30, 3)

    from PIL import Image
    for cls in ("cat", "dog"):
        os.makedirs(str(tmp_path / "imgs" / cls))
        for i in range(2):
            Image.fromarray(np.full((6, 5, 3), 40 if cls == "cat" else 200, dtype=np.uint8)).save(str(tmp_path / "imgs" / cls / ("%d.png" % i)))
    ds = mx.gluon.data.vision.ImageFolderDataset(str(tmp_path / "imgs"))
    assert ds.synsets == ["cat", "dog"] and len(ds) == 4 and ds[3][1] == 1 and ds[0][0].shape == (6, 5, 3)
    rec = mx.recordio.MXIndexedRecordIO(str(tmp_path / "d.idx"), str(tmp_path / "d.rec"), "w")
    for i in range(3):
        rec.write_idx(i, mx.recordio.pack_img(mx.recordio.IRHeader(0, float(i), i, 0), np.full((4, 4, 3), 10 * i, dtype=np.uint8), img_fmt=".png"))
    rec.close()
    rd = mx.gluon.data.vision.ImageRecordDataset(str(tmp_path / "d.rec"))
    assert len(rd) == 3 and rd[2][1] == 2.0 and int(rd[2][0].asnumpy()[0, 0, 0]) == 20
    assert len(mx.gluon.data.RecordFileDataset(str(tmp_path / "d.rec"))) == 3
    c100 = mx.gluon.data.vision.CIFAR100(root=str(tmp_path / "none"), fine_label=True, train=False)
    assert c100.synthetic and c100[0][0].shape == (32, 32, 3)


def test_module_conveniences_and_python_loss_module(tmp_path):
    rs = np.random.RandomState(0)
    x = rs.randn(64, 4).astype(np.float32); y = (x[:, 0] > 0).astype(np.float32)
    it = mx.io.NDArrayIter(x, y, batch_size=16)
    mod = mx.mod.Module(mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.Variable("data"), num_hidden=2, name="fc"), name="softmax"))
    mod.fit(it, num_epoch=10, optimizer="adam", optimizer_params={"learning_rate": 0.05}, initializer=mx.init.Xavier())
    assert mod.data_shapes[0].shape == (16, 4) and mod.label_names == ["softmax_label"] and mod.output_shapes == [("softmax_output", (16, 2))]
    outs = [o for o, _, _ in mod.iter_predict(it)]
    assert len(outs) == 4 and outs[0][0].shape == (16, 2)
    mod.save_params(str(tmp_path / "p.params"))
    w = mod.get_params()[0]["fc_weight"].asnumpy().copy()
    mod.set_params({k: v * 0 for k, v in mod.get_params()[0].items()}, {})
    mod.load_params(str(tmp_path / "p.params"))
    np.testing.assert_allclose(mod.get_params()[0]["fc_weight"].asnumpy(), w)
    mod.forward(mx.io.DataBatch([mx.nd.array(x[:5])], [mx.nd.array(y[:5])]), is_train=False)       # different batch size re-binds on the fly
    assert mod.get_outputs()[0].shape == (5, 2)
    np.testing.assert_allclose(mod.get_params()[0]["fc_weight"].asnumpy(), w)
    mon = mx.monitor.Monitor(1, pattern="fc_output"); mod.install_monitor(mon)
    mon.tic(); mod.forward(mx.io.DataBatch([mx.nd.array(x[:5])], [mx.nd.array(y[:5])]), is_train=False)
    assert [k for _, k, _ in mon.toc()] == ["fc_output"]

    # a Python loss at the end of a SequentialModule: gradient of 0.5 * ||scores - onehot||^2
    def grad(scores, labels):
        return scores - mx.nd.one_hot(labels, 2)
    feat = mx.mod.Module(mx.sym.FullyConnected(mx.sym.Variable("data"), num_hidden=2, name="lin"), label_names=None)
    seq = mx.mod.SequentialModule().add(feat).add(mx.mod.PythonLossModule(grad_func=grad), take_labels=True, auto_wiring=True)
    seq.fit(it, num_epoch=15, optimizer="sgd", optimizer_params={"learning_rate": 0.1}, initializer=mx.init.Xavier())
    assert dict(seq.score(it, "acc"))["accuracy"] > 0.9

    class Proto(mx.io.DataIter):                                       # low-level iterator protocol
        def __init__(self):
            super().__init__(2); self.i = 0

        def iter_next(self):
            self.i += 1
            return self.i <= 3

        def getdata(self): return [mx.nd.ones((2, 2)) * self.i]
        def getlabel(self): return [mx.nd.zeros((2,))]
    assert [float(b.data[0].asnumpy()[0, 0]) for b in Proto()] == [1, 2, 3]
    assert mx.misc.FactorScheduler(10, 0.5)(25) == 0.0025
    p = mx.gluon.Parameter("emb", shape=(6, 2), stype="row_sparse", grad_stype="row_sparse"); p.initialize(mx.init.One())
    rsd = p.row_sparse_data(mx.nd.array([4, 1, 4]))
    assert rsd.stype == "row_sparse" and rsd.indices.asnumpy().tolist() == [1, 4] and rsd.data.shape == (2, 2)


def test_test_utils_helpers():
    tu = mx.test_utils
    x = mx.sym.Variable("x"); y = mx.sym.Variable("y")
    s = mx.sym.broadcast_mul(x, y)
    a = np.random.RandomState(0).randn(3, 4).astype(np.float32); b = np.random.RandomState(1).randn(3, 4).astype(np.float32)
    tu.check_symbolic_forward(s, [a, b], [a * b])
    tu.check_symbolic_backward(s, {"x": a, "y": b}, [np.ones((3, 4), np.float32)], {"x": b, "y": a})
    with pytest.raises(AssertionError):
        tu.check_symbolic_forward(s, [a, b], [a * b + 1])
    assert tu.simple_forward(s, x=a, y=b).shape == (3, 4) and tu.check_speed(s, location={"x": a, "y": b}, N=3) > 0
    arr, (vals, idx) = tu.rand_sparse_ndarray((6, 3), "row_sparse", density=0.5)
    assert arr.stype == "row_sparse" and vals.shape == (len(idx), 3)
    csr, (indptr, indices, data) = tu.rand_sparse_ndarray((5, 7), "csr", density=0.4, shuffle_csr_indices=True)
    import scipy.sparse as sp
    np.testing.assert_allclose(csr.tostype("default").asnumpy(), sp.csr_matrix((data, indices, indptr), shape=(5, 7)).toarray(), rtol=1e-6)
    assert tu.create_sparse_array_zd((4, 2), "row_sparse", 0.0).tostype("default").asnumpy().sum() == 0
    v = mx.nd.ones((3,)); assert tu.same_array(v, v) and not tu.same_array(v, v.copy())
    assert tu.almost_equal_ignore_nan(np.array([1.0, np.nan]), np.array([1.0, 2.0]))
    tu.assert_exception(lambda: 1 / 0, ZeroDivisionError)
    assert tu.np_reduce(np.ones((2, 3, 4)), (0, 2), True, np.sum).shape == (1, 3, 1) and tu.np_reduce(np.ones((2, 3)), None, False, np.sum) == 6
    tu.compare_optimizer(mx.optimizer.SGD(learning_rate=0.1, momentum=0.9), mx.optimizer.SGD(learning_rate=0.1, momentum=0.9), (4, 3), np.float32)
    import scipy.stats as ss
    buckets, probs = tu.gen_buckets_probs_with_ppf(lambda q: ss.norm.ppf(q, 0, 1), 5)
    tu.verify_generator(lambda n: mx.nd.random.normal(0, 1, shape=(n,)).asnumpy(), buckets, probs, nsamples=20000, nrepeat=3)
    with pytest.raises(AssertionError):
        tu.verify_generator(lambda n: mx.nd.random.uniform(-1, 1, shape=(n,)).asnumpy(), buckets, probs, nsamples=20000, nrepeat=3)
    assert tu.mean_check(lambda n: np.random.normal(2, 1, n), 2, 1, 20000) and tu.var_check(lambda n: np.random.normal(2, 1, n), 1, 20000)
    with tu.EnvManager("GEOMX_TMP_VAR", "7"):
        assert os.environ["GEOMX_TMP_VAR"] == "7"
    assert "GEOMX_TMP_VAR" not in os.environ
    with pytest.raises(IOError):
        tu.download("http://example.invalid/x.bin", dirname="/nonexistent")
    os.environ["GEOMX_SYNTHETIC_SIZE"] = "64"
    try:
        d = tu.get_mnist("/nonexistent")
        tr, va = tu.get_mnist_iterator(8, (1, 28, 28), num_parts=2, part_index=1, path="/nonexistent")
    finally:
        os.environ.pop("GEOMX_SYNTHETIC_SIZE")
    assert d["train_data"].shape == (64, 1, 28, 28) and d["train_data"].max() <= 1.0 and tr.provide_data[0].shape == (8, 1, 28, 28)
    it = tu.DummyIter(va); b1 = next(it); b2 = next(it)
    assert b1 is b2 and tu.get_im2rec_path().endswith("im2rec.py")
    calls = []

    @tu.retry(3)
    def flaky():
        calls.append(1)
        assert len(calls) >= 2
    flaky(); assert len(calls) == 2


def test_detection_augmenters_and_iterator(tmp_path):
    from PIL import Image
    rs = np.random.RandomState(0)
    entries = []
    for i in range(5):
        Image.fromarray(rs.randint(0, 255, (40, 60, 3)).astype(np.uint8)).save(str(tmp_path / ("%d.png" % i)))
        objs = [[i % 3, 0.1, 0.2, 0.5, 0.7], [1, 0.4, 0.4, 0.9, 0.95]][:1 + i % 2]
        entries.append([2, 5] + [v for o in objs for v in o] + ["%d.png" % i])
    it = mx.image.ImageDetIter(2, (3, 32, 32), imglist=entries, path_root=str(tmp_path), rand_crop=0.5, rand_pad=0.5, rand_mirror=True,
                               mean=True, std=True, brightness=0.1)
    assert it.provide_label[0].shape == (2, 2, 5)
    nb = 0
    for b in it:
        nb += 1
        lab = b.label[0].asnumpy()
        assert b.data[0].shape == (2, 3, 32, 32)
        v = lab[lab[:, :, 0] >= 0]
        assert len(v) >= 2 and (v[:, 1:5] >= -1e-6).all() and (v[:, 1:5] <= 1 + 1e-6).all() and (v[:, 3] > v[:, 1]).all() and (v[:, 4] > v[:, 2]).all()
    assert nb == 3
    lab = np.array([[0, 0.1, 0.2, 0.5, 0.7]], dtype=np.float32)
    img = mx.nd.array(rs.randint(0, 255, (40, 60, 3)).astype(np.uint8))
    fi, fl = mx.image.DetHorizontalFlipAug(1.0)(img, lab)
    np.testing.assert_allclose(fl, [[0, 0.5, 0.2, 0.9, 0.7]], atol=1e-6)
    assert np.array_equal(fi.asnumpy(), img.asnumpy()[:, ::-1])
    pi, pl = mx.image.DetRandomPadAug(area_range=(1.5, 2.0))(img, lab)
    H, W = pi.shape[:2]
    assert H >= 40 and W >= 60 and H * W >= 1.4 * 40 * 60
    # the padded image still shows the object at the moved box: compare the box size in pixels
    assert abs((pl[0, 3] - pl[0, 1]) * W - 0.4 * 60) < 1e-3 and abs((pl[0, 4] - pl[0, 2]) * H - 0.5 * 40) < 1e-3
    ci, cl = mx.image.DetRandomCropAug(min_object_covered=0.5, area_range=(0.3, 0.9))(img, lab)
    assert ci.shape[0] <= 40 and ci.shape[1] <= 60 and (cl[:, 1:5] >= 0).all() and (cl[:, 1:5] <= 1).all()
    sel = mx.image.CreateMultiRandCropAugmenter(min_object_covered=[0.1, 0.5], area_range=[(0.1, 1.0), (0.3, 0.9)])
    assert len(sel.aug_list) == 2 and isinstance(json_ok(sel.dumps()), list)
    it.reset()
    assert next(it.draw_next(mean=np.array([123.68, 116.28, 103.53]), std=np.array([58.395, 57.12, 57.375]))).shape == (32, 32, 3)
    val = mx.image.ImageDetIter(2, (3, 32, 32), imglist=entries[:1], path_root=str(tmp_path))
    assert val.label_shape == (1, 5) and it.sync_label_shape(val).label_shape == (2, 5)
    with pytest.raises(RuntimeError):
        mx.image.ImageDetIter._parse_label([2, 5, 0, 0.5, 0.5, 0.4, 0.4])          # xmax < xmin: no valid object


def json_ok(x):
    import json
    json.dumps(x)
    return x


def test_demo_script_saves_and_loads_params(tmp_path):
    """``--save-params`` / ``--load-params`` of the demo scripts (stand-alone CPU run): the saved file reproduces the trained accuracy."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GEOMX_SYNTHETIC_SIZE="512", GEOMX_SEED="1")
    for k in ("DMLC_ROLE", "DMLC_PS_ROOT_URI", "RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = str(tmp_path / "cnn.params")
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "cnn.py"), "--cpu", "--max-iters", "30", "--eval-every", "30", "--save-params", out],
                       env=env, capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    acc = float(r.stdout.strip().splitlines()[-1].split("Test Acc")[1])
    assert acc > 0.9 and os.path.exists(out)
    sys.path.insert(0, os.path.join(root, "examples"))
    try:
        import common
        os.environ["GEOMX_SYNTHETIC_SIZE"] = "512"
        net = common.build_net(mx.cpu(), 32)
        _, test = common.make_loaders(32, 1, 0, "/nonexistent", False)
        before = common.accuracy(test, net, mx.cpu())
        net.load_parameters(out)
        assert common.accuracy(test, net, mx.cpu()) > 0.9 > before
    finally:
        sys.path.pop(0); os.environ.pop("GEOMX_SYNTHETIC_SIZE", None)


def test_executor_manager_and_notebook_logger():
    from geomx_b200 import executor_manager as em
    rs = np.random.RandomState(0)
    x = rs.randn(32, 4).astype(np.float32); y = (x[:, 0] > 0).astype(np.float32)
    it = mx.io.NDArrayIter(x, y, batch_size=8)
    sym = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.Variable("data"), num_hidden=2, name="fc"), name="softmax")
    m = em.DataParallelExecutorManager(sym, mx.cpu(), it)
    m.set_params({"fc_weight": mx.nd.array(rs.randn(2, 4).astype(np.float32)), "fc_bias": mx.nd.zeros((2,))}, {})
    upd = mx.optimizer.get_updater(mx.optimizer.SGD(learning_rate=0.5, rescale_grad=1 / 8))
    acc = mx.metric.Accuracy()
    for _ in range(10):
        it.reset(); acc.reset()
        for b in it:
            m.load_data_batch(b); m.forward(is_train=True); m.backward()
            for i, (ws, gs) in enumerate(zip(m.param_arrays, m.grad_arrays)):
                upd(i, gs[0], ws[0])
            m.update_metric(acc, b.label)
    assert acc.get()[1] > 0.9
    a, b_ = {}, {}
    m.copy_to(a, b_)
    assert sorted(a) == ["fc_bias", "fc_weight"] and em._split_input_slice(10, [1, 1, 2]) == [slice(0, 2), slice(2, 4), slice(4, 10)]
    with pytest.raises(ValueError):
        em._split_input_slice(2, [1, 1, 1, 1])
    log = mx.notebook.callback.PandasLogger(batch_size=8, frequent=2)
    mod = mx.mod.Module(sym)
    mod.fit(it, num_epoch=2, **log.callback_args())
    assert len(log.train_df) == 4 and {"accuracy", "records_per_sec", "epoch", "minibatch_count"} <= set(log.train_df.columns) and len(log.epoch_df) == 2
    with pytest.raises(ImportError):
        mx.notebook.callback.LiveLearningCurve()
