"""MXNet-shaped API on CPU: NDArray, autograd, gluon, optimizers, local kvstore (reference: python/mxnet/*)."""
import numpy as np
import pytest
import torch

import geomx_b200 as mx


def _cnn():
    net = mx.gluon.nn.Sequential()
    net.add(mx.gluon.nn.Conv2D(channels=16, kernel_size=5, activation="relu"), mx.gluon.nn.MaxPool2D(pool_size=2, strides=2),
            mx.gluon.nn.Conv2D(channels=32, kernel_size=5, activation="relu"), mx.gluon.nn.MaxPool2D(pool_size=2, strides=2),
            mx.gluon.nn.Dense(256, activation="relu"), mx.gluon.nn.Dense(128, activation="relu"), mx.gluon.nn.Dense(10))
    return net


def test_ndarray_basics():
    a = mx.nd.array([[1, 2], [3, 4]])
    assert a.shape == (2, 2) and a.size == 4 and a.dtype == np.float32
    b = (a * 2 + 1) / 2
    np.testing.assert_allclose(b.asnumpy(), [[1.5, 2.5], [3.5, 4.5]])
    assert a.astype("float16").dtype == np.float16
    assert float(a.mean().asscalar()) == 2.5
    assert a.argmax(axis=1).asnumpy().tolist() == [1.0, 1.0]
    a[:] = 7
    assert a.asnumpy().sum() == 28
    c = mx.nd.zeros((3,), ctx=mx.cpu())
    a2 = a.copyto(mx.cpu())
    assert a2 is not a and (a2 == a).asnumpy().all()
    assert c.context == mx.cpu()


def test_gpu_ctx_raises_without_device():
    if torch.cuda.is_available():
        pytest.skip("has gpu")
    with pytest.raises(mx.base.MXNetError):
        mx.nd.zeros((1,), ctx=mx.gpu(0))


def test_cnn_param_inventory():
    net = _cnn()
    net.initialize(force_reinit=True, ctx=mx.cpu(), init=mx.init.Xavier())
    net(mx.nd.random.uniform(shape=(32, 1, 28, 28)))
    params = list(net.collect_params().values())
    sizes = [int(np.prod(p.shape)) for p in params]
    # 10 tensors, 178 762 elements
    assert sizes == [400, 16, 12800, 32, 131072, 256, 32768, 128, 1280, 10]
    assert sum(sizes) == 178762


def test_autograd_write_and_add():
    x = mx.nd.array([1.0, 2.0, 3.0]); x.attach_grad()
    with mx.autograd.record():
        y = (x * x).sum()
    y.backward()
    np.testing.assert_allclose(x.grad.asnumpy(), [2, 4, 6])
    with mx.autograd.record():
        y = (x * 3).sum()
    y.backward()
    np.testing.assert_allclose(x.grad.asnumpy(), [3, 3, 3])  # 'write' overwrites
    z = mx.nd.array([1.0, 1.0]); z.attach_grad("add")
    for _ in range(2):
        with mx.autograd.record():
            (z * 2).sum().backward()
    np.testing.assert_allclose(z.grad.asnumpy(), [4, 4])


def test_train_converges_with_trainer():
    torch.manual_seed(0)
    net = _cnn(); net.initialize(init=mx.init.Xavier())
    x = mx.nd.random.uniform(shape=(32, 1, 28, 28)); y = mx.nd.array(np.arange(32) % 10)
    loss = mx.gluon.loss.SoftmaxCrossEntropyLoss()
    tr = mx.gluon.Trainer(net.collect_params(), "adam", {"learning_rate": 0.01}, kvstore=None)
    first = last = None
    for i in range(30):
        with mx.autograd.record():
            l = loss(net(x), y)
        l.backward(); tr.step(32)
        v = float(l.mean().asscalar())
        first = v if first is None else first; last = v
    assert last < first * 0.7


def test_local_kvstore_semantics():
    kv = mx.kv.create("local")
    kv.init(3, mx.nd.ones((2, 3)))
    out = mx.nd.zeros((2, 3)); kv.pull(3, out=out)
    assert out.asnumpy().sum() == 6
    kv.push(3, [mx.nd.ones((2, 3)) * 2, mx.nd.ones((2, 3)) * 3])   # multi-value push sums
    kv.pull(3, out=out)
    np.testing.assert_allclose(out.asnumpy(), 5 * np.ones((2, 3)))   # no updater: assign merged
    kv._set_updater(lambda k, g, w: w.__iadd__(g * 2))
    kv.push(3, mx.nd.ones((2, 3))); kv.pull(3, out=out)
    np.testing.assert_allclose(out.asnumpy(), 7 * np.ones((2, 3)))
    # list keys, str keys cannot mix with int keys
    kv2 = mx.kv.create("local"); kv2.init(["a", "b"], [mx.nd.ones((1,)), mx.nd.ones((1,))])
    with pytest.raises(mx.base.MXNetError):
        kv2.init(5, mx.nd.ones((1,)))
    assert kv.type == "local" and kv.rank == 0 and kv.num_workers == 1


def test_kvstore_optimizer_matches_trainer():
    torch.manual_seed(1)
    w0 = torch.randn(10)
    g = torch.randn(10)
    kv = mx.kv.create("local"); kv.set_optimizer(mx.optimizer.Adam(learning_rate=0.1))
    kv.init(0, mx.nd.array(w0)); out = mx.nd.zeros((10,))
    ref = w0.clone().requires_grad_(True); opt = torch.optim.Adam([ref], lr=0.1, eps=1e-8)
    for _ in range(3):
        kv.push(0, mx.nd.array(g)); kv.pull(0, out=out)
        ref.grad = g.clone(); opt.step()
    np.testing.assert_allclose(out.asnumpy(), ref.detach().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["sgd", "adam", "dcasgd", "nag", "rmsprop", "adagrad", "adadelta", "ftrl", "adamax",
                                  "nadam", "signum", "ftml", "sgld"])
def test_optimizers_step(name):
    o = mx.optimizer.create(name, learning_rate=0.01)
    upd = mx.optimizer.get_updater(o)
    w = mx.nd.ones((5,)); g = mx.nd.ones((5,)) * 0.5
    before = w.asnumpy().copy()
    upd(0, g, w); upd(0, g, w)
    assert np.isfinite(w.asnumpy()).all() and not np.allclose(before, w.asnumpy())
    st = upd.get_states(dump_optimizer=True)
    upd2 = mx.optimizer.get_updater(mx.optimizer.create(name)); upd2.set_states(st)
    assert type(upd2.optimizer).__name__.lower() == name


def test_trainer_save_load_states(tmp_path):
    net = mx.gluon.nn.Dense(4, in_units=3); net.initialize()
    tr = mx.gluon.Trainer(net.collect_params(), "adam", {"learning_rate": 0.1}, kvstore=None)
    x = mx.nd.ones((2, 3))
    with mx.autograd.record():
        l = net(x).sum()
    l.backward(); tr.step(2)
    f = str(tmp_path / "t.states"); tr.save_states(f); tr.load_states(f)
    assert tr.optimizer.num_update == 1


def test_dataloader_and_split_sampler():
    ds = mx.gluon.data.vision.MNIST(root="/nonexistent", train=True)
    assert ds.synthetic and ds[0][0].shape == (28, 28, 1)
    tf = mx.gluon.data.vision.transforms.Compose([mx.gluon.data.vision.transforms.Resize((28, 28)),
                                                   mx.gluon.data.vision.transforms.ToTensor()])
    dl = mx.gluon.data.DataLoader(ds.transform_first(tf), 32, num_workers=0)
    X, y = next(iter(dl))
    assert X.shape == (32, 1, 28, 28) and y.shape == (32,) and float(X.max().asscalar()) <= 1.0
    parts = mx.gluon.utils.split_and_load(X, [mx.cpu()])
    assert parts[0].shape == (32, 1, 28, 28)
    # thread workers and process workers (batches come back through shared memory) deliver the same batches in the same order
    small = mx.gluon.data.ArrayDataset(np.random.RandomState(0).rand(50, 3, 4).astype(np.float32), np.arange(50, dtype=np.float32))
    ref = [(a.asnumpy(), b.asnumpy()) for a, b in mx.gluon.data.DataLoader(small, batch_size=8)]
    for kw in ({"num_workers": 2}, {"num_workers": 2, "worker_type": "process"}, {"num_workers": 3, "worker_type": "process", "prefetch": 2}):
        got = [(a.asnumpy(), b.asnumpy()) for a, b in mx.gluon.data.DataLoader(small, batch_size=8, **kw)]
        assert len(got) == len(ref) == 7 and all((p[0] == q[0]).all() and (p[1] == q[1]).all() for p, q in zip(ref, got)), kw


def test_row_sparse_ndarray_and_local_kvstore():
    """row_sparse storage + kv.row_sparse_pull / sparse push on the local store (python/mxnet/ndarray/sparse.py, kvstore_local.h:357-417)."""
    import numpy as np
    dense = np.zeros((6, 3), dtype=np.float32); dense[1] = 1.0; dense[4] = [1, 2, 3]
    rs = mx.nd.array(dense).tostype("row_sparse")
    assert rs.stype == "row_sparse" and rs.indices.asnumpy().tolist() == [1, 4] and rs.data.shape == (2, 3)
    assert np.array_equal(rs.tostype("default").asnumpy(), dense)
    rs2 = mx.nd.sparse.row_sparse_array((np.ones((2, 3), dtype=np.float32), [4, 0]), shape=(6, 3))
    assert rs2.indices.asnumpy().tolist() == [0, 4]
    s = mx.nd.sparse.add(rs, rs2)
    assert s.indices.asnumpy().tolist() == [0, 1, 4] and np.allclose(s.tostype("default").asnumpy()[4], [2, 3, 4])
    assert rs.retain(mx.nd.array([4, 5], dtype="int64")).indices.asnumpy().tolist() == [4]
    z = mx.nd.sparse.zeros("row_sparse", (6, 3))
    assert z.data.shape == (0, 3) and z.tostype("default").asnumpy().sum() == 0

    kv = mx.kv.create("local")
    w = mx.nd.array(np.arange(18, dtype=np.float32).reshape(6, 3))
    kv.init("emb", w)
    out = mx.nd.sparse.zeros("row_sparse", (6, 3))
    kv.row_sparse_pull("emb", out=out, row_ids=mx.nd.array([4, 1, 4], dtype="int64"))
    assert out.indices.asnumpy().tolist() == [1, 4] and np.array_equal(out.data.asnumpy(), w.asnumpy()[[1, 4]])
    dense_out = mx.nd.zeros((6, 3))
    kv.row_sparse_pull("emb", out=dense_out, row_ids=mx.nd.array([2], dtype="int64"))
    assert np.array_equal(dense_out.asnumpy()[2], w.asnumpy()[2]) and dense_out.asnumpy()[[0, 1, 3, 4, 5]].sum() == 0
    kv.push("emb", rs)                      # without an updater the store takes the pushed value (densified)
    chk = mx.nd.zeros((6, 3)); kv.pull("emb", out=chk)
    assert np.array_equal(chk.asnumpy(), dense)


def test_recordio_roundtrip(tmp_path):
    """mx.recordio: sequential + indexed files, payloads containing the magic word, image records (python/mxnet/recordio.py)."""
    import struct
    import numpy as np
    from geomx_b200 import recordio
    p = str(tmp_path / "a.rec")
    w = recordio.MXRecordIO(p, "w")
    payloads = [b"hello", b"", struct.pack("<I", 0xced7230a) * 2 + b"xyz", bytes(range(256)) * 3]
    for b in payloads:
        w.write(b)
    w.close()
    r = recordio.MXRecordIO(p, "r")
    assert [r.read() for _ in payloads] == payloads and r.read() is None
    wi = recordio.MXIndexedRecordIO(str(tmp_path / "b.idx"), str(tmp_path / "b.rec"), "w")
    img = (np.arange(8 * 8 * 3) % 255).astype(np.uint8).reshape(8, 8, 3)
    for i in range(5):
        wi.write_idx(i, recordio.pack_img(recordio.IRHeader(0, float(i), i, 0), img, img_fmt=".png"))
    wi.write_idx(7, recordio.pack(recordio.IRHeader(0, [1.0, 2.0, 3.0], 7, 0), b"raw"))
    wi.close()
    ri = recordio.MXIndexedRecordIO(str(tmp_path / "b.idx"), str(tmp_path / "b.rec"), "r")
    assert ri.keys == [0, 1, 2, 3, 4, 7]
    h, dec = recordio.unpack_img(ri.read_idx(3))
    assert h.label == 3.0 and h.id == 3 and np.array_equal(dec, img)
    h, raw = recordio.unpack(ri.read_idx(7))
    assert list(h.label) == [1.0, 2.0, 3.0] and raw == b"raw"
    # native reader: memory-mapped scan + random access by offset (multi-chunk records re-assembled), several records per call
    rd = recordio.RecordReader(p)
    assert len(rd) == len(payloads) and rd.offsets == recordio._scan_python(p) and [rd.read(o) for o in rd.offsets] == payloads
    assert rd.read_many(rd.offsets[::-1], threads=3) == payloads[::-1]
    with pytest.raises(Exception):
        rd.read(rd.offsets[-1] + 4)                       # not the start of a record
    # the offsets are what an .idx file stores: an un-indexed .rec is iterated lazily through them by ImageIter
    assert recordio.scan_offsets(str(tmp_path / "b.rec")) == [ri.idx[k] for k in ri.keys]
    wc = recordio.MXRecordIO(str(tmp_path / "c.rec"), "w")
    for i in range(5):
        wc.write(recordio.pack_img(recordio.IRHeader(0, float(i), i, 0), np.full((8, 8, 3), 20 * i, dtype=np.uint8), img_fmt=".png"))
    wc.close()
    it = mx.image.ImageIter(batch_size=2, data_shape=(3, 8, 8), path_imgrec=str(tmp_path / "c.rec"), aug_list=[])
    got = [(b.data[0].asnumpy()[:, 0, 0, 0].tolist(), b.label[0].asnumpy().tolist(), b.pad) for b in it]
    assert got == [([0.0, 20.0], [0.0, 1.0], 0), ([40.0, 60.0], [2.0, 3.0], 0), ([80.0, 0.0], [4.0, 0.0], 1)]


def test_symbol_and_module_fit(tmp_path):
    """mx.sym graph + mx.mod.Module (bind / init_params / init_optimizer(kvstore) / fit / score / predict / checkpoint) on a separable toy task,
    single context and two (CPU) contexts with a 'local' kvstore (module.py:363-668, executor_group.py)."""
    import numpy as np
    rng = np.random.RandomState(0)
    X = rng.randn(256, 1, 8, 8).astype(np.float32)
    y = (X.reshape(256, -1)[:, :32].sum(1) > X.reshape(256, -1)[:, 32:].sum(1)).astype(np.float32)
    data = mx.sym.Variable("data")
    net = mx.sym.Convolution(data, kernel=(3, 3), num_filter=4, name="c0")
    net = mx.sym.Activation(net, "relu")
    net = mx.sym.Pooling(net, kernel=(2, 2), stride=(2, 2), pool_type="max")
    net = mx.sym.Flatten(net)
    net = mx.sym.FullyConnected(net, num_hidden=16, name="fc0")
    net = mx.sym.Activation(net, "relu")
    net = mx.sym.FullyConnected(net, num_hidden=2, name="fc1")
    net = mx.sym.SoftmaxOutput(net, name="softmax")
    assert net.list_arguments() == ["data", "c0_weight", "c0_bias", "fc0_weight", "fc0_bias", "fc1_weight", "fc1_bias", "softmax_label"]
    args, outs, _ = net.infer_shape(data=(32, 1, 8, 8))
    assert outs == [(32, 2)] and args[3] == (16, 36)
    for ctxs, kvs in (([mx.cpu()], None), ([mx.cpu(0), mx.cpu(1)], "local")):
        it = mx.io.NDArrayIter(X, y, batch_size=32, shuffle=False)
        mod = mx.mod.Module(net, context=ctxs)
        mod.fit(it, num_epoch=12, optimizer="adam", optimizer_params={"learning_rate": 0.01}, kvstore=kvs, initializer=mx.init.Xavier())
        acc = dict(mod.score(mx.io.NDArrayIter(X, y, batch_size=32), "acc"))["accuracy"]
        assert acc > 0.9, acc
        pred = mod.predict(mx.io.NDArrayIter(X, y, batch_size=32))
        assert pred.shape == (256, 2)
    prefix = str(tmp_path / "toy")
    mod.save_checkpoint(prefix, 3)
    mod2 = mx.mod.Module.load(prefix, 3)
    mod2.bind(data_shapes=[("data", (32, 1, 8, 8))], label_shapes=[("softmax_label", (32,))], for_training=False)
    mod2.init_params()
    acc2 = dict(mod2.score(mx.io.NDArrayIter(X, y, batch_size=32), "acc"))["accuracy"]
    assert abs(acc2 - acc) < 1e-6


def test_text_and_record_iterators(tmp_path):
    """mx.io.CSVIter / LibSVMIter (native parsers, csrc/runtime/text_io.h), ImageRecordIter over a RecordIO file, PrefetchingIter."""
    import numpy as np
    from geomx_b200 import recordio
    d = np.arange(40, dtype=np.float32).reshape(10, 4) / 7
    np.savetxt(tmp_path / "d.csv", d, delimiter=",", fmt="%.6f")
    np.savetxt(tmp_path / "l.csv", np.arange(10), fmt="%d")
    it = mx.io.CSVIter(data_csv=str(tmp_path / "d.csv"), data_shape=(4,), label_csv=str(tmp_path / "l.csv"), batch_size=5)
    b = next(iter(it))
    assert b.data[0].shape == (5, 4) and np.allclose(b.data[0].asnumpy(), d[:5], atol=1e-5) and b.label[0].asnumpy().tolist() == [0, 1, 2, 3, 4]
    (tmp_path / "s.libsvm").write_text("1 0:1.5 3:2\n0 2:-1\n# comment\n1 1:4 2:5 3:6\n")
    it = mx.io.LibSVMIter(data_libsvm=str(tmp_path / "s.libsvm"), data_shape=(4,), batch_size=3)
    b = next(iter(it))
    assert np.array_equal(b.data[0].asnumpy(), np.array([[1.5, 0, 0, 2], [0, 0, -1, 0], [0, 4, 5, 6]], dtype=np.float32))
    assert b.label[0].asnumpy().tolist() == [1, 0, 1]
    w = recordio.MXIndexedRecordIO(str(tmp_path / "i.idx"), str(tmp_path / "i.rec"), "w")
    for i in range(6):
        img = np.full((12, 12, 3), 10 * i, dtype=np.uint8)
        w.write_idx(i, recordio.pack_img(recordio.IRHeader(0, float(i % 2), i, 0), img, img_fmt=".png"))
    w.close()
    it = mx.io.ImageRecordIter(path_imgrec=str(tmp_path / "i.rec"), path_imgidx=str(tmp_path / "i.idx"), data_shape=(3, 8, 8), batch_size=4,
                               rand_crop=True, rand_mirror=True, scale=1 / 255.0)
    pf = mx.io.PrefetchingIter(it)
    batches = list(pf)
    assert len(batches) == 2 and batches[0].data[0].shape == (4, 3, 8, 8) and batches[1].pad == 2
    assert np.allclose(batches[0].data[0].asnumpy()[1], 10 / 255.0) and batches[0].label[0].asnumpy().tolist() == [0, 1, 0, 1]
    pf.reset()
    assert len(list(pf)) == 2


def test_sync_batchnorm_two_ranks():
    """SyncBatchNorm over 2 gloo ranks == BatchNorm over the concatenated batch (forward, input grad, gamma grad, running stats)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ); env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29877", os.path.join(here, "_syncbn_worker.py")], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=180)
    assert "SYNCBN PASS" in r.stdout, r.stdout[-2000:]


def test_host_storage_pool_and_resources():
    """mx.storage: size-bucketed host pool (reuse, rounding, release) and reproducible per-device seed streams (storage.cc / resource.cc)."""
    import gc
    import torch
    from geomx_b200 import runtime, storage
    if not runtime.available():
        pytest.skip("native runtime not built")
    pool = storage.HostPool(pin=False)
    a = pool.empty(1000, torch.float32)
    a.fill_(3.0)
    assert a.numel() == 1000 and float(a.sum()) == 3000.0
    p0 = a.data_ptr()
    st = pool.stats()
    assert st["used_bytes"] == 4096 and st["num_system_alloc"] == 1
    view = a[10:20]
    del a; gc.collect()
    assert pool.stats()["used_bytes"] == 4096                   # a view keeps the block alive
    del view; gc.collect()
    b = pool.empty(900, torch.float32)                         # same bucket: served from the pool
    assert b.data_ptr() == p0 and pool.stats()["num_pool_hits"] == 1
    big = pool.empty((3 << 20) // 4 + 1, torch.float32)        # > 1 MiB: power-of-two bucket
    assert pool.stats()["used_bytes"] == 4096 + (4 << 20)
    del b, big; gc.collect()
    assert pool.stats()["used_bytes"] == 0 and pool.stats()["pooled_bytes"] == 4096 + (4 << 20)
    pool.release_all()
    assert pool.stats()["pooled_bytes"] == 0
    pool.seed(42); s1 = [pool.next_seed(0), pool.next_seed(0), pool.next_seed(1)]
    pool.seed(42); s2 = [pool.next_seed(0), pool.next_seed(0), pool.next_seed(1)]
    assert s1 == s2 and len(set(s1)) == 3


def test_callbacks_monitor_and_test_utils(caplog):
    """mx.callback.Speedometer / mx.monitor.Monitor on a Module, mx.test_utils numeric-gradient check against autograd."""
    import logging
    import numpy as np
    X = np.random.RandomState(1).randn(64, 6).astype(np.float32); y = (X[:, 0] > 0).astype(np.float32)
    net = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.Variable("data"), num_hidden=2, name="fc"), name="softmax")
    mod = mx.mod.Module(net)
    speed = mx.callback.Speedometer(16, frequent=2)
    with caplog.at_level(logging.INFO):
        mod.fit(mx.io.NDArrayIter(X, y, batch_size=16), num_epoch=2, optimizer_params={"learning_rate": 0.1}, batch_end_callback=speed)
    assert speed.last_speed is not None and any("samples/sec" in r.getMessage() for r in caplog.records)
    mon = mx.monitor.Monitor(1, pattern="fc_.*")
    mon.install(mod._execs[0]); mon.tic()
    mod.forward_backward(next(iter(mx.io.NDArrayIter(X, y, batch_size=16))))
    names = [k for _, k, _ in mon.toc()]
    assert "fc_weight" in names and "fc_weight_grad" in names
    mx.test_utils.check_numeric_gradient(lambda a, w: mx.nd.dot(a, w).tanh() * 2.0, [np.random.randn(3, 4), np.random.randn(4, 2)])
    mx.test_utils.assert_almost_equal(mx.nd.array([1.0, 2.0]), np.array([1.0, 2.0 + 1e-7]))


def test_cpu_shared_and_pinned_contexts():
    """mx.cpu_shared(): POSIX shared memory another process can map (CPUSharedStorageManager, the DataLoader hand-off); mx.cpu_pinned():
    page-locked memory when a CUDA driver is present, plain host memory otherwise."""
    import torch.multiprocessing as tmp
    a = mx.nd.array(np.arange(6, dtype=np.float32).reshape(2, 3), ctx=mx.cpu_shared())
    assert str(a.context) == "cpu_shared(0)" and a._t.is_shared()
    b = mx.nd.zeros((4,), ctx=mx.cpu_shared(0))
    ctx = tmp.get_context("spawn")
    p = ctx.Process(target=_write_shared, args=(b._t,))
    p.start(); p.join(60)
    assert p.exitcode == 0 and b.asnumpy().tolist() == [1.0, 2.0, 3.0, 4.0]      # the child wrote into the same pages
    c = mx.nd.ones((3,)).as_in_context(mx.cpu_shared())
    assert c._t.is_shared() and str(c.context) == "cpu_shared(0)"
    d = mx.nd.array([1, 2, 3], ctx=mx.cpu_pinned())
    assert str(d.context) == "cpu_pinned(0)" and d._t.is_pinned() == torch.cuda.is_available()


def _write_shared(t):
    t.copy_(torch.tensor([1.0, 2.0, 3.0, 4.0]))


def test_engine_device_pools_exceptions_and_delete():
    """mx.engine: per-device worker pools (compute / copy per GPU, normal / priority on the CPU; threaded_engine_perdevice.cc), exceptions
    surface at the wait points (threaded_engine.h OnComplete / var exception), DeleteVariable after pending ops."""
    import threading
    import time
    from geomx_b200 import engine, runtime
    if not runtime.available():
        pytest.skip("native runtime not built")
    engine.wait_all()
    before = engine.stats()
    seen, lock = {}, threading.Lock()

    def tag(name, dt=0.0):
        def f():
            time.sleep(dt)
            with lock:
                seen.setdefault(name, threading.current_thread().ident)
        return f
    v0, v1 = engine.new_variable(), engine.new_variable()
    # a slow op on GPU 0's compute pool must not delay GPU 1's pool nor the copy pool of GPU 0
    t0 = time.time()
    for _ in range(4):
        engine.push(tag("g0", 0.15), mutable_vars=[], ctx=mx.gpu(0))          # 4 x 0.15 s on 2 threads = 0.3 s of queueing in that pool
    engine.push(tag("g1"), mutable_vars=[v1], ctx=mx.gpu(1))
    engine.push(tag("c0"), mutable_vars=[v0], ctx=0, prop=engine.COPY)
    engine.wait_for_var(v1); engine.wait_for_var(v0)
    assert time.time() - t0 < 0.14, "independent pools were blocked behind gpu0's queue"
    engine.push(tag("prio"), prop=engine.PRIORITY); engine.push(tag("cpu"))
    engine.wait_all()
    after = engine.stats()
    delta = {k: after.get(k, 0) - before.get(k, 0) for k in after}
    assert delta["gpu0"] == 4 and delta["gpu1"] == 1 and delta["gpu0/copy"] == 1 and delta["cpu"] >= 1 and delta["priority"] >= 1
    assert len({seen["g0"], seen["g1"], seen["c0"]}) == 3                       # three different worker threads
    # exceptions: remembered on the written variable, raised by the next wait on it; later ops still run
    def boom():
        raise ValueError("op failed")
    ran = []
    engine.push(boom, mutable_vars=[v0])
    engine.push(lambda: ran.append(1), mutable_vars=[v0])
    with pytest.raises(Exception, match="op failed"):
        engine.wait_for_var(v0)
    assert ran == [1]
    engine.wait_for_var(v0)                                                      # reported once
    engine.push(boom, mutable_vars=[v1])
    with pytest.raises(Exception, match="op failed"):
        engine.wait_all()
    engine.wait_all()
    # DeleteVariable: after the ops already pushed on it
    n = engine.get().num_variables()
    v2 = engine.new_variable()
    engine.push(tag("last", 0.02), mutable_vars=[v2])
    engine.delete_variable(v2)
    engine.wait_all()
    assert "last" in seen and engine.get().num_variables() == n


def test_engine_ordering_and_async_save(tmp_path):
    """mx.engine: writers are exclusive and ordered, readers run in between, priorities order ready ops; nd.save_async writes a snapshot."""
    import threading
    import time
    import numpy as np
    from geomx_b200 import engine, runtime
    if not runtime.available():
        pytest.skip("native runtime not built")
    v = engine.new_variable()
    log, lock = [], threading.Lock()

    def op(tag, dt=0.0):
        def f():
            time.sleep(dt)
            with lock:
                log.append(tag)
        return f
    engine.push(op("w1", 0.05), mutable_vars=[v])
    engine.push(op("r1"), const_vars=[v]); engine.push(op("r2"), const_vars=[v])
    engine.push(op("w2"), mutable_vars=[v])
    engine.push(op("r3"), const_vars=[v])
    engine.wait_for_var(v)                        # a read of v: returns once w2 is done (r3 may still be running next to it)
    assert log[0] == "w1" and set(log[1:3]) == {"r1", "r2"} and log[3] == "w2"
    engine.wait_all()
    assert log[4] == "r3"
    with pytest.raises(Exception):
        engine.push(op("bad"), const_vars=[v], mutable_vars=[v])
    a = mx.nd.array(np.arange(6, dtype=np.float32))
    path = str(tmp_path / "ck.params")
    mx.nd.save_async(path, {"w": a})
    a[:] = -1.0                                   # later in-place updates must not leak into the checkpoint
    mx.nd.save_async(path + "2", [a])
    mx.nd.waitall()
    assert mx.nd.load(path)["w"].asnumpy().tolist() == [0, 1, 2, 3, 4, 5] and mx.nd.load(path + "2")[0].asnumpy().tolist() == [-1.0] * 6
