"""mx.operator (CustomOp), mx.random, mx.visualization, mx.image, extra metrics."""
import numpy as np

import geomx_b200 as mx


def test_custom_op_forward_backward():
    @mx.operator.register("scaled_sigmoid")
    class ScaledSigmoidProp(mx.operator.CustomOpProp):
        def __init__(self, scale="1.0"):
            super().__init__(need_top_grad=True)
            self.scale = float(scale)

        def list_arguments(self): return ["data"]
        def list_outputs(self): return ["output"]
        def infer_shape(self, in_shape): return in_shape, [in_shape[0]], []

        def create_operator(self, ctx, shapes, dtypes):
            scale = self.scale

            class Op(mx.operator.CustomOp):
                def forward(self, is_train, req, in_data, out_data, aux):
                    self.assign(out_data[0], req[0], mx.nd.sigmoid(in_data[0]) * scale)

                def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
                    y = out_data[0] / scale
                    self.assign(in_grad[0], req[0], out_grad[0] * scale * y * (1 - y))
            return Op()
    x = mx.nd.array(np.linspace(-2, 2, 12, dtype=np.float32).reshape(3, 4)); x.attach_grad()
    with mx.autograd.record():
        y = mx.nd.Custom(x, op_type="scaled_sigmoid", scale=3.0)
        loss = (y * y).sum()
    loss.backward()
    s = 1 / (1 + np.exp(-x.asnumpy()))
    assert np.allclose(y.asnumpy(), 3 * s, atol=1e-6) and np.allclose(x.grad.asnumpy(), 2 * 3 * s * 3 * s * (1 - s), atol=1e-5)
    assert "scaled_sigmoid" in mx.operator.get_all_registered_operators()


def test_random_seed_and_visualization(capsys):
    mx.random.seed(7); a = mx.random.uniform(shape=(5,)).asnumpy()
    mx.random.seed(7); b = mx.random.uniform(shape=(5,)).asnumpy()
    assert np.array_equal(a, b)
    net = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.Activation(mx.sym.Convolution(mx.sym.Variable("data"), kernel=(3, 3), num_filter=4, name="c"),
                                                                       "relu"), num_hidden=10, name="fc"), name="softmax")
    total = mx.viz.print_summary(net, shape={"data": (1, 1, 8, 8)})
    assert total == 4 * 9 + 4 + 10 * 4 * 36 + 10 and "fc(FullyConnected)" in capsys.readouterr().out


def test_image_pipeline(tmp_path):
    from PIL import Image
    rng = np.random.RandomState(0)
    files = []
    for i in range(5):
        arr = rng.randint(0, 255, size=(20 + i, 30, 3)).astype(np.uint8)
        p = tmp_path / ("img%d.png" % i); Image.fromarray(arr).save(p); files.append((arr, str(p)))
    img = mx.image.imread(files[0][1])
    assert img.shape == (20, 30, 3) and np.array_equal(img.asnumpy(), files[0][0])
    assert mx.image.resize_short(img, 10).shape[:2] == (10, 15)
    crop, box = mx.image.center_crop(img, (16, 16))
    assert crop.shape == (16, 16, 3) and box == (7, 2, 16, 16)
    it = mx.image.ImageIter(batch_size=2, data_shape=(3, 16, 16), imglist=[[float(i % 2), f[1]] for i, f in enumerate(files)], path_root="",
                            rand_crop=True, rand_mirror=True, mean=True, std=True)
    batches = list(it)
    assert len(batches) == 3 and batches[0].data[0].shape == (2, 3, 16, 16) and batches[2].pad == 1
    assert batches[0].label[0].asnumpy().tolist() == [0.0, 1.0]


def test_rtc_compiles_for_sm100a_without_a_gpu():
    """mx.rtc: NVRTC produces an sm_100a cubin on a GPU-less box; syntax errors surface as MXNetError with the compiler log."""
    import pytest
    try:
        mod = mx.rtc.CudaModule('extern "C" __global__ void axpy(const float* x, float* y, float a, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] += a * x[i]; }')
    except ImportError:
        pytest.skip("cuda-python not installed")
    assert len(mod.cubin) > 1000
    k = mod.get_kernel("axpy", "const float *x, float *y, float a, int n")
    assert [p[0] for p in k._params] == [True, True, False, False]
    with pytest.raises(mx.MXNetError):
        mx.rtc.CudaModule("this is not CUDA")


def test_bucketing_and_sequential_modules():
    import numpy as np
    import geomx_b200 as mx

    # ---- bucketing: per-length graphs sharing one set of weights
    def sym_gen(T):
        d = mx.sym.Variable("data")                                   # (N, T) token ids in [0, 8)
        emb = mx.sym.one_hot(d, depth=8)                              # (N, T, 8) via the generic nd bridge
        h = mx.sym.mean(emb, axis=1)                                  # (N, 8): bag of words, independent of T
        out = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(h, num_hidden=2, name="fc"), name="softmax")
        return out, ("data",), ("softmax_label",)

    rs = np.random.RandomState(0)

    def batch(T, n=16):
        y = rs.randint(0, 2, n)
        x = np.where(rs.rand(n, T) < 0.7, y[:, None] * 4 + rs.randint(0, 4, (n, T)), rs.randint(0, 8, (n, T)))   # class decides the token range
        return mx.io.DataBatch([mx.nd.array(x.astype(np.float32))], [mx.nd.array(y.astype(np.float32))], bucket_key=T,
                               provide_data=[mx.io.DataDesc("data", (n, T), "float32", "NT")], provide_label=[mx.io.DataDesc("softmax_label", (n,), "float32", "N")])

    bm = mx.mod.BucketingModule(sym_gen, default_bucket_key=10)
    bm.bind([("data", (16, 10))], [("softmax_label", (16,))])
    bm.init_params(mx.init.Xavier())
    bm.init_optimizer(optimizer="adam", optimizer_params={"learning_rate": 0.05})
    for step in range(60):
        b = batch([5, 10, 7][step % 3])
        bm.forward_backward(b); bm.update()
    assert set(bm._buckets) == {5, 7, 10}
    w10 = bm._buckets[10]._execs[0].arg_dict["fc_weight"]; w5 = bm._buckets[5]._execs[0].arg_dict["fc_weight"]
    assert w10 is w5                                                   # literally the same array
    m = mx.metric.Accuracy()
    for T in (5, 7, 10):
        b = batch(T, 16)
        bm.forward(b, is_train=False); bm.update_metric(m, b.label)
    assert m.get()[1] > 0.9

    # ---- sequential: two modules chained, gradients cross the boundary
    d = mx.sym.Variable("data")
    m1 = mx.mod.Module(mx.sym.Activation(mx.sym.FullyConnected(d, num_hidden=8, name="a"), act_type="relu"), label_names=None)
    m2 = mx.mod.Module(mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.Variable("data"), num_hidden=2, name="b"), name="softmax"))
    seq = mx.mod.SequentialModule().add(m1).add(m2, take_labels=True, auto_wiring=True)
    x = rs.randn(64, 4).astype(np.float32); y = (x[:, 0] + x[:, 1] > 0).astype(np.float32)
    it = mx.io.NDArrayIter(x, y, batch_size=16)
    seq.fit(it, num_epoch=25, optimizer="adam", optimizer_params={"learning_rate": 0.05}, initializer=mx.init.Xavier())
    assert dict(seq.score(it, "acc"))["accuracy"] > 0.9
    arg, _ = seq.get_params()
    assert {"a_weight", "b_weight"} <= set(arg)


def test_rnn_bucket_iter_registry_log_features(tmp_path):
    import geomx_b200 as mx
    sents = [["a", "b", "c"], ["a", "b"], ["c", "c", "c", "a", "b"], ["b", "a", "c"], ["a"], ["b", "b"]] * 4
    enc, vocab = mx.rnn.encode_sentences(sents, invalid_label=0, start_label=1)
    assert vocab["a"] == 1 and enc[0] == [1, 2, 3]
    it = mx.rnn.BucketSentenceIter(enc, batch_size=4, buckets=[2, 3, 5], invalid_label=0)
    keys = []
    for b in it:
        assert b.data[0].shape == (4, b.bucket_key) and b.provide_data[0].shape == (4, b.bucket_key)
        d, l = b.data[0].asnumpy(), b.label[0].asnumpy()
        assert (l[:, :-1] == d[:, 1:]).all() and (l[:, -1] == 0).all()
        keys.append(b.bucket_key)
    assert sorted(set(keys)) == [2, 3, 5] and it.default_bucket_key == 5

    class Base:                                                         # registry helpers
        pass
    reg = mx.registry.get_register_func(Base, "thing"); alias = mx.registry.get_alias_func(Base, "thing"); create = mx.registry.get_create_func(Base, "thing")

    @alias("w", "widget2")
    @reg
    class Widget(Base):
        def __init__(self, size=1):
            self.size = size
    assert create("widget", size=3).size == 3 and create('["w", {"size": 5}]').size == 5 and isinstance(create(Widget()), Widget)
    assert set(mx.registry.get_registry(Base)) == {"widget", "w", "widget2"}

    log = mx.log.get_logger("geomx_test_log", filename=str(tmp_path / "x.log"), level=mx.log.INFO)
    log.info("hello %d", 3)
    assert "hello 3" in open(str(tmp_path / "x.log")).read()
    f = mx.runtime.Features()
    assert f.is_enabled("DIST_KVSTORE") == mx.runtime.available() and "TCGEN05" in f
    assert mx.executor.Executor is mx.symbol.Executor and mx.libinfo.find_include_path().endswith("csrc")


def test_profiler_trace_objects_and_hooks(tmp_path):
    import json
    import time
    import numpy as np
    import geomx_b200 as mx
    prof = mx.profiler
    fn = str(tmp_path / "trace.json")
    prof.set_config(filename=fn, profile_all=True, aggregate_stats=True)
    prof.set_state("run")
    net = mx.gluon.nn.Sequential()
    net.add(mx.gluon.nn.Dense(8, activation="relu"), mx.gluon.nn.Dense(2))
    net.initialize()
    net(mx.nd.array(np.ones((4, 3), dtype=np.float32)))
    kv = mx.kv.create("local")
    kv.init(3, mx.nd.ones((2,))); kv.push(3, mx.nd.ones((2,))); out = mx.nd.zeros((2,)); kv.pull(3, out=out)
    d = mx.sym.Variable("data")
    ex = mx.sym.FullyConnected(d, num_hidden=2, name="fcp").simple_bind(mx.cpu(), data=(1, 3))
    ex.forward()
    dom = prof.Domain("app")
    with dom.new_task("stage"):
        time.sleep(0.002)
    c = dom.new_counter("items", 5); c += 3; c -= 1
    dom.new_marker("here").mark()
    prof.pause()
    net(mx.nd.array(np.ones((4, 3), dtype=np.float32)))                # not recorded while paused
    prof.resume()
    table = prof.dumps()
    assert "KVStorePush" in table and "stage" in table
    agg = json.loads(prof.dumps(format="json"))
    assert agg["KVStorePush"]["count"] == 1 and agg["stage"]["total_us"] >= 1500
    prof.dump()
    tr = json.load(open(fn))
    names = [e["name"] for e in tr["traceEvents"]]
    assert {"KVStorePush", "KVStorePull", "fcp", "stage", "items", "here"} <= set(names)
    assert sum(1 for e in tr["traceEvents"] if e.get("cat") == "block") == 3        # Sequential + two Dense, once
    counters = [e for e in tr["traceEvents"] if e["name"] == "items"]
    assert [list(e["args"].values())[0] for e in counters][-1] == 7
    assert all(e["ph"] in ("X", "C", "i", "M", "B", "E") for e in tr["traceEvents"])
    assert not prof.is_active()                                         # dump(finished=True) stops the profiler
    # continuous dump keeps rewriting the file while running
    fn2 = str(tmp_path / "cont.json")
    prof.set_config(filename=fn2, continuous_dump=True, dump_period=0.05)
    prof.set_state("run")
    with prof.scope("tick"):
        pass
    time.sleep(0.3)
    assert "tick" in open(fn2).read()
    prof.set_config(continuous_dump=False); prof.set_state("stop")


def _torchrun_cpu(nproc, env, timeout=240):
    import json
    import os
    import socket
    import subprocess
    import sys
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    e = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    for k in ("DMLC_ROLE", "DMLC_PS_ROOT_URI", "DMLC_ROLE_GLOBAL", "DMLC_PS_GLOBAL_ROOT_URI", "RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env)
    import tempfile
    outdir = tempfile.mkdtemp()
    e["TEST_OUT_DIR"] = outdir
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(here, "_collective_worker.py")], env=e, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    res = [json.load(open(os.path.join(outdir, f))) for f in sorted(os.listdir(outdir))]
    assert len(res) == nproc, (r.stdout + r.stderr)[-2000:]
    return sorted(res, key=lambda d: d["rank"])


def test_collective_kvstore_under_torchrun_on_cpu():
    """torchrun without GPUs: ``mx.kv.create('dist_sync')`` is the gloo collective store with the fabric store's two-tier semantics."""
    import pytest
    res = _torchrun_cpu(4, {"TEST_MODE": "sgd", "GEOMX_NUM_PARTIES": "2"})
    gsum = 0.5 * (1 + 2 + 3 + 4)
    for r in res:
        assert r["type"] == "KVStoreCollective" and r["num_workers"] == 2 and r["num_all_workers"] == 4 and r["party_rank"] == r["rank"] % 2
        assert r["init"] == [1.0, 2.0]                                            # world rank 0's values win
        for t, vals in enumerate(r["vals"]):
            assert vals == pytest.approx([1.0 - 0.1 * gsum * (t + 1), 2.0 - 0.1 * gsum * (t + 1)], abs=1e-5)
    # HFA, K2 = 2: round 1 stops at the party tier (party sums of value/num_workers), round 2 is the mean over parties of the party values
    res = _torchrun_cpu(4, {"TEST_MODE": "hfa", "GEOMX_NUM_PARTIES": "2", "MXNET_KVSTORE_USE_HFA": "1", "MXNET_KVSTORE_HFA_K2": "2"})
    assert [r["vals"][0][0] for r in res] == pytest.approx([1.5, 1.5, 3.5, 3.5])
    assert all(r["vals"][1][0] == pytest.approx((2 * 1.5 + 2 * 3.5) / 2) for r in res)
    # MixedSync: SGD is linear, so applying the party aggregates one after the other equals the synchronous result
    res = _torchrun_cpu(4, {"TEST_MODE": "async", "GEOMX_NUM_PARTIES": "2", "TEST_STEPS": "1"})
    assert all(r["vals"][0][0] == pytest.approx(1.0 - 0.1 * gsum, abs=1e-5) for r in res)


def test_optimizer_spec_is_static_and_adam_clip_order():
    """the native server spec is only used when it is complete (no scheduler / multipliers); Adam clips grad + wd*w like adam_update."""
    import numpy as np
    import geomx_b200 as mx
    o = mx.optimizer.Adam(learning_rate=0.01)
    assert o.spec_is_static()
    assert not mx.optimizer.Adam(learning_rate=0.01, lr_scheduler=mx.lr_scheduler.FactorScheduler(step=10, factor=0.5)).spec_is_static()
    o2 = mx.optimizer.SGD(learning_rate=0.1, wd=1e-4, param_idx2name={0: "fc_weight", 1: "fc_bias"})
    assert not o2.spec_is_static()            # wd_mult = 0 on the bias
    o3 = mx.optimizer.SGD(learning_rate=0.1, param_idx2name={0: "fc_weight", 1: "fc_bias"})
    assert o3.spec_is_static()                # no weight decay: the multipliers are irrelevant
    # Adam, one step, clip below |g + wd*w|: m = (1-b1)*clip(g + wd*w)
    w = mx.nd.array(np.array([10.0, -10.0], dtype=np.float32)); g = mx.nd.array(np.array([0.5, -0.5], dtype=np.float32))
    opt = mx.optimizer.Adam(learning_rate=0.1, wd=0.1, clip_gradient=1.0)
    st = opt.create_state(0, w)
    opt.update(0, w, g, st)
    m = st[0].asnumpy()
    assert np.allclose(m, 0.1 * np.array([1.0, -1.0]), atol=1e-6), m      # clip(0.5 + 1.0) = 1.0, not clip(0.5) + 1.0 = 1.5


def test_group2ctx_model_parallel_plumbing():
    """mx.AttrScope(ctx_group=...) + simple_bind(group2ctx=...): annotated variables are allocated on their group's device, operators run there,
    cross-device copies are differentiable; results equal the single-device executor (reference: PlaceDevice + _CrossDeviceCopy)."""
    import numpy as np
    import geomx_b200 as mx
    with mx.AttrScope(ctx_group="stage1"):
        data = mx.sym.Variable("data")
        act = mx.sym.Activation(mx.sym.FullyConnected(data, num_hidden=8, name="fc1"), act_type="relu", name="act1")
    with mx.AttrScope(ctx_group="stage2"):
        out = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(act, num_hidden=3, name="fc2"), name="softmax")
    ex = out.simple_bind(mx.cpu(0), group2ctx={"stage1": mx.cpu(0), "stage2": mx.cpu(1)}, data=(4, 5), softmax_label=(4,))
    assert ex.arg_dict["fc2_weight"].context == mx.cpu(1) and ex.arg_dict["fc1_weight"].context == mx.cpu(0)
    rng = np.random.RandomState(3)
    for n, a in ex.arg_dict.items():
        a[:] = (rng.randint(0, 3, a.shape) if n == "softmax_label" else rng.randn(*a.shape) * 0.3).astype(np.float32)
    ref = out.simple_bind(mx.cpu(0), data=(4, 5), softmax_label=(4,))
    for n in ex.arg_dict:
        ref.arg_dict[n][:] = ex.arg_dict[n].asnumpy()
    o = ex.forward(is_train=True); ex.backward()
    r = ref.forward(is_train=True); ref.backward()
    assert np.allclose(o[0].asnumpy(), r[0].asnumpy(), atol=1e-6)
    for n in ("fc1_weight", "fc2_weight", "fc1_bias"):
        assert np.allclose(ex.grad_dict[n].asnumpy(), ref.grad_dict[n].asnumpy(), atol=1e-6)
