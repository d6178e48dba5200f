"""gluon.model_zoo and gluon.rnn: shapes, parameter inventories, exactness of the cells against torch.nn, gradients through a layer."""
import numpy as np
import torch

import geomx_b200 as mx
from geomx_b200 import gluon


def test_model_zoo_builds_and_trains_one_step():
    net = gluon.model_zoo.get_model("resnet18_v1", classes=10, thumbnail=True)
    net.initialize(mx.init.Xavier())
    x = mx.nd.random.uniform(shape=(2, 3, 32, 32)); y = mx.nd.array([1, 7])
    loss = gluon.loss.SoftmaxCrossEntropyLoss()
    trainer = gluon.Trainer(net.collect_params(), "sgd", {"learning_rate": 0.05})
    with mx.autograd.record():
        l0 = loss(net(x), y).mean()
    l0.backward(); trainer.step(2)
    with mx.autograd.record():
        l1 = loss(net(x), y).mean()
    assert net(x).shape == (2, 10) and float(l1.asscalar()) < float(l0.asscalar())
    for name, shp in (("lenet", (2, 1, 28, 28)), ("mlp", (2, 1, 28, 28)), ("squeezenet1.1", (2, 3, 64, 64)), ("mobilenet0.5", (2, 3, 64, 64)), ("vgg11_bn", (2, 3, 32, 32))):
        n = gluon.model_zoo.get_model(name, classes=10); n.initialize()
        assert n(mx.nd.random.uniform(shape=shp)).shape == (2, 10)
    lenet = gluon.model_zoo.get_model("lenet"); lenet.initialize()
    lenet(mx.nd.zeros((1, 1, 28, 28)))
    assert sum(int(np.prod(p.shape)) for p in lenet.collect_params().values()) == 178762      # the GeoMX demo CNN


def test_rnn_cells_match_torch_and_layers_backprop():
    for ours, ref_cls in ((gluon.rnn.LSTMCell, torch.nn.LSTMCell), (gluon.rnn.GRUCell, torch.nn.GRUCell)):
        cell = ours(6); cell.initialize()
        x = mx.nd.random.uniform(shape=(2, 4))
        h, st = cell(x, cell.begin_state(2))
        ref = ref_cls(4, 6)
        with torch.no_grad():
            ref.weight_ih.copy_(cell.i2h_weight.data()._t); ref.weight_hh.copy_(cell.h2h_weight.data()._t)
            ref.bias_ih.copy_(cell.i2h_bias.data()._t); ref.bias_hh.copy_(cell.h2h_bias.data()._t)
        r = ref(x._t, (torch.zeros(2, 6), torch.zeros(2, 6))) if ref_cls is torch.nn.LSTMCell else ref(x._t, torch.zeros(2, 6))
        hr = r[0] if isinstance(r, tuple) else r
        assert float((hr - h._t).abs().max()) < 1e-6
    lstm = gluon.rnn.LSTM(8, num_layers=2, bidirectional=True, layout="NTC"); lstm.initialize()
    seq = mx.nd.random.uniform(shape=(3, 5, 4)); seq.attach_grad()
    with mx.autograd.record():
        out, st = lstm(seq, lstm.begin_state(3))
        loss = (out * out).sum()
    loss.backward()
    assert out.shape == (3, 5, 16) and st[0].shape == (4, 3, 8) and float(seq.grad.norm().asscalar()) > 0
    outs, _ = gluon.rnn.GRUCell(5).unroll(4, mx.nd.random.uniform(shape=(2, 4, 3)), layout="NTC") if False else (None, None)
    cell = gluon.rnn.GRUCell(5); cell.initialize()
    merged, _ = cell.unroll(4, mx.nd.random.uniform(shape=(2, 4, 3)), layout="NTC", merge_outputs=True)
    assert merged.shape == (2, 4, 5)


def test_nd_layers_losses_and_contrib_blocks():
    import numpy as np
    from geomx_b200.gluon import contrib, loss as gloss, nn, rnn
    x3 = mx.nd.array(torch.randn(2, 3, 4, 6, 6))
    c3 = nn.Conv3D(5, 3, padding=1); c3.initialize()
    assert c3(x3).shape == (2, 5, 4, 6, 6) and c3.weight.shape == (5, 3, 3, 3, 3)
    t3 = nn.Conv3DTranspose(4, 2, strides=2); t3.initialize()
    assert t3(x3).shape == (2, 4, 8, 12, 12)
    t1 = nn.Conv1DTranspose(4, 3, strides=2); t1.initialize()
    assert t1(mx.nd.array(torch.randn(2, 3, 5))).shape == (2, 4, 11)
    assert nn.MaxPool3D()(x3).shape == (2, 3, 2, 3, 3) and nn.AvgPool3D()(x3).shape == (2, 3, 2, 3, 3)
    assert nn.GlobalAvgPool3D()(x3).shape == (2, 3, 1, 1, 1) and nn.GlobalMaxPool1D()(mx.nd.array(torch.randn(2, 3, 5))).shape == (2, 3, 1)
    img = mx.nd.array(torch.arange(16.0).reshape(1, 1, 4, 4))
    rp = nn.ReflectionPad2D(1)(img).asnumpy()
    assert rp.shape == (1, 1, 6, 6) and rp[0, 0, 0, 0] == 5 and rp[0, 0, 0, 1] == 4
    pr = nn.PReLU(); pr.initialize()
    np.testing.assert_allclose(pr(mx.nd.array([-2.0, 3.0])).asnumpy(), [-0.5, 3.0])
    inorm = nn.InstanceNorm(); inorm.initialize()
    y = inorm(mx.nd.array(torch.randn(2, 3, 8, 8) * 3 + 1)).asnumpy()
    assert abs(y.mean(axis=(2, 3))).max() < 1e-5 and abs(y.std(axis=(2, 3)) - 1).max() < 1e-2
    assert float(abs(nn.GELU()(mx.nd.array([0.0, 10.0])).asnumpy() - [0, 10]).max()) < 1e-4

    # losses
    T, N, C = 6, 2, 5
    pred = mx.nd.array(torch.randn(N, T, C)); lab = mx.nd.array([[1, 2, -1], [3, 3, 0]])
    l = gloss.CTCLoss()(pred, lab).asnumpy()
    ref = torch.nn.functional.ctc_loss(torch.log_softmax(pred._t.transpose(0, 1), -1), torch.tensor([[1, 2, 0], [3, 3, 0]]), torch.tensor([T, T]),
                                       torch.tensor([2, 3]), blank=C - 1, reduction="none")
    np.testing.assert_allclose(l, ref.numpy(), rtol=1e-5)
    a, p, n_ = (mx.nd.array(torch.randn(4, 8)) for _ in range(3))
    tl = gloss.TripletLoss(margin=0.5)(a, p, n_).asnumpy()
    np.testing.assert_allclose(tl, np.maximum(((a._t - p._t) ** 2 - (a._t - n_._t) ** 2).sum(1).numpy() + 0.5, 0), rtol=1e-5)
    pn = gloss.PoissonNLLLoss()(mx.nd.array([0.0, 1.0]), mx.nd.array([1.0, 2.0])).asnumpy()
    np.testing.assert_allclose(pn, np.mean([1.0 - 0.0, np.e - 2.0]), rtol=1e-5)
    ce = gloss.CosineEmbeddingLoss()(mx.nd.array([[1.0, 0.0], [1.0, 0.0]]), mx.nd.array([[1.0, 0.0], [1.0, 0.0]]), mx.nd.array([1, -1])).asnumpy()
    np.testing.assert_allclose(ce, [0.0, 1.0], atol=1e-6)

    # contrib blocks
    cc = contrib.nn.HybridConcurrent(axis=1)
    cc.add(nn.Dense(3)); cc.add(contrib.nn.Identity()); cc.initialize()
    assert cc(mx.nd.array(torch.randn(2, 4))).shape == (2, 7)
    assert contrib.nn.PixelShuffle2D(2)(mx.nd.array(torch.randn(1, 8, 3, 3))).shape == (1, 2, 6, 6)
    se = contrib.nn.SparseEmbedding(10, 4); se.initialize()
    assert se(mx.nd.array([1, 3, 3])).shape == (3, 4)
    assert list(contrib.data.IntervalSampler(7, 3)) == [0, 3, 6, 1, 4, 2, 5] and list(contrib.data.IntervalSampler(7, 3, rollover=False)) == [0, 3, 6]

    # recurrent cells
    seq = mx.nd.array(torch.randn(2, 5, 4))
    bi = rnn.BidirectionalCell(rnn.LSTMCell(6), rnn.GRUCell(3)); bi.initialize()
    out, st = bi.unroll(5, seq)
    assert out.shape == (2, 5, 9) and len(st) == 3
    res = rnn.ResidualCell(rnn.RNNCell(4)); res.initialize()
    o, _ = res.unroll(5, seq); assert o.shape == (2, 5, 4)
    zc = rnn.ZoneoutCell(rnn.RNNCell(4), 0.5, 0.5); zc.initialize()
    with mx.autograd.record():
        o, _ = zc.unroll(5, seq)
    assert o.shape == (2, 5, 4)
    vd = contrib.rnn.VariationalDropoutCell(rnn.LSTMCell(4), drop_inputs=0.5, drop_outputs=0.5); vd.initialize()
    with mx.autograd.record():
        o, _ = vd.unroll(5, seq)
    zero_cols = (o.asnumpy() == 0).all(axis=1)                       # the output mask is shared by every time step
    assert zero_cols.any() and o.shape == (2, 5, 4)
    lp = contrib.rnn.LSTMPCell(8, 3); lp.initialize()
    o, st = lp.unroll(5, seq)
    assert o.shape == (2, 5, 3) and st[0].shape == (2, 3) and st[1].shape == (2, 8)
    for cls, ns in ((contrib.rnn.Conv2DRNNCell, 1), (contrib.rnn.Conv2DLSTMCell, 2), (contrib.rnn.Conv2DGRUCell, 1)):
        cell = cls((3, 8, 8), 5, 3, 3, i2h_pad=1); cell.initialize()
        s0 = cell.begin_state(2)
        assert len(s0) == ns and s0[0].shape == (2, 5, 8, 8)
        with mx.autograd.record():
            o, s1 = cell(mx.nd.array(torch.randn(2, 3, 8, 8)), s0)
            o2, _ = cell(mx.nd.array(torch.randn(2, 3, 8, 8)), s1)
            L = o2.sum()
        L.backward()
        assert o2.shape == (2, 5, 8, 8) and float(cell.h2h_weight.grad().asnumpy().__abs__().sum()) > 0


def test_model_zoo_covers_the_reference_table_with_the_published_sizes():
    from geomx_b200.gluon.model_zoo import vision
    names = set(vision._models)
    want = {"resnet%d_v%d" % (n, v) for n in (18, 34, 50, 101, 152) for v in (1, 2)} | {"vgg%d%s" % (n, s) for n in (11, 13, 16, 19) for s in ("", "_bn")} | \
        {"alexnet", "densenet121", "densenet161", "densenet169", "densenet201", "squeezenet1.0", "squeezenet1.1", "inceptionv3", "mobilenet1.0",
         "mobilenet0.75", "mobilenet0.5", "mobilenet0.25", "mobilenetv2_1.0", "mobilenetv2_0.75", "mobilenetv2_0.5", "mobilenetv2_0.25"}
    assert want <= names, sorted(want - names)

    def count(net):
        return sum(int(np.prod(p.shape)) for p in net.collect_params().values() if p.grad_req != "null")
    # learnable-parameter counts of the ImageNet configurations (the numbers every framework reports for these architectures)
    published = {"resnet50_v1": 25557032, "resnet18_v1": 11689512, "densenet121": 7978856, "squeezenet1.0": 1248424, "squeezenet1.1": 1235496,
                 "inceptionv3": 23834568, "mobilenet1.0": 4231976, "alexnet": 61100840}
    for name, n in published.items():
        net = vision.get_model(name); net.initialize()
        hw = 299 if name == "inceptionv3" else 224
        y = net(mx.nd.random.uniform(shape=(1, 3, hw, hw)))
        assert y.shape == (1, 1000) and count(net) == n, (name, count(net))
    for name in ("resnet50_v2", "mobilenetv2_0.5", "resnet18_v2"):          # pre-activation / inverted-residual families: forward + backward
        net = vision.get_model(name, classes=7); net.initialize()
        x = mx.nd.random.uniform(shape=(2, 3, 64, 64))
        with mx.autograd.record():
            loss = net(x).sum()
        loss.backward()
        conv_w = next(p for n, p in net.collect_params().items() if n.endswith("weight") and p.grad_req != "null")
        assert net(x).shape == (2, 7) and float(conv_w.grad().asnumpy().__abs__().sum()) > 0
    assert vision.get_resnet(2, 34, classes=3, thumbnail=True) is not None and vision.get_vgg(19, batch_norm=True) is not None
