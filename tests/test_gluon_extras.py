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
