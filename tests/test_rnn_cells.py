"""``mx.rnn`` symbolic cells (python/mxnet/rnn/rnn_cell.py): numerics against the imperative gluon cells, fused <-> unfused equivalence and
weight (un)packing, stacking / modifier cells, per-variable initializers, checkpoints, training through Module."""
import numpy as np
import pytest

import geomx_b200 as mx

T, N, C, H = 4, 3, 5, 6


def _weights(rs, prefix, G, cin=C):
    return {prefix + "i2h_weight": rs.randn(G * H, cin).astype(np.float32) * 0.3, prefix + "i2h_bias": rs.randn(G * H).astype(np.float32) * 0.1,
            prefix + "h2h_weight": rs.randn(G * H, H).astype(np.float32) * 0.3, prefix + "h2h_bias": rs.randn(G * H).astype(np.float32) * 0.1}


def _run(symbol, feed):
    return symbol.bind(mx.cpu(), {k: mx.nd.array(feed[k]) if isinstance(feed[k], np.ndarray) else feed[k] for k in symbol.list_arguments()}).forward()


@pytest.mark.parametrize("name,G", [("lstm", 4), ("gru", 3), ("rnn", 1)])
def test_symbolic_cells_match_gluon_cells(name, G):
    rs = np.random.RandomState(0)
    x = rs.randn(N, T, C).astype(np.float32)
    cell = {"lstm": mx.rnn.LSTMCell, "gru": mx.rnn.GRUCell, "rnn": mx.rnn.RNNCell}[name](H, prefix=name + "_")
    outs, states = cell.unroll(T, mx.sym.Variable("data"), layout="NTC", merge_outputs=True)
    W = _weights(rs, name + "_", G)
    y = _run(outs, dict(W, data=x))[0].asnumpy()
    g = {"lstm": mx.gluon.rnn.LSTMCell, "gru": mx.gluon.rnn.GRUCell, "rnn": mx.gluon.rnn.RNNCell}[name](H, input_size=C, prefix="g_")
    g.initialize()
    for k, v in W.items():
        getattr(g, k.split("_", 1)[1]).set_data(mx.nd.array(v))
    ref, ref_states = g.unroll(T, mx.nd.array(x), layout="NTC", merge_outputs=True)
    assert y.shape == (N, T, H) and np.allclose(y, ref.asnumpy(), atol=1e-6)
    last = _run(mx.sym.Group(states), dict(W, data=x))
    assert len(last) == len(ref_states) and all(np.allclose(a.asnumpy(), b.asnumpy(), atol=1e-6) for a, b in zip(last, ref_states))
    # step list in, step list out, TNC layout
    steps, _ = cell.unroll(T, [mx.sym.Variable("t%d" % i) for i in range(T)], layout="TNC", merge_outputs=False)
    assert len(steps) == T
    y2 = _run(steps[-1], dict(W, **{"t%d" % i: x[:, i] for i in range(T)}))[0].asnumpy()
    assert np.allclose(y2, y[:, -1], atol=1e-6)
    # per-gate weights <-> packed weights
    nd_w = {k: mx.nd.array(v) for k, v in W.items()}
    un = cell.unpack_weights(nd_w)
    assert len(un) == 4 * G and all(np.array_equal(cell.pack_weights(un)[k].asnumpy(), W[k]) for k in W)


@pytest.mark.parametrize("mode,bi,L", [("lstm", True, 2), ("gru", False, 2), ("rnn_tanh", True, 1), ("rnn_relu", False, 1)])
def test_fused_cell_equals_its_unfused_stack(mode, bi, L):
    rs = np.random.RandomState(1)
    x = rs.randn(N, T, C).astype(np.float32)
    fused = mx.rnn.FusedRNNCell(H, num_layers=L, mode=mode, bidirectional=bi, prefix="f_", get_next_state=True)
    fo, fs = fused.unroll(T, mx.sym.Variable("data"), layout="NTC", merge_outputs=True)
    n = sum(int(np.prod(s)) for _, s in fused._layout(C))
    flat = mx.nd.array(rs.randn(n).astype(np.float32) * 0.2)
    yf = _run(fo, {"data": x, "f_parameters": flat})[0].asnumpy()
    assert yf.shape == (N, T, H * (2 if bi else 1)) and len(fs) == (2 if mode == "lstm" else 1)
    stack = fused.unfuse()
    so, ss = stack.unroll(T, mx.sym.Variable("data"), layout="NTC", merge_outputs=True)
    un = fused.unpack_weights({"f_parameters": flat})
    assert "f_parameters" not in un and ("f_l0_i2h_i_weight" in un) == (mode == "lstm")
    ys = _run(so, dict(stack.pack_weights(dict(un)), data=x))[0].asnumpy()
    assert np.allclose(yf, ys, atol=1e-6)
    assert np.array_equal(fused.pack_weights(un)["f_parameters"].asnumpy(), flat.asnumpy())
    with pytest.raises(mx.MXNetError):
        fused(mx.sym.Variable("x"), [])


def test_stacked_modifier_and_bidirectional_cells():
    rs = np.random.RandomState(2)
    x = rs.randn(N, T, H).astype(np.float32)                    # input size = hidden size so that the residual connection fits
    base = mx.rnn.GRUCell(H, prefix="g0_")
    res = mx.rnn.ResidualCell(base)
    with pytest.raises(mx.MXNetError):
        base.begin_state()                                      # a wrapped cell may only be used through its wrapper
    W0 = _weights(rs, "g0_", 3, cin=H)
    yo, _ = res.unroll(T, mx.sym.Variable("data"), merge_outputs=True)
    plain = mx.rnn.GRUCell(H, prefix="g0_")
    po, _ = plain.unroll(T, mx.sym.Variable("data"), merge_outputs=True)
    assert np.allclose(_run(yo, dict(W0, data=x))[0].asnumpy(), _run(po, dict(W0, data=x))[0].asnumpy() + x, atol=1e-6)
    # zoneout with p = 0 is the identity; dropout with p = 0 too; a stack steps through all cells
    stack = mx.rnn.SequentialRNNCell()
    stack.add(mx.rnn.ZoneoutCell(mx.rnn.LSTMCell(H, prefix="l0_"), 0.0, 0.0))
    stack.add(mx.rnn.DropoutCell(0.0))
    stack.add(mx.rnn.LSTMCell(H, prefix="l1_"))
    assert len(stack.state_info) == 4
    so, ss = stack.unroll(T, mx.sym.Variable("data"), merge_outputs=True)
    W = dict(_weights(rs, "l0_", 4, cin=H), **_weights(rs, "l1_", 4, cin=H))
    y = _run(so, dict(W, data=x))[0].asnumpy()
    a, _ = mx.rnn.LSTMCell(H, prefix="l0_").unroll(T, mx.sym.Variable("data"), merge_outputs=True)
    b, _ = mx.rnn.LSTMCell(H, prefix="l1_").unroll(T, a, merge_outputs=True)
    assert len(ss) == 4 and np.allclose(y, _run(b, dict(W, data=x))[0].asnumpy(), atol=1e-6)
    # bidirectional: forward half equals the forward cell, backward half the cell run over the reversed sequence
    bi = mx.rnn.BidirectionalCell(mx.rnn.GRUCell(H, prefix="fw_"), mx.rnn.GRUCell(H, prefix="bw_"))
    bo, bs = bi.unroll(T, mx.sym.Variable("data"), merge_outputs=True)
    Wb = dict(_weights(rs, "fw_", 3, cin=H), **_weights(rs, "bw_", 3, cin=H))
    yb = _run(bo, dict(Wb, data=x))[0].asnumpy()
    fw, _ = mx.rnn.GRUCell(H, prefix="fw_").unroll(T, mx.sym.Variable("data"), merge_outputs=True)
    bw, _ = mx.rnn.GRUCell(H, prefix="bw_").unroll(T, mx.sym.Variable("data"), merge_outputs=True)
    assert yb.shape == (N, T, 2 * H) and len(bs) == 2
    assert np.allclose(yb[..., :H], _run(fw, dict(Wb, data=x))[0].asnumpy(), atol=1e-6)
    assert np.allclose(yb[..., H:], _run(bw, dict(Wb, data=x[:, ::-1].copy()))[0].asnumpy()[:, ::-1], atol=1e-6)
    with pytest.raises(mx.MXNetError):
        bi(mx.sym.Variable("x"), [])


def test_begin_state_variables_params_sharing_and_lstm_bias_init():
    cell = mx.rnn.LSTMCell(H, prefix="enc_", forget_bias=2.5)
    states = cell.begin_state(func=mx.sym.Variable)
    assert [s.name for s in states] == ["enc_begin_state_0", "enc_begin_state_1"]
    out, _ = cell(mx.sym.Variable("x"), states)
    assert set(out.list_arguments()) == {"x", "enc_begin_state_0", "enc_begin_state_1", "enc_i2h_weight", "enc_i2h_bias", "enc_h2h_weight", "enc_h2h_bias"}
    assert [tuple(s) for s in cell.state_shape] == [(0, H), (0, H)] and cell.begin_state(batch_size=7)[0].bind(mx.cpu(), {}).forward()[0].shape == (7, H)
    shared = mx.rnn.LSTMCell(H, prefix="dec_", params=cell.params)          # a second cell on the first one's parameters
    o2, _ = shared(mx.sym.Variable("x"), shared.begin_state(func=mx.sym.Variable))
    assert "enc_i2h_weight" in o2.list_arguments() and "dec_i2h_weight" not in o2.list_arguments()
    # the i2h bias variable names its own initializer: Module.init_params honours it over the global one
    outs, _ = mx.rnn.LSTMCell(H, prefix="m_", forget_bias=2.5).unroll(T, mx.sym.Variable("data"), merge_outputs=True)
    net = mx.sym.LinearRegressionOutput(mx.sym.sum(outs, axis=(1, 2)), mx.sym.Variable("softmax_label"), name="out")
    assert "lstmbias" in net.attr_dict()["m_i2h_bias"]["__init__"]
    mod = mx.mod.Module(net, data_names=("data",), label_names=("softmax_label",))
    mod.bind(data_shapes=[("data", (N, T, C))], label_shapes=[("softmax_label", (N,))])
    mod.init_params(mx.init.Uniform(0.1))
    b = mod.get_params()[0]["m_i2h_bias"].asnumpy()
    assert np.array_equal(b[H:2 * H], np.full(H, 2.5, dtype=np.float32)) and not b[:H].any() and not b[2 * H:].any()
    assert np.abs(mod.get_params()[0]["m_h2h_bias"].asnumpy()).max() <= 0.1


def test_rnn_checkpoint_and_training_through_module(tmp_path):
    rs = np.random.RandomState(3)
    fused = mx.rnn.FusedRNNCell(H, num_layers=1, mode="gru", prefix="r_")
    outs, _ = fused.unroll(T, mx.sym.Variable("data"), merge_outputs=True)
    pred = mx.sym.FullyConnected(mx.sym.Flatten(outs), num_hidden=1, name="fc")
    net = mx.sym.LinearRegressionOutput(pred, mx.sym.Variable("softmax_label"), name="out")
    X = rs.randn(32, T, C).astype(np.float32)
    y = X.sum((1, 2), keepdims=False).reshape(32, 1).astype(np.float32) * 0.1
    it = mx.io.NDArrayIter(X, y, batch_size=8, label_name="softmax_label")
    mod = mx.mod.Module(net, data_names=("data",), label_names=("softmax_label",))
    prefix = str(tmp_path / "rnn")
    mod.fit(it, num_epoch=12, optimizer="adam", optimizer_params={"learning_rate": 0.02}, initializer=mx.init.Xavier(), eval_metric="mse",
            epoch_end_callback=mx.rnn.do_rnn_checkpoint(fused, prefix, period=12))
    it.reset()
    mse = dict(mod.score(it, "mse"))["mse"]
    assert mse < 0.5 * float((y ** 2).mean()), mse
    # the file holds per-gate weights; loading for the fused cell packs them back, loading for the unfused stack keeps layer matrices
    import os
    assert os.path.exists(prefix + "-0012.params")
    raw = mx.nd.load(prefix + "-0012.params")
    assert "arg:r_l0_i2h_r_weight" in raw and "arg:r_parameters" not in raw
    _, arg, _ = mx.rnn.load_rnn_checkpoint(fused, prefix, 12)
    assert np.allclose(arg["r_parameters"].asnumpy(), mod.get_params()[0]["r_parameters"].asnumpy())
    _, arg2, _ = mx.rnn.load_rnn_checkpoint(fused.unfuse(), prefix, 12)
    assert arg2["r_l0_i2h_weight"].shape == (3 * H, C)
