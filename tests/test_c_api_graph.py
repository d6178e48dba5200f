"""Symbol / Executor / imperative / autograd / RecordIO / DataIter groups of the flat C ABI (csrc/runtime/c_api_graph.cc, c_api_io.cc), driven
through ctypes only and checked against PyTorch fp32 on the CPU and against the Python front end (JSON, RecordIO interchange)."""
import ctypes
import struct

import numpy as np
import pytest
import torch
import torch.nn.functional as TF

import _capi as C
from _capi import ck, lib, u32, vp, cp


def _cnn():
    """The reference's demo CNN (examples/cnn.py:56-66) composed through GXSymbol*."""
    x = C.var("data")
    c0 = C.op("Convolution", "conv0", [x], kernel="(5, 5)", num_filter=16)
    a0 = C.op("Activation", "relu0", [c0], act_type="relu")
    p0 = C.op("Pooling", "pool0", [a0], kernel="(2, 2)", stride="(2, 2)", pool_type="max")
    c1 = C.op("Convolution", "conv1", [p0], kernel="(5, 5)", num_filter=32)
    a1 = C.op("Activation", "relu1", [c1], act_type="relu")
    p1 = C.op("Pooling", "pool1", [a1], kernel="(2, 2)", stride="(2, 2)", pool_type="max")
    f = C.op("Flatten", "flat", [p1])
    d0 = C.op("FullyConnected", "fc0", [f], num_hidden=256)
    r0 = C.op("Activation", "relu2", [d0], act_type="relu")
    d1 = C.op("FullyConnected", "fc1", [r0], num_hidden=128)
    r1 = C.op("Activation", "relu3", [d1], act_type="relu")
    d2 = C.op("FullyConnected", "fc2", [r1], num_hidden=10)
    return C.op("SoftmaxOutput", "softmax", kwinputs={"data": d2, "label": C.var("softmax_label")}, normalization="batch")


def test_symbol_compose_infer_json_roundtrip(tmp_path):
    net = _cnn()
    args = C.list_arguments(net)
    assert args == ["data", "conv0_weight", "conv0_bias", "conv1_weight", "conv1_bias", "fc0_weight", "fc0_bias", "fc1_weight", "fc1_bias", "fc2_weight", "fc2_bias",
                    "softmax_label"]
    assert C.list_outputs(net) == ["softmax_output"] and C.list_aux(net) == []
    a, o, x, complete = C.infer_shape(net, data=(32, 1, 28, 28))
    assert complete and o == [(32, 10)] and x == []
    assert dict(zip(args, a))["fc0_weight"] == (256, 512) and dict(zip(args, a))["conv1_weight"] == (32, 16, 5, 5) and dict(zip(args, a))["softmax_label"] == (32,)
    assert sum(int(np.prod(s)) for n, s in zip(args, a) if n not in ("data", "softmax_label")) == 178762     # the parameter count BASELINE.md quotes
    # nothing known: partial inference reports incomplete, the strict form fails with the name of the first undetermined node
    a2, _, _, complete2 = C.infer_shape(net, partial=True)
    assert not complete2 and a2[0] == ()
    with pytest.raises(RuntimeError, match="cannot be determined"):
        C.infer_shape(net)
    with pytest.raises(RuntimeError, match="expected"):
        C.infer_shape(net, data=(32, 1, 28, 28), fc0_weight=(256, 100))
    # JSON: native -> Python front end -> native, same graph
    js = C.sym_json(net)
    import geomx_b200 as mx
    psym = mx.sym.load_json(js)
    assert psym.list_arguments() == args
    _, pout, _ = psym.infer_shape(data=(32, 1, 28, 28))
    assert pout == [(32, 10)]
    back = C.sym_from_json(psym.tojson(nnvm=True))
    assert C.list_arguments(back) == args and C.infer_shape(back, data=(8, 1, 28, 28))[1] == [(8, 10)]
    own = C.sym_from_json(psym.tojson())                       # this framework's own dialect loads too
    assert C.list_arguments(own) == args
    fname = str(tmp_path / "net-symbol.json").encode()
    ck(lib().GXSymbolSaveToFile(net, fname))
    h = vp(); ck(lib().GXSymbolCreateFromFile(fname, ctypes.byref(h)))
    assert C.sym_json(h) == js
    # internals / outputs / children / attributes / copy / print
    internals = vp(); ck(lib().GXSymbolGetInternals(net, ctypes.byref(internals)))
    names = C.list_outputs(internals)
    assert "fc1_output" in names and "conv0_weight" in names and names[-1] == "softmax_output"
    fc1 = vp(); ck(lib().GXSymbolGetOutput(internals, names.index("fc1_output"), ctypes.byref(fc1)))
    assert C.list_outputs(fc1) == ["fc1_output"] and "fc2_weight" not in C.list_arguments(fc1)
    kids = vp(); ck(lib().GXSymbolGetChildren(fc1, ctypes.byref(kids)))
    assert C.list_outputs(kids) == ["relu2_output", "fc1_weight", "fc1_bias"]
    out, ok = cp(), ctypes.c_int()
    ck(lib().GXSymbolGetAttr(fc1, b"num_hidden", ctypes.byref(out), ctypes.byref(ok))); assert ok.value == 1 and out.value == b"128"
    w = C.var("w"); ck(lib().GXSymbolSetAttr(w, b"__lr_mult__", b"0.1"))
    ck(lib().GXSymbolGetAttr(w, b"lr_mult", ctypes.byref(out), ctypes.byref(ok))); assert ok.value == 1 and out.value == b"0.1"
    ck(lib().GXSymbolGetAttr(w, b"nope", ctypes.byref(out), ctypes.byref(ok))); assert ok.value == 0
    n = u32(); ck(lib().GXSymbolGetNumOutputs(internals, ctypes.byref(n))); assert n.value == len(names)
    ck(lib().GXSymbolGetName(net, ctypes.byref(out), ctypes.byref(ok))); assert out.value == b"softmax"
    cpy = vp(); ck(lib().GXSymbolCopy(net, ctypes.byref(cpy))); assert C.sym_json(cpy) == js
    ck(lib().GXSymbolPrint(net, ctypes.byref(out))); assert b"Op:Convolution, Name=conv0" in out.value
    grp = vp(); ck(lib().GXSymbolCreateGroup(2, C.handles([fc1, net]), ctypes.byref(grp))); assert C.list_outputs(grp) == ["fc1_output", "softmax_output"]
    # operator table
    ops = C.str_list(lib().GXListAllOpNames)
    assert {"Convolution", "FullyConnected", "BatchNorm", "SoftmaxOutput", "broadcast_mul", "dot"} <= set(ops)
    nc, creators = u32(), ctypes.POINTER(vp)()
    ck(lib().GXSymbolListAtomicSymbolCreators(ctypes.byref(nc), ctypes.byref(creators)))
    info = {}
    for i in range(nc.value):
        nm, desc, na, an, at, ad, kv, rt = cp(), cp(), u32(), ctypes.POINTER(cp)(), ctypes.POINTER(cp)(), ctypes.POINTER(cp)(), cp(), cp()
        ck(lib().GXSymbolGetAtomicSymbolInfo(vp(creators[i]), ctypes.byref(nm), ctypes.byref(desc), ctypes.byref(na), ctypes.byref(an), ctypes.byref(at), ctypes.byref(ad),
                                             ctypes.byref(kv), ctypes.byref(rt)))
        info[nm.value.decode()] = ([an[j].decode() for j in range(na.value)], kv.value.decode())
    assert info["Convolution"][0][:2] == ["kernel", "num_filter"] and info["Concat"][1] == "num_args"
    # errors are reported, not thrown across the ABI
    bad = vp()
    assert lib().GXSymbolCreateAtomicSymbolByName(b"NoSuchOp", 0, None, None, ctypes.byref(bad)) == -1 and "not registered" in C.err()
    assert lib().GXSymbolCreateFromJSON(b'{"nodes": [{"op": "null", "name": "x", "inputs": [[3, 0, 0]]}], "heads": [[0, 0, 0]]}', ctypes.byref(bad)) == -1
    for h in (net, back, own, internals, fc1, kids, cpy, grp, w):
        ck(lib().GXSymbolFree(h))


def _torch_cnn(params, X, y):
    p = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
    h = TF.max_pool2d(torch.relu(TF.conv2d(X, p["conv0_weight"], p["conv0_bias"])), 2)
    h = TF.max_pool2d(torch.relu(TF.conv2d(h, p["conv1_weight"], p["conv1_bias"])), 2).flatten(1)
    h = torch.relu(TF.linear(h, p["fc0_weight"], p["fc0_bias"]))
    h = torch.relu(TF.linear(h, p["fc1_weight"], p["fc1_bias"]))
    logits = TF.linear(h, p["fc2_weight"], p["fc2_bias"])
    TF.cross_entropy(logits, y.long(), reduction="mean").backward()
    return torch.softmax(logits, 1).detach().numpy(), {k: v.grad.numpy() for k, v in p.items()}


def test_executor_cnn_matches_torch_and_trains():
    net = _cnn()
    B = 16
    ex, args, grads, aux = C.simple_bind(net, {"data": (B, 1, 28, 28)}, no_grad=("data", "softmax_label"))
    assert set(grads) == set(args) - {"data", "softmax_label"} and aux == {}
    rng = np.random.RandomState(0)
    params = {}
    for k, h in args.items():
        if k in ("data", "softmax_label"):
            continue
        shp = C.nd_shape(h)
        params[k] = (rng.randn(*shp) * (0.1 if k.endswith("weight") else 0.01)).astype(np.float32)
        C.nd_set(h, params[k])
    X = rng.rand(B, 1, 28, 28).astype(np.float32); y = rng.randint(0, 10, B).astype(np.float32)
    C.nd_set(args["data"], X); C.nd_set(args["softmax_label"], y)
    prob = C.forward(ex, True)[0]
    C.backward(ex)
    tprob, tgrad = _torch_cnn(params, torch.tensor(X), torch.tensor(y))
    assert np.allclose(prob, tprob, atol=1e-5)
    for k in params:
        g = C.nd_get(grads[k])
        assert np.allclose(g, tgrad[k], rtol=1e-3, atol=1e-4 * max(1.0, np.abs(tgrad[k]).max())), (k, np.abs(g - tgrad[k]).max())
    # a few SGD steps through the C ABI only: the loss goes down
    def loss():
        p = C.forward(ex, True)[0]
        return float(-np.log(p[np.arange(B), y.astype(int)] + 1e-12).mean())
    l0 = loss()
    for _ in range(30):
        loss(); C.backward(ex)
        for k in params:
            params[k] = params[k] - 0.05 * C.nd_get(grads[k]); C.nd_set(args[k], params[k])
    assert loss() < 0.8 * l0
    out = cp(); ck(lib().GXExecutorPrint(ex, ctypes.byref(out))); assert b"Op:Convolution, Name=conv1" in out.value
    # grad_req add accumulates; explicit Bind with caller-owned arrays
    x = C.var("x"); s = C.op("_mul_scalar", "twice", [x], scalar=2.0)
    xa, ga = C.nd_create(np.ones((2, 3))), C.nd_create(np.full((2, 3), 5.0))
    e2 = vp()
    ck(lib().GXExecutorBind(s, 1, 0, 1, C.handles([xa]), C.handles([ga]), (u32 * 1)(3), 0, None, ctypes.byref(e2)))
    assert np.allclose(C.forward(e2, True)[0], 2.0)
    hg = C.nd_create(np.full((2, 3), 0.5)); C.backward(e2, [hg])
    assert np.allclose(C.nd_get(ga), 6.0)
    assert lib().GXExecutorBackward(e2, 1, C.handles([C.nd_create(np.ones(5))])) == -1 and "does not match" in C.err()
    ck(lib().GXExecutorFree(e2)); ck(lib().GXExecutorFree(ex))


def _t(a, grad=True):
    return torch.tensor(np.asarray(a, dtype=np.float32), requires_grad=grad)


def test_imperative_autograd_matches_torch():
    rng = np.random.RandomState(1)
    xn, wn, gn, bn = rng.randn(4, 3, 6, 6), rng.randn(5, 3, 3, 3) * 0.3, rng.rand(5) + 0.5, rng.randn(5) * 0.1
    emb_idx, emb_w = np.array([[0, 2], [3, 2]], dtype=np.float32), rng.randn(4, 6)
    x, w, g, b, ew = [C.nd_create(a) for a in (xn, wn, gn, bn, emb_w)]
    grads = [C.nd_create(np.zeros_like(a)) for a in (xn, wn, gn, bn, emb_w)]
    C.mark_variables([x, w, g, b, ew], grads)
    mm, mv = C.nd_create(np.zeros(5)), C.nd_create(np.ones(5))
    with C.record():
        c = C.invoke("Convolution", [x, w], kernel="(3, 3)", num_filter=5, pad="(1, 1)", stride="(2, 2)", no_bias=True)      # (4, 5, 3, 3)
        n = C.invoke("BatchNorm", [c, g, b, mm, mv], fix_gamma=False, eps=1e-5, momentum=0.9)
        a = C.invoke("LeakyReLU", [n], act_type="leaky", slope=0.1)
        p = C.invoke("Pooling", [a], kernel="(2, 2)", stride="(1, 1)", pad="(1, 1)", pool_type="avg")                          # (4, 5, 4, 4)
        t = C.invoke("transpose", [p], axes="(0, 2, 3, 1)")
        r = C.invoke("Reshape", [t], shape="(4, -1)")                                                                           # (4, 80)
        e = C.invoke("Embedding", [C.nd_create(emb_idx), ew], input_dim=4, output_dim=6)                                         # (2, 2, 6)
        e2 = C.invoke("Reshape", [e], shape="(4, 6)")
        cat = C.invoke("Concat", [r, e2], dim=1)                                                                                # (4, 86)
        sm = C.invoke("log_softmax", [cat], axis=-1)
        m = C.invoke("mean", [sm], axis="(1,)", keepdims=True)                                                                  # (4, 1)
        bm = C.invoke("broadcast_mul", [cat, m])
        d = C.invoke("dot", [bm, cat], transpose_b=True)                                                                        # (4, 4)
        out = C.invoke("sum", [C.invoke("tanh", [d])])
    sym = vp(); ck(lib().GXAutogradGetSymbol(out, ctypes.byref(sym)))
    assert [a_ for a_ in C.list_arguments(sym) if a_.startswith("var")] == ["var0", "var1", "var2", "var3", "var4"] and "Op:dot" in _print(sym)
    C.ag_backward([out])
    with pytest.raises(RuntimeError, match="already freed"):
        C.ag_backward([out])
    # the same computation in torch
    tx, tw, tg, tb, tew = _t(xn), _t(wn), _t(gn), _t(bn), _t(emb_w)
    tc = TF.conv2d(tx, tw, None, stride=2, padding=1)
    tn = TF.batch_norm(tc, torch.zeros(5), torch.ones(5), tg, tb, training=True, momentum=0.1, eps=1e-5)
    ta = TF.leaky_relu(tn, 0.1)
    tp = TF.avg_pool2d(ta, 2, 1, 1, count_include_pad=True)
    tr = tp.permute(0, 2, 3, 1).reshape(4, -1)
    te = tew[torch.tensor(emb_idx).long()].reshape(4, 6)
    tcat = torch.cat([tr, te], 1)
    tm = torch.log_softmax(tcat, -1).mean(1, keepdim=True)
    tout = torch.tanh((tcat * tm) @ tcat.t()).sum()
    tout.backward()
    assert np.allclose(C.nd_get(out), tout.item(), rtol=1e-4)
    for h, tt, name in zip(grads, (tx, tw, tg, tb, tew), "x w gamma beta emb".split()):
        got, want = C.nd_get(h), tt.grad.numpy()
        assert np.allclose(got, want, rtol=2e-3, atol=2e-5 * max(1.0, np.abs(want).max())), (name, np.abs(got - want).max())
    # running statistics were updated in place (momentum 0.9 towards the batch statistics)
    assert np.allclose(C.nd_get(mm), 0.1 * tc.detach().mean((0, 2, 3)).numpy(), atol=1e-5)
    # not recording: no history; dropout is the identity outside training mode and a mask inside
    y = C.invoke("relu", [x])
    with pytest.raises(RuntimeError, match="not computed while recording"):
        C.ag_backward([y])
    ones = C.nd_create(np.ones((64, 64)))
    assert np.array_equal(C.nd_get(C.invoke("Dropout", [ones], p=0.5)), np.ones((64, 64)))
    with C.record(train=True):
        dm = C.nd_get(C.invoke("Dropout", [ones], p=0.5))
    assert set(np.unique(dm)) == {0.0, 2.0} and 0.35 < (dm == 0).mean() < 0.65
    # wrong input count / dtype are errors
    with pytest.raises(RuntimeError, match="inputs given"):
        C.invoke("dot", [x])
    gh = vp(); ck(lib().GXNDArrayGetGrad(x, ctypes.byref(gh))); assert gh.value == grads[0].value
    det = vp(); ck(lib().GXNDArrayDetach(out, ctypes.byref(det))); assert np.allclose(C.nd_get(det), C.nd_get(out))


def _print(sym):
    out = cp(); ck(lib().GXSymbolPrint(sym, ctypes.byref(out)))
    return out.value.decode()


@pytest.mark.parametrize("case", ["sigmoid_logistic", "maxpool_full", "softmax_ignore", "grouped_conv", "broadcast_div_max", "scalars"])
def test_operator_gradients(case):
    rng = np.random.RandomState(2)
    if case == "sigmoid_logistic":
        xn, yn = rng.randn(6, 3), rng.rand(6, 3)
        net = C.op("LogisticRegressionOutput", "lro", kwinputs={"data": C.op("FullyConnected", "fc", [C.var("data")], num_hidden=3, no_bias=True), "label": C.var("label")},
                   grad_scale=2.0)
        ex, args, grads, _ = C.simple_bind(net, {"data": (6, 3)}, no_grad=("label",))
        wn = rng.randn(3, 3)
        C.nd_set(args["data"], xn); C.nd_set(args["fc_weight"], wn); C.nd_set(args["label"], yn)
        C.forward(ex, True); C.backward(ex)
        tx, tw = _t(xn), _t(wn)
        z = tx @ tw.t()                      # the op's gradient (p - y) * grad_scale / num_output is taken w.r.t. its INPUT (regression_output-inl.h:200)
        z.backward((torch.sigmoid(z.detach()) - torch.tensor(yn, dtype=torch.float32)) * 2.0 / 3)
        assert np.allclose(C.nd_get(grads["data"]), tx.grad.numpy(), atol=1e-5) and np.allclose(C.nd_get(grads["fc_weight"]), tw.grad.numpy(), atol=1e-5)
    elif case == "maxpool_full":
        xn = rng.randn(2, 3, 7, 7)
        net = C.op("MakeLoss", "l", [C.op("Pooling", "p", [C.var("data")], kernel="(3, 3)", stride="(2, 2)", pad="(1, 1)", pool_type="max", pooling_convention="full")])
        ex, args, grads, _ = C.simple_bind(net, {"data": (2, 3, 7, 7)})
        C.nd_set(args["data"], xn)
        out = C.forward(ex, True)[0]; C.backward(ex)
        tx = _t(xn)
        tp = TF.max_pool2d(tx, 3, 2, 1, ceil_mode=True)
        tp.sum().backward()
        assert out.shape == tuple(tp.shape) and np.allclose(out, tp.detach().numpy()) and np.allclose(C.nd_get(grads["data"]), tx.grad.numpy())
    elif case == "softmax_ignore":
        xn = rng.randn(2, 4, 5); yn = rng.randint(0, 4, (2, 5)).astype(np.float32); yn[0, 1] = -1; yn[1, 3] = -1
        net = C.op("SoftmaxOutput", "sm", kwinputs={"data": C.var("data"), "label": C.var("label")}, multi_output=True, use_ignore=True, ignore_label=-1,
                   normalization="valid")
        ex, args, grads, _ = C.simple_bind(net, {"data": (2, 4, 5)}, no_grad=("label",))
        assert C.nd_shape(args["label"]) == (2, 5)
        C.nd_set(args["data"], xn); C.nd_set(args["label"], yn)
        C.forward(ex, True); C.backward(ex)
        tx = _t(xn)
        TF.cross_entropy(tx, torch.tensor(yn).long(), ignore_index=-1, reduction="mean").backward()
        assert np.allclose(C.nd_get(grads["data"]), tx.grad.numpy(), atol=1e-6)
    elif case == "grouped_conv":
        xn, wn, bn = rng.randn(2, 4, 6, 5), rng.randn(6, 2, 3, 2), rng.randn(6)
        net = C.op("MakeLoss", "l", [C.op("square", "sq", [C.op("Convolution", "c", [C.var("data")], kernel="(3, 2)", num_filter=6, num_group=2, dilate="(2, 1)", pad="(2, 0)")])])
        ex, args, grads, _ = C.simple_bind(net, {"data": (2, 4, 6, 5)})
        for k, v in (("data", xn), ("c_weight", wn), ("c_bias", bn)):
            C.nd_set(args[k], v)
        out = C.forward(ex, True)[0]; C.backward(ex)
        tx, tw, tb = _t(xn), _t(wn), _t(bn)
        to = TF.conv2d(tx, tw, tb, dilation=(2, 1), padding=(2, 0), groups=2) ** 2
        to.sum().backward()
        assert np.allclose(out, to.detach().numpy(), rtol=1e-4, atol=1e-5)
        for k, t in (("data", tx), ("c_weight", tw), ("c_bias", tb)):
            assert np.allclose(C.nd_get(grads[k]), t.grad.numpy(), rtol=1e-3, atol=1e-4), k
    elif case == "broadcast_div_max":
        an, bn = rng.rand(3, 1, 4) + 0.5, rng.rand(2, 1) + 0.5
        a, b = C.var("a"), C.var("b")
        net = C.op("MakeLoss", "l", [C.op("broadcast_maximum", "mx", [C.op("broadcast_div", "dv", [a, b]), C.op("broadcast_sub", "sb", [a, b])])])
        ex, args, grads, _ = C.simple_bind(net, {"a": (3, 1, 4), "b": (2, 1)})
        C.nd_set(args["a"], an); C.nd_set(args["b"], bn)
        out = C.forward(ex, True)[0]; C.backward(ex)
        ta, tb = _t(an), _t(bn)
        to = torch.maximum(ta / tb, ta - tb); to.sum().backward()
        assert out.shape == (3, 2, 4) and np.allclose(out, to.detach().numpy(), atol=1e-6)
        assert np.allclose(C.nd_get(grads["a"]), ta.grad.numpy(), atol=1e-5) and np.allclose(C.nd_get(grads["b"]), tb.grad.numpy(), atol=1e-5)
    else:
        xn = rng.rand(3, 4) + 0.5
        x = C.var("x")
        h = C.op("_rdiv_scalar", "rd", [C.op("_power_scalar", "pw", [C.op("_rminus_scalar", "rm", [x], scalar=3.0)], scalar=1.5)], scalar=2.0)
        h = C.op("elemwise_add", "ad", [C.op("softsign", "ss", [C.op("sqrt", "sq", [h])]), C.op("BlockGrad", "bg", [C.op("exp", "ex", [x])])])
        net = C.op("MakeLoss", "l", [C.op("add_n", "an", [h, C.op("clip", "cl", [x], a_min=0.8, a_max=1.2), C.op("log", "lg", [x])])], grad_scale=0.5)
        ex, args, grads, _ = C.simple_bind(net, {"x": (3, 4)})
        C.nd_set(args["x"], xn)
        out = C.forward(ex, True)[0]; C.backward(ex)
        tx = _t(xn)
        th = torch.sqrt(2.0 / (3.0 - tx) ** 1.5)
        to = th / (1 + th.abs()) + torch.exp(tx).detach() + tx.clamp(0.8, 1.2) + torch.log(tx)
        (0.5 * to).sum().backward()
        assert np.allclose(out, to.detach().numpy(), rtol=1e-5) and np.allclose(C.nd_get(grads["x"]), tx.grad.numpy(), rtol=1e-4, atol=1e-6)
    ck(lib().GXExecutorFree(ex))


def test_batchnorm_inference_and_aux_states():
    x = C.var("data")
    net = C.op("BatchNorm", "bn", [x], fix_gamma=True, eps=1e-3)
    assert C.list_arguments(net) == ["data", "bn_gamma", "bn_beta"] and C.list_aux(net) == ["bn_moving_mean", "bn_moving_var"]
    a, o, aux, ok = C.infer_shape(net, data=(4, 3, 2, 2))
    assert ok and a == [(4, 3, 2, 2), (3,), (3,)] and aux == [(3,), (3,)]
    ex, args, grads, auxs = C.simple_bind(net, {"data": (4, 3, 2, 2)})
    rng = np.random.RandomState(3)
    xn = rng.randn(4, 3, 2, 2).astype(np.float32)
    C.nd_set(args["data"], xn); C.nd_set(args["bn_gamma"], np.full(3, 7.0)); C.nd_set(args["bn_beta"], np.array([1.0, 2.0, 3.0]))
    C.nd_set(auxs["bn_moving_mean"], np.array([0.5, -0.5, 0.0])); C.nd_set(auxs["bn_moving_var"], np.array([4.0, 1.0, 0.25]))
    out = C.forward(ex, False)[0]                      # inference: running statistics, gamma fixed to 1
    want = (xn - np.array([0.5, -0.5, 0.0]).reshape(1, 3, 1, 1)) / np.sqrt(np.array([4.0, 1.0, 0.25]).reshape(1, 3, 1, 1) + 1e-3) + np.array([1.0, 2.0, 3.0]).reshape(1, 3, 1, 1)
    assert np.allclose(out, want, atol=1e-5)
    assert np.allclose(C.nd_get(auxs["bn_moving_mean"]), [0.5, -0.5, 0.0])
    out = C.forward(ex, True)[0]                       # training: batch statistics, running statistics move
    assert np.allclose(out.mean((0, 2, 3)), [1.0, 2.0, 3.0], atol=1e-4)
    assert np.allclose(C.nd_get(auxs["bn_moving_mean"]), 0.9 * np.array([0.5, -0.5, 0.0]) + 0.1 * xn.mean((0, 2, 3)), atol=1e-5)
    C.backward(ex, [C.nd_create(np.ones((4, 3, 2, 2)))])
    assert np.allclose(C.nd_get(grads["bn_gamma"]), 0.0) and np.allclose(C.nd_get(grads["bn_beta"]), 16.0)
    ck(lib().GXExecutorFree(ex))


def test_recordio_c_api_interchange_with_python(tmp_path):
    import geomx_b200 as mx
    magic = struct.pack("<I", 0xced7230a)
    payloads = [b"hello", b"", b"x" * 1001, b"abcd" + magic + b"tail", magic + magic + b"zz", bytes(range(256)) * 3]
    path = str(tmp_path / "c.rec")
    w = vp(); ck(lib().GXRecordIOWriterCreate(path.encode(), ctypes.byref(w)))
    offs = []
    for p in payloads:
        pos = ctypes.c_size_t(); ck(lib().GXRecordIOWriterTell(w, ctypes.byref(pos))); offs.append(pos.value)
        ck(lib().GXRecordIOWriterWriteRecord(w, p, ctypes.c_size_t(len(p))))
    ck(lib().GXRecordIOWriterFree(w))
    r = mx.recordio.MXRecordIO(path, "r")              # the Python front end reads what the C API wrote
    assert [r.read() for _ in payloads] == payloads and r.read() is None
    r.close()
    path2 = str(tmp_path / "py.rec")
    pw = mx.recordio.MXRecordIO(path2, "w")
    for p in payloads:
        pw.write(p)
    pw.close()
    assert open(path, "rb").read() == open(path2, "rb").read()          # byte-identical files
    rd = vp(); ck(lib().GXRecordIOReaderCreate(path2.encode(), ctypes.byref(rd)))
    got = []
    while True:
        buf, size = ctypes.POINTER(ctypes.c_char)(), ctypes.c_size_t()
        ck(lib().GXRecordIOReaderReadRecord(rd, ctypes.byref(buf), ctypes.byref(size)))
        if not buf:
            break
        got.append(ctypes.string_at(buf, size.value))
    assert got == payloads
    ck(lib().GXRecordIOReaderSeek(rd, ctypes.c_size_t(offs[3])))
    buf, size = ctypes.POINTER(ctypes.c_char)(), ctypes.c_size_t()
    ck(lib().GXRecordIOReaderReadRecord(rd, ctypes.byref(buf), ctypes.byref(size)))
    assert ctypes.string_at(buf, size.value) == payloads[3]
    ck(lib().GXRecordIOReaderFree(rd))
    # a damaged file is an error, not a crash
    bad = str(tmp_path / "bad.rec")
    open(bad, "wb").write(open(path, "rb").read()[:30] + b"\x00" * 7)
    rd = vp(); ck(lib().GXRecordIOReaderCreate(bad.encode(), ctypes.byref(rd)))
    rcs = []
    for _ in range(4):
        rcs.append(lib().GXRecordIOReaderReadRecord(rd, ctypes.byref(buf), ctypes.byref(size)))
    assert -1 in rcs
    ck(lib().GXRecordIOReaderFree(rd))


def test_data_iterators_c_api(tmp_path):
    rng = np.random.RandomState(4)
    img = rng.randint(0, 256, (50, 28, 28)).astype(np.uint8); lab = rng.randint(0, 10, 50).astype(np.uint8)
    ip, lp = str(tmp_path / "img.idx"), str(tmp_path / "lab.idx")
    open(ip, "wb").write(struct.pack(">IIII", 0x803, 50, 28, 28) + img.tobytes())
    open(lp, "wb").write(struct.pack(">II", 0x801, 50) + lab.tobytes())
    n, creators = u32(), ctypes.POINTER(vp)()
    ck(lib().GXListDataIters(ctypes.byref(n), ctypes.byref(creators)))
    byname = {}
    for i in range(n.value):
        nm, desc, na, an, at, ad = cp(), cp(), u32(), ctypes.POINTER(cp)(), ctypes.POINTER(cp)(), ctypes.POINTER(cp)()
        ck(lib().GXDataIterGetIterInfo(vp(creators[i]), ctypes.byref(nm), ctypes.byref(desc), ctypes.byref(na), ctypes.byref(an), ctypes.byref(at), ctypes.byref(ad)))
        byname[nm.value.decode()] = vp(creators[i])
    assert set(byname) == {"MNISTIter", "CSVIter"}

    def make(name, **kw):
        h = vp()
        rc = lib().GXDataIterCreateIter(byname[name], len(kw), C.strs(list(kw.keys())), C.strs([str(v) for v in kw.values()]), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(C.err())
        return h

    def batches(it):
        out = []
        while True:
            more = ctypes.c_int(); ck(lib().GXDataIterNext(it, ctypes.byref(more)))
            if not more.value:
                return out
            d, l, pad = vp(), vp(), ctypes.c_int()
            ck(lib().GXDataIterGetData(it, ctypes.byref(d))); ck(lib().GXDataIterGetLabel(it, ctypes.byref(l))); ck(lib().GXDataIterGetPadNum(it, ctypes.byref(pad)))
            idx, ni = ctypes.POINTER(ctypes.c_uint64)(), ctypes.c_uint64()
            ck(lib().GXDataIterGetIndex(it, ctypes.byref(idx), ctypes.byref(ni)))
            out.append((C.nd_get(d), C.nd_get(l), pad.value, [idx[i] for i in range(ni.value)]))
    it = make("MNISTIter", image=ip, label=lp, batch_size=16, shuffle=0)
    bs = batches(it)
    assert len(bs) == 3 and bs[0][0].shape == (16, 1, 28, 28) and bs[0][1].shape == (16,)            # 50 // 16 full batches, the partial one is dropped
    assert np.allclose(bs[1][0][:, 0], img[16:32] / 256.0) and np.array_equal(bs[1][1], lab[16:32]) and bs[1][3] == list(range(16, 32))
    ck(lib().GXDataIterBeforeFirst(it)); assert len(batches(it)) == 3
    ck(lib().GXDataIterFree(it))
    it = make("MNISTIter", image=ip, label=lp, batch_size=10, shuffle=1, seed=7, flat=1, num_parts=2, part_index=1)
    bs = batches(it)
    seen = sorted(i for b in bs for i in b[3])
    assert bs[0][0].shape == (10, 784) and len(bs) == 2 and len(set(seen)) == 20 and bs[0][3] != list(range(10))
    assert all(np.array_equal(b[1], lab[25:][b[3]]) for b in bs)
    ck(lib().GXDataIterFree(it))
    dn, ln = rng.randn(7, 6).astype(np.float32), rng.randint(0, 3, (7, 1)).astype(np.float32)
    dp, lpath = str(tmp_path / "d.csv"), str(tmp_path / "l.csv")
    np.savetxt(dp, dn, delimiter=",", fmt="%.8g"); np.savetxt(lpath, ln, delimiter=",", fmt="%g")
    it = make("CSVIter", data_csv=dp, data_shape="(2, 3)", label_csv=lpath, batch_size=3)
    bs = batches(it)
    assert [b[2] for b in bs] == [0, 0, 2] and bs[0][0].shape == (3, 2, 3) and np.allclose(bs[2][0][0].ravel(), dn[6], atol=1e-6)
    assert np.allclose(bs[2][0][1].ravel(), dn[0], atol=1e-6) and np.array_equal(bs[2][1], [ln[6, 0], ln[0, 0], ln[1, 0]])          # round_batch wraps to the start
    ck(lib().GXDataIterFree(it))
    with pytest.raises(RuntimeError, match="columns"):
        make("CSVIter", data_csv=dp, data_shape="(5,)", batch_size=2)
    with pytest.raises(RuntimeError, match="required"):
        make("MNISTIter", image=ip)
    v = ctypes.c_int(); ck(lib().GXGetVersion(ctypes.byref(v))); assert v.value == 10400


def test_python_free_c_library_and_c_example(tmp_path):
    """lib/libgeomx_capi.so exports every function the public header declares, depends on neither libpython nor torch, and the pure-C
    training example (examples/c_api/train_cnn.c) builds against it, trains the demo CNN and serves its checkpoint through GXPred*."""
    import os
    import re
    import shutil
    import subprocess
    so = os.path.join(C.ROOT, "geomx_b200", "lib", "libgeomx_capi.so")
    if not os.path.exists(so):
        from geomx_b200 import build
        build.build_capi()
    header = open(os.path.join(C.ROOT, "geomx_b200", "include", "geomx", "c_api.h")).read()
    declared = set(re.findall(r"\b(GX[A-Za-z0-9]+)\s*\(", header)) - {"GXEngineFn"}
    exported = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout.splitlines() if " T " in l}
    assert len(declared) > 130 and not (declared - exported), sorted(declared - exported)
    pyext = {l.split()[-1] for l in subprocess.run(["nm", "-D", "--defined-only", C.lib()._name], capture_output=True, text=True, check=True).stdout.splitlines() if " T " in l}
    assert not (declared - pyext), sorted(declared - pyext)          # the Python extension carries the same ABI
    needed = subprocess.run(["ldd", so], capture_output=True, text=True, check=True).stdout
    assert "python" not in needed and "torch" not in needed, needed
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "train_cnn")
    subprocess.run([cc, "-O2", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(C.ROOT, "geomx_b200", "include"), os.path.join(C.ROOT, "examples", "c_api", "train_cnn.c"),
                    "-L", os.path.dirname(so), "-lgeomx_capi", "-Wl,-rpath," + os.path.dirname(so), "-lm", "-o", exe], check=True)
    r = subprocess.run([exe, "40", str(tmp_path / "cnn")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "predictor agrees with the executor on 32/32" in r.stdout
    # the checkpoint the C program wrote is a regular checkpoint for the Python front end
    import geomx_b200 as mx
    sym, arg, aux = mx.model.load_checkpoint(str(tmp_path / "cnn"), 1)
    assert sym.list_outputs() == ["softmax_output"] and tuple(arg["fc0_weight"].shape) == (256, 512) and aux == {}


def test_ndarray_views_raw_bytes_profile_objects_and_knobs(tmp_path):
    a = C.nd_create(np.arange(24, dtype=np.float32).reshape(4, 3, 2))
    s, t, r = vp(), vp(), vp()
    ck(lib().GXNDArraySlice(a, 1, 3, ctypes.byref(s))); assert np.array_equal(C.nd_get(s), np.arange(24).reshape(4, 3, 2)[1:3])
    ck(lib().GXNDArrayAt(a, 2, ctypes.byref(t))); assert np.array_equal(C.nd_get(t), np.arange(24).reshape(4, 3, 2)[2])
    ck(lib().GXNDArrayReshape(a, 3, (ctypes.c_int * 3)(0, -1, 1), ctypes.byref(r))); assert C.nd_shape(r) == (4, 6, 1)
    assert lib().GXNDArraySlice(a, 3, 9, ctypes.byref(s)) == -1 and lib().GXNDArrayReshape(a, 2, (ctypes.c_int * 2)(5, -1), ctypes.byref(r)) == -1
    dt, di, st = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    ck(lib().GXNDArrayGetContext(a, ctypes.byref(dt), ctypes.byref(di))); ck(lib().GXNDArrayGetStorageType(a, ctypes.byref(st)))
    assert (dt.value, di.value, st.value) == (1, 0, 0)
    ck(lib().GXNDArrayWaitToRead(a)); ck(lib().GXNDArrayWaitToWrite(a)); ck(lib().GXNDArrayWaitAll())
    n, buf = ctypes.c_size_t(), ctypes.POINTER(ctypes.c_char)()
    ck(lib().GXNDArraySaveRawBytes(a, ctypes.byref(n), ctypes.byref(buf)))
    raw = ctypes.string_at(buf, n.value)
    b = vp(); ck(lib().GXNDArrayLoadFromRawBytes(raw, ctypes.c_size_t(len(raw)), ctypes.byref(b)))
    assert np.array_equal(C.nd_get(b), C.nd_get(a))
    assert lib().GXNDArrayLoadFromRawBytes(raw[:20], ctypes.c_size_t(20), ctypes.byref(b)) == -1
    none = vp(); ck(lib().GXNDArrayCreateNone(ctypes.byref(none))); ck(lib().GXNDArrayGetStorageType(none, ctypes.byref(st))); assert st.value == -1
    # profiler objects end up in the chrome trace
    prof = str(tmp_path / "objs.json").encode()
    ck(lib().GXSetProfilerConfig(1, (cp * 1)(b"filename"), (cp * 1)(prof))); ck(lib().GXSetProfilerState(1))
    dom, task, ev, ctr = vp(), vp(), vp(), vp()
    ck(lib().GXProfileCreateDomain(b"capi_domain", ctypes.byref(dom))); ck(lib().GXProfileCreateTask(dom, b"capi_task", ctypes.byref(task)))
    ck(lib().GXProfileCreateEvent(b"capi_event", ctypes.byref(ev))); ck(lib().GXProfileCreateCounter(dom, b"capi_counter", ctypes.byref(ctr)))
    assert lib().GXProfileDurationStop(task) == -1
    for h in (task, ev):
        ck(lib().GXProfileDurationStart(h)); ck(lib().GXProfileDurationStop(h))
    ck(lib().GXProfileSetCounter(ctr, ctypes.c_uint64(5))); ck(lib().GXProfileAdjustCounter(ctr, ctypes.c_int64(-2)))
    ck(lib().GXDumpProfile(1)); ck(lib().GXSetProfilerState(0))
    trace = open(prof.decode()).read()
    assert "capi_task" in trace and "capi_event" in trace and '"capi_counter":3' in trace.replace(" ", "")
    for h in (dom, task, ev, ctr):
        ck(lib().GXProfileDestroyHandle(h))
    v, prev = ctypes.c_int(), ctypes.c_int()
    ck(lib().GXSetNumOMPThreads(3)); ck(lib().GXGetNumOMPThreads(ctypes.byref(v))); assert v.value == 3
    ck(lib().GXEngineSetBulkSize(4, ctypes.byref(prev))); ck(lib().GXEngineSetBulkSize(prev.value, ctypes.byref(v))); assert v.value == 4
    ck(lib().GXGetGPUCount(ctypes.byref(v))); assert v.value >= 0
    ck(lib().GXNotifyShutdown())


def test_python_free_parameter_server_job(tmp_path):
    """examples/c_api/ps_node.c: scheduler + server (optimizer = a C callback through GXKVStoreRunServerEx) + two workers, four processes of one
    C binary linked against libgeomx_capi.so only — the HiPS TCP plane without an interpreter in any process."""
    import os
    import shutil
    import socket
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    libdir = os.path.join(C.ROOT, "geomx_b200", "lib")
    exe = str(tmp_path / "ps_node")
    subprocess.run([cc, "-O2", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(C.ROOT, "geomx_b200", "include"), os.path.join(C.ROOT, "examples", "c_api", "ps_node.c"),
                    "-L", libdir, "-lgeomx_capi", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = {k: v for k, v in os.environ.items() if not k.startswith(("DMLC_", "PS_")) and k not in ("RANK", "WORLD_SIZE")}
    env.update({"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(port), "DMLC_NUM_SERVER": "1", "DMLC_NUM_WORKER": "2", "DMLC_NUM_ALL_WORKER": "2"})
    procs = [subprocess.Popen([exe], env=dict(env, DMLC_ROLE=role), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for role in ("scheduler", "server", "worker", "worker")]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=120)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert all(p.returncode == 0 for p in procs), outs
    assert "server done: 2 updates, 1 controller commands" in outs[1]
    for o in outs[2:]:
        vals = [float(l.split()[-1]) for l in o.splitlines() if l.startswith("RESULT")]
        assert len(vals) == 2 and abs(vals[0] - 0.85) < 1e-6 and abs(vals[1] - 0.70) < 1e-6, o          # 1 - 0.1 * (0.5 + 1.0) per round


def _pad_nd(t, widths, mode, value=0.0):
    """numpy-style N-D padding for the torch reference (torch's F.pad has rank restrictions in the non-constant modes)."""
    out = t
    for ax, (b, a) in enumerate(widths):
        if b == 0 and a == 0:
            continue
        n = out.shape[ax]
        idx = torch.arange(-b, n + a)
        if mode == "edge":
            idx = idx.clamp(0, n - 1)
        elif mode == "reflect":
            idx = torch.where(idx < 0, -idx, idx); idx = torch.where(idx >= n, 2 * (n - 1) - idx, idx)
        if mode == "constant":
            shape = list(out.shape); shape[ax] = b; left = torch.full(shape, value); shape[ax] = a; right = torch.full(shape, value)
            out = torch.cat([left, out, right], ax)
        else:
            out = out.index_select(ax, idx)
    return out


# (case id, operator, attributes, input shapes by name, torch reference over the inputs in ListArguments order, inputs without gradient)
_RNG = np.random.RandomState(5)
_IDX = {"take_idx": np.array([[2, 0], [9, 1]], dtype=np.float32), "pick_idx": _RNG.randint(0, 4, (3, 5)).astype(np.float32),
        "cond": (_RNG.rand(3, 4) > 0.5).astype(np.float32), "hot": np.array([[1, 0, 3]], dtype=np.float32), "lab": np.array([2, 0, 1, 2], dtype=np.float32)}
_TIER2 = [
    ("layernorm", "LayerNorm", dict(axis=1, eps=1e-5), {"data": (3, 5, 2)}, lambda x, g, b: TF.layer_norm(x.transpose(1, 2), (5,), g, b, 1e-5).transpose(1, 2), ()),
    ("instancenorm", "InstanceNorm", dict(eps=1e-3), {"data": (2, 3, 4, 2)}, lambda x, g, b: TF.instance_norm(x, weight=g, bias=b, eps=1e-3), ()),
    ("l2norm_channel", "L2Normalization", dict(mode="channel", eps=1e-6), {"data": (2, 3, 4)}, lambda x: x / torch.sqrt((x * x).sum(1, keepdim=True) + 1e-6), ()),
    ("l2norm_instance", "L2Normalization", dict(mode="instance", eps=1e-6), {"data": (2, 3, 4)}, lambda x: x / torch.sqrt((x * x).sum((1, 2), keepdim=True) + 1e-6), ()),
    ("l2norm_spatial", "L2Normalization", dict(mode="spatial", eps=1e-6), {"data": (2, 3, 4)}, lambda x: x / torch.sqrt((x * x).sum(2, keepdim=True) + 1e-6), ()),
    ("lrn", "LRN", dict(nsize=3, alpha=0.3, beta=0.75, knorm=2.0), {"data": (2, 5, 3, 2)}, lambda x: TF.local_response_norm(x, 3, alpha=0.3, beta=0.75, k=2.0), ()),
    ("deconv", "Deconvolution", dict(kernel="(3, 2)", num_filter=4, stride="(2, 2)", pad="(1, 0)", adj="(1, 0)", num_group=2, no_bias=False), {"data": (2, 4, 3, 4)},
     lambda x, w, b: TF.conv_transpose2d(x, w, b, stride=2, padding=(1, 0), output_padding=(1, 0), groups=2), ()),
    ("deconv_dilated", "Deconvolution", dict(kernel="(2, 2)", num_filter=3, dilate="(2, 1)"), {"data": (1, 2, 3, 3)}, lambda x, w: TF.conv_transpose2d(x, w, None, dilation=(2, 1)), ()),
    ("upsampling", "UpSampling", dict(scale=2, sample_type="nearest", num_args=1), {"arg0": (2, 2, 3, 2)}, lambda x: TF.interpolate(x, scale_factor=2, mode="nearest"), ()),
    ("softmax_ce", "softmax_cross_entropy", {}, {"data": (4, 3), "label": "lab"}, lambda x, l: TF.cross_entropy(x, l.long(), reduction="sum").reshape(1), ("label",)),
    ("smooth_l1", "smooth_l1", dict(scalar=2.0), {"data": (4, 5)}, lambda x: torch.where(x.abs() < 0.25, 2.0 * x * x, x.abs() - 0.125), ()),
    ("slice_axis", "slice_axis", dict(axis=1, begin=-3, end="None"), {"data": (2, 5, 3)}, lambda x: x[:, -3:], ()),
    ("slice", "slice", dict(begin="(None, 1)", end="(2, -1)"), {"data": (3, 5, 2)}, lambda x: x[:2, 1:-1], ()),
    ("swapaxis", "SwapAxis", dict(dim1=0, dim2=2), {"data": (2, 3, 4)}, lambda x: x.transpose(0, 2), ()),
    ("tile", "tile", dict(reps="(2, 1, 3)"), {"data": (2, 3)}, lambda x: x.repeat(2, 1, 3), ()),
    ("repeat", "repeat", dict(repeats=3, axis=1), {"data": (2, 3, 2)}, lambda x: x.repeat_interleave(3, 1), ()),
    ("pad_constant", "Pad", dict(mode="constant", pad_width="(0, 0, 1, 2, 2, 0)", constant_value=1.5), {"data": (2, 3, 2)}, lambda x: _pad_nd(x, [(0, 0), (1, 2), (2, 0)], "constant", 1.5), ()),
    ("pad_edge", "Pad", dict(mode="edge", pad_width="(0, 0, 0, 0, 1, 2, 2, 1)"), {"data": (1, 2, 3, 3)}, lambda x: _pad_nd(x, [(0, 0), (0, 0), (1, 2), (2, 1)], "edge"), ()),
    ("pad_reflect", "Pad", dict(mode="reflect", pad_width="(0, 0, 0, 0, 2, 1, 1, 2)"), {"data": (1, 2, 3, 4)}, lambda x: _pad_nd(x, [(0, 0), (0, 0), (2, 1), (1, 2)], "reflect"), ()),
    ("squeeze", "squeeze", dict(axis="(1,)"), {"data": (3, 1, 2, 1)}, lambda x: x.squeeze(1), ()),
    ("broadcast_to", "broadcast_to", dict(shape="(0, 4, 3)"), {"data": (2, 1, 1)}, lambda x: x.expand(2, 4, 3), ()),
    ("broadcast_axis", "broadcast_axis", dict(axis="(0, 2)", size="(3, 2)"), {"data": (1, 4, 1)}, lambda x: x.expand(3, 4, 2), ()),
    ("reverse", "reverse", dict(axis="(0, 2)"), {"data": (2, 3, 4)}, lambda x: x.flip(0, 2), ()),
    ("take", "take", dict(axis=1), {"a": (3, 4, 2), "indices": "take_idx"}, lambda a, i: a[:, i.long().clamp(0, 3)], ("indices",)),
    ("pick", "pick", dict(axis=1, keepdims=True), {"data": (3, 4, 5), "index": "pick_idx"}, lambda d, i: d.gather(1, i.long().unsqueeze(1)), ("index",)),
    ("one_hot", "one_hot", dict(depth=4, on_value=2.0, off_value=-1.0), {"indices": "hot"}, lambda i: TF.one_hot(i.long(), 4).float() * 3.0 - 1.0, ("indices",)),
    ("where", "where", {}, {"condition": "cond", "x": (3, 4), "y": (3, 4)}, lambda c, x, y: torch.where(c != 0, x, y), ("condition",)),
    ("cast", "Cast", dict(dtype="float32"), {"data": (2, 3)}, lambda x: x * 1.0, ()),
    ("max", "max", dict(axis="(1,)", keepdims=True), {"data": (3, 4, 2)}, lambda x: x.amax(1, keepdim=True), ()),
    ("min_all", "min", {}, {"data": (3, 4)}, lambda x: x.amin().reshape(1), ()),
    ("prod", "prod", dict(axis="(0, 2)"), {"data": (2, 3, 2)}, lambda x: x.prod(2).prod(0), ()),
    ("norm", "norm", dict(axis="(1,)"), {"data": (3, 4)}, lambda x: x.norm(dim=1), ()),
    ("argmax", "argmax", dict(axis=1), {"data": (3, 5, 2)}, lambda x: x.argmax(1).float(), ("data",)),
    ("argmin_keep", "argmin", dict(axis=-1, keepdims=True), {"data": (3, 5)}, lambda x: x.argmin(-1, keepdim=True).float(), ("data",)),
    ("power", "broadcast_power", {}, {"lhs": (3, 1, 2), "rhs": (4, 1)}, lambda l, r: l ** r, ()),
    ("greater_equal", "broadcast_greater_equal", {}, {"lhs": (3, 4), "rhs": (1, 4)}, lambda l, r: (l >= r).float(), ("lhs", "rhs")),
    ("not_equal", "broadcast_not_equal", {}, {"lhs": (3, 4), "rhs": (3, 4)}, lambda l, r: (l != r).float(), ("lhs", "rhs")),
    ("max_scalar", "_maximum_scalar", dict(scalar=1.0), {"data": (3, 4)}, lambda x: x.clamp(min=1.0), ()),
    ("min_scalar", "_minimum_scalar", dict(scalar=1.0), {"data": (3, 4)}, lambda x: x.clamp(max=1.0), ()),
    ("rpower_scalar", "_rpower_scalar", dict(scalar=1.7), {"data": (3, 4)}, lambda x: 1.7 ** x, ()),
] + [(u, u, {}, {"data": (3, 4)}, f, ()) for u, f in [
    ("sin", torch.sin), ("cos", torch.cos), ("tan", torch.tan), ("arctan", torch.atan), ("sinh", torch.sinh), ("cosh", torch.cosh), ("log1p", torch.log1p), ("expm1", torch.expm1),
    ("log2", torch.log2), ("log10", torch.log10), ("rsqrt", torch.rsqrt), ("reciprocal", torch.reciprocal), ("cbrt", lambda x: x ** (1.0 / 3)), ("erf", torch.erf),
    ("floor", torch.floor), ("ceil", torch.ceil), ("sign", torch.sign)]] + [
    ("arcsin", "arcsin", {}, {"data": (3, 4)}, lambda x: torch.asin(x / 2), ()), ("arccos", "arccos", {}, {"data": (3, 4)}, lambda x: torch.acos(x / 2), ())]


@pytest.mark.parametrize("case", _TIER2, ids=[c[0] for c in _TIER2])
def test_second_tier_operators_match_torch(case):
    cid, opname, attrs, shapes, ref, no_grad = case
    rng = np.random.RandomState(abs(hash(cid)) % 1000)
    names = list(shapes.keys())
    values = {}
    for k, v in shapes.items():
        values[k] = _IDX[v] if isinstance(v, str) else (rng.rand(*v) + 0.5).astype(np.float32)        # (0.5, 1.5): inside every domain used here
    halve = cid in ("arcsin", "arccos")
    ins = [C.var(k) for k in names]
    if halve:
        ins = [C.op("_mul_scalar", "half", ins, scalar=0.5)]
    sym = C.op(opname, "n", ins, **attrs)
    args = C.list_arguments(sym)
    extra = [a for a in args if a not in names]                                   # learned parameters the operator declared itself (gamma, weight, ...)
    a_, o_, _, _ = C.infer_shape(sym, partial=True, **{k: values[k].shape for k in names})
    for a, shp in zip(args, a_):
        if a in extra:
            values[a] = (rng.rand(*shp) + 0.5).astype(np.float32)
    ex, hargs, hgrads, _ = C.simple_bind(sym, {k: values[k].shape for k in names}, no_grad=tuple(no_grad))
    for k in args:
        C.nd_set(hargs[k], values[k])
    out = C.forward(ex, True)[0]
    tin = [torch.tensor(values[k], requires_grad=k not in no_grad) for k in args]
    tout = ref(*tin)
    assert out.shape == tuple(tout.shape), (out.shape, tuple(tout.shape))
    assert np.allclose(out, tout.detach().numpy(), rtol=1e-4, atol=1e-5), np.abs(out - tout.detach().numpy()).max()
    head = rng.rand(*out.shape).astype(np.float32)
    C.backward(ex, [C.nd_create(head)])
    if tout.requires_grad:
        tout.backward(torch.tensor(head))
        for k, t in zip(args, tin):
            if k in no_grad:
                continue
            want = t.grad.numpy() if t.grad is not None else np.zeros_like(values[k])
            got = C.nd_get(hgrads[k])
            assert np.allclose(got, want, rtol=2e-3, atol=2e-5 * max(1.0, np.abs(want).max())), (k, np.abs(got - want).max())
    ck(lib().GXExecutorFree(ex))


def test_multi_output_operator_slice_channel():
    """SliceChannel / split: the multi-output operator — secondary outputs in graphs, JSON, the executor and the autograd history."""
    rng = np.random.RandomState(9)
    x = C.var("data")
    parts = C.op("SliceChannel", "sp", [x], num_outputs=3, axis=1)
    assert C.list_outputs(parts) == ["sp_output0", "sp_output1", "sp_output2"]
    p0, p2 = vp(), vp()
    ck(lib().GXSymbolGetOutput(parts, 0, ctypes.byref(p0))); ck(lib().GXSymbolGetOutput(parts, 2, ctypes.byref(p2)))
    # y = tanh(part0) * part2, part1 unused; plus part2 itself as a second head
    y = C.op("elemwise_mul", "m", [C.op("tanh", "t", [p0]), p2])
    grp = vp(); ck(lib().GXSymbolCreateGroup(2, C.handles([y, p2]), ctypes.byref(grp)))
    assert C.list_outputs(grp) == ["m_output", "sp_output2"]
    _, o, _, ok = C.infer_shape(grp, data=(2, 6, 4))
    assert ok and o == [(2, 2, 4), (2, 2, 4)]
    with pytest.raises(RuntimeError, match="not divisible"):
        C.infer_shape(grp, data=(2, 5, 4))
    js = C.sym_json(grp)
    import json as _json
    doc = _json.loads(js)
    sp = [i for i, n in enumerate(doc["nodes"]) if n["name"] == "sp"][0]
    assert doc["node_row_ptr"][sp + 1] - doc["node_row_ptr"][sp] == 3 and [sp, 2, 0] in doc["heads"]
    again = C.sym_from_json(js)
    assert C.list_outputs(again) == ["m_output", "sp_output2"] and C.sym_json(again) == js
    import geomx_b200 as mx
    psym = mx.sym.load_json(js)                              # the Python front end reads the secondary outputs
    assert psym.list_outputs() == ["m_output", "sp_output2"] and psym.infer_shape(data=(2, 6, 4))[1] == [(2, 2, 4), (2, 2, 4)]
    internals = vp(); ck(lib().GXSymbolGetInternals(grp, ctypes.byref(internals)))
    assert [n for n in C.list_outputs(internals) if n.startswith("sp_")] == ["sp_output0", "sp_output1", "sp_output2"]
    ex, args, grads, _ = C.simple_bind(grp, {"data": (2, 6, 4)})
    xn = rng.randn(2, 6, 4).astype(np.float32)
    C.nd_set(args["data"], xn)
    out = C.forward(ex, True)
    h0, h1 = rng.rand(2, 2, 4).astype(np.float32), rng.rand(2, 2, 4).astype(np.float32)
    C.backward(ex, [C.nd_create(h0), C.nd_create(h1)])
    tx = torch.tensor(xn, requires_grad=True)
    a, b, c = tx.split(2, 1)
    ty = torch.tanh(a) * c
    (ty * torch.tensor(h0)).sum().backward(retain_graph=True); (c * torch.tensor(h1)).sum().backward()
    assert np.allclose(out[0], ty.detach().numpy(), atol=1e-6) and np.allclose(out[1], c.detach().numpy())
    assert np.allclose(C.nd_get(grads["data"]), tx.grad.numpy(), atol=1e-6) and np.allclose(C.nd_get(grads["data"])[:, 2:4], 0.0)
    ck(lib().GXExecutorFree(ex))
    # squeeze_axis through the imperative path, several outputs, one of them never used
    v = C.nd_create(xn[:, :3]); g = C.nd_create(np.zeros((2, 3, 4)))
    C.mark_variables([v], [g])
    n, outs = ctypes.c_int(0), ctypes.POINTER(vp)()
    with C.record():
        ck(lib().GXImperativeInvokeByName(b"split", 1, C.handles([v]), ctypes.byref(n), ctypes.byref(outs), 3, C.strs(["num_outputs", "axis", "squeeze_axis"]),
                                          C.strs(["3", "1", "True"])))
        rows = [vp(outs[i]) for i in range(n.value)]
        z = C.invoke("broadcast_add", [C.invoke("exp", [rows[2]]), rows[0]])
    assert n.value == 3 and C.nd_shape(rows[1]) == (2, 4)
    sym = vp(); ck(lib().GXAutogradGetSymbol(z, ctypes.byref(sym)))
    assert "SliceChannel" in _print(sym)
    C.ag_backward([z])
    want = np.zeros((2, 3, 4), dtype=np.float32); want[:, 0] = 1.0; want[:, 2] = np.exp(xn[:, 2])
    assert np.allclose(C.nd_get(g), want, rtol=1e-5)
    bad = vp()
    assert lib().GXSymbolCreateFromJSON(js.replace('[%d, 2, 0]' % sp, '[%d, 7, 0]' % sp).encode(), ctypes.byref(bad)) == -1 and "has no output 7" in C.err()


def test_python_free_distributed_training(tmp_path):
    """examples/c_api/dist_train_cnn.c: the reference's cnn.py flow (init, per-step forward / backward, push + pull of every parameter) as four C
    processes — two workers computing with GXExecutor*, a server applying SGD natively, a scheduler.  Workers must end with identical parameters."""
    import os
    import re
    import shutil
    import socket
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    libdir = os.path.join(C.ROOT, "geomx_b200", "lib")
    exe = str(tmp_path / "dist_train_cnn")
    subprocess.run([cc, "-O2", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(C.ROOT, "geomx_b200", "include"), os.path.join(C.ROOT, "examples", "c_api", "dist_train_cnn.c"),
                    "-L", libdir, "-lgeomx_capi", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = {k: v for k, v in os.environ.items() if not k.startswith(("DMLC_", "PS_")) and k not in ("RANK", "WORLD_SIZE")}
    env.update({"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(port), "DMLC_NUM_SERVER": "1", "DMLC_NUM_WORKER": "2", "DMLC_NUM_ALL_WORKER": "2"})
    procs = [subprocess.Popen([exe, "30"], env=dict(env, DMLC_ROLE=role), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for role in ("scheduler", "server", "worker", "worker")]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=240)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert all(p.returncode == 0 for p in procs), outs
    finals = [re.search(r"FINAL rank (\d) of 2 loss ([\d.]+) -> ([\d.]+) checksum ([-\d.]+)", o) for o in outs[2:]]
    assert all(finals), outs
    assert sorted(m.group(1) for m in finals) == ["0", "1"]
    assert finals[0].group(4) == finals[1].group(4)                      # the same parameters on both workers, to the last printed digit
    assert all(float(m.group(3)) < 0.3 * float(m.group(2)) for m in finals)


def test_predict_api_falls_back_to_the_general_executor(tmp_path):
    """Graphs with operators outside the planned predictor's set (LayerNorm, SliceChannel, Deconvolution ...) are served by the general executor
    behind the same GXPred* ABI; graphs inside the set keep the planned engine."""
    rng = np.random.RandomState(11)
    x = C.var("data")
    h = C.op("Deconvolution", "up", [x], kernel="(2, 2)", stride="(2, 2)", num_filter=4)                    # (N, 4, 8, 8)
    parts = C.op("SliceChannel", "sp", [h], num_outputs=2, axis=1)
    a, b = vp(), vp()
    ck(lib().GXSymbolGetOutput(parts, 0, ctypes.byref(a))); ck(lib().GXSymbolGetOutput(parts, 1, ctypes.byref(b)))
    h = C.op("broadcast_mul", "gate", [C.op("sigmoid", "s", [a]), b])                                       # (N, 2, 8, 8)
    h = C.op("LayerNorm", "ln", [C.op("Flatten", "f", [h])], axis=-1)
    net = C.op("softmax", "prob", [C.op("FullyConnected", "fc", [h], num_hidden=3)])
    args = C.list_arguments(net)
    shapes, _, _, _ = C.infer_shape(net, data=(2, 3, 4, 4))
    values = {n: (rng.randn(*s) * 0.5).astype(np.float32) for n, s in zip(args, shapes)}
    fname = str(tmp_path / "m.params").encode()
    keys = [n for n in args if n != "data"]
    ck(lib().GXNDArraySave(fname, len(keys), C.handles([C.nd_create(values[k]) for k in keys]), C.strs(["arg:" + k for k in keys])))
    blob = open(fname, "rb").read()
    js = C.sym_json(net).encode()

    def create(json_bytes, batch, out_keys=None):
        p = vp()
        ind, dims = (u32 * 2)(0, 4), (u32 * 4)(batch, 3, 4, 4)
        if out_keys:
            rc = lib().GXPredCreatePartialOut(json_bytes, blob, len(blob), 1, 0, 1, C.strs(["data"]), ind, dims, len(out_keys), C.strs(out_keys), ctypes.byref(p))
        else:
            rc = lib().GXPredCreate(json_bytes, blob, len(blob), 1, 0, 1, C.strs(["data"]), ind, dims, ctypes.byref(p))
        if rc != 0:
            raise RuntimeError(C.err())
        return p

    def run(p, xin, n_out):
        ck(lib().GXPredSetInput(p, b"data", xin.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), xin.size)); ck(lib().GXPredForward(p))
        out = np.empty(n_out, dtype=np.float32)
        ck(lib().GXPredGetOutput(p, 0, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), out.size))
        return out
    p = create(js, 2)
    eng = ctypes.c_int(); ck(lib().GXPredGetEngine(p, ctypes.byref(eng))); assert eng.value == 2
    xin = rng.rand(2, 3, 4, 4).astype(np.float32)
    got = run(p, xin, 6).reshape(2, 3)
    t = {k: torch.tensor(v) for k, v in values.items()}
    th = TF.conv_transpose2d(torch.tensor(xin), t["up_weight"], None, stride=2)
    ta, tb = th.split(2, 1)
    tn = TF.layer_norm((torch.sigmoid(ta) * tb).flatten(1), (128,), t["ln_gamma"], t["ln_beta"], 1e-5)
    want = torch.softmax(TF.linear(tn, t["fc_weight"], t["fc_bias"]), -1).numpy()
    assert np.allclose(got, want, atol=1e-5)
    sd, nd_ = ctypes.POINTER(u32)(), u32()
    ck(lib().GXPredGetOutputShape(p, 0, ctypes.byref(sd), ctypes.byref(nd_))); assert [sd[i] for i in range(nd_.value)] == [2, 3]
    # reshape to another batch on the same parameters; internal output by name; plan statistics
    p5 = vp(); ck(lib().GXPredReshape(1, C.strs(["data"]), (u32 * 2)(0, 4), (u32 * 4)(5, 3, 4, 4), p, ctypes.byref(p5)))
    x5 = rng.rand(5, 3, 4, 4).astype(np.float32)
    assert np.allclose(run(p5, x5, 15).reshape(5, 3)[:2], run(create(js, 5), x5, 15).reshape(5, 3)[:2])
    pin = create(js, 2, ["sp_output1"])
    assert np.allclose(run(pin, xin, 2 * 2 * 8 * 8).reshape(2, 2, 8, 8), tb.numpy(), atol=1e-5)
    arena, nops = ctypes.c_uint64(), u32()
    ck(lib().GXPredGetPlan(p, ctypes.byref(arena), ctypes.byref(nops))); assert nops.value == 8 and arena.value > 0
    with pytest.raises(RuntimeError, match="unknown input"):
        bad = np.zeros(3, np.float32)
        if lib().GXPredSetInput(p, b"fc_weight", bad.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 3) != 0:
            raise RuntimeError(C.err())
    # a graph inside the planned set keeps engine 1
    small = C.op("softmax", "prob", [C.op("FullyConnected", "fc", [C.op("Flatten", "f", [C.var("data")])], num_hidden=3)])
    w = {"fc_weight": rng.randn(3, 48).astype(np.float32), "fc_bias": np.zeros(3, np.float32)}
    ck(lib().GXNDArraySave(fname, 2, C.handles([C.nd_create(w[k]) for k in w]), C.strs(["arg:" + k for k in w])))
    blob = open(fname, "rb").read()
    ps = create(C.sym_json(small).encode(), 2)
    ck(lib().GXPredGetEngine(ps, ctypes.byref(eng))); assert eng.value == 1
    for h_ in (p, p5, pin, ps):
        ck(lib().GXPredFree(h_))
    # a missing parameter is reported with both engines' reasons
    with pytest.raises(RuntimeError, match="general executor could not run the graph either"):
        create(js, 2)
