"""``mx.nd.contrib`` + the spatial / optimizer-update / sampling operator families against numpy / plain PyTorch references."""
import numpy as np
import torch

import geomx_b200 as mx

nd = mx.nd


def test_box_iou_and_nms():
    a = nd.array([[0, 0, 2, 2], [1, 1, 3, 3]]); b = nd.array([[0, 0, 2, 2], [2, 2, 3, 3], [0, 0, 4, 4]])
    iou = nd.contrib.box_iou(a, b).asnumpy()
    assert iou.shape == (2, 3)
    np.testing.assert_allclose(iou[0], [1.0, 0.0, 0.25], atol=1e-6)
    np.testing.assert_allclose(iou[1, 0], 1.0 / 7.0, atol=1e-6)
    # rows: id, score, box
    d = nd.array([[0, 0.9, 0, 0, 2, 2], [0, 0.8, 0.1, 0.1, 2, 2], [1, 0.7, 0.1, 0.1, 2, 2], [0, 0.05, 5, 5, 6, 6], [0, 0.6, 5, 5, 6, 6]])
    out = nd.contrib.box_nms(d, overlap_thresh=0.5, valid_thresh=0.1, coord_start=2, score_index=1, id_index=0).asnumpy()
    assert out[:3, 1].tolist() == [np.float32(0.9), np.float32(0.7), np.float32(0.6)]     # 0.8 suppressed by 0.9 (same class); class 1 kept
    assert (out[3:] == -1).all()
    out2 = nd.contrib.box_nms(d, overlap_thresh=0.5, valid_thresh=0.1, coord_start=2, score_index=1, id_index=0, force_suppress=True).asnumpy()
    assert out2[:2, 1].tolist() == [np.float32(0.9), np.float32(0.6)] and (out2[2:] == -1).all()


def test_bipartite_matching():
    s = nd.array([[0.5, 0.6], [0.1, 0.2], [0.3, 0.4]])
    r, c = nd.contrib.bipartite_matching(s, threshold=1e-12)
    assert r.asnumpy().tolist() == [1, -1, 0] and c.asnumpy().tolist() == [2, 0]


def test_multibox_prior_target_detection_round_trip():
    feat = nd.zeros((1, 8, 4, 4))
    anchors = nd.contrib.MultiBoxPrior(feat, sizes=(0.4, 0.2), ratios=(1, 2))
    assert anchors.shape == (1, 4 * 4 * 3, 4)
    a0 = anchors.asnumpy()[0, 0]
    np.testing.assert_allclose(a0, [0.125 - 0.2, 0.125 - 0.2, 0.125 + 0.2, 0.125 + 0.2], atol=1e-6)
    gt = np.array([[[1, 0.1, 0.1, 0.5, 0.55], [-1, -1, -1, -1, -1]]], dtype=np.float32)
    N = anchors.shape[1]
    cls_pred = nd.zeros((1, 3, N))
    loc_t, loc_m, cls_t = nd.contrib.MultiBoxTarget(anchors, nd.array(gt), cls_pred)
    ct = cls_t.asnumpy()[0]
    assert (ct == 2).sum() >= 1 and set(np.unique(ct)) <= {0.0, 2.0}
    assert loc_m.asnumpy().reshape(N, 4)[ct == 2].all() and not loc_m.asnumpy().reshape(N, 4)[ct == 0].any()
    # decode the targets of the matched anchors as if they were predictions: must give the ground truth box back
    prob = np.zeros((1, 3, N), dtype=np.float32); prob[0, 0] = 1.0
    pos = np.where(ct == 2)[0]
    prob[0, 0, pos] = 0.05; prob[0, 2, pos] = 0.95
    det = nd.contrib.MultiBoxDetection(nd.array(prob), loc_t, anchors, nms_threshold=0.5).asnumpy()[0]
    assert det[0, 0] == 1 and abs(det[0, 1] - 0.95) < 1e-6
    np.testing.assert_allclose(det[0, 2:], gt[0, 0, 1:], atol=1e-4)
    assert (det[1:, 0] == -1).all()                                                # duplicates of the same object are suppressed


def test_roi_ops_and_proposal():
    x = nd.array(np.arange(2 * 1 * 8 * 8, dtype=np.float32).reshape(2, 1, 8, 8))
    rois = nd.array([[0, 0, 0, 7, 7], [1, 2, 2, 5, 5]])
    out = nd.contrib.ROIAlign(x, rois, (2, 2), 1.0, sample_ratio=2)
    assert out.shape == (2, 1, 2, 2)
    assert float(out[1].asnumpy().mean()) > float(out[0].asnumpy().mean())
    p = nd.ROIPooling(x, rois, (2, 2), 1.0).asnumpy()
    assert p[0, 0, 1, 1] == 63 and p[1, 0, 1, 1] == 64 + 5 * 8 + 5
    B, A, H, W = 1, 3, 4, 4
    g = torch.Generator().manual_seed(0)
    cls = nd.array(torch.rand(B, 2 * A, H, W, generator=g)); bbox = nd.array(torch.randn(B, 4 * A, H, W, generator=g) * 0.1)
    r, s = nd.contrib.Proposal(cls, bbox, nd.array([[64, 64, 1.0]]), rpn_pre_nms_top_n=30, rpn_post_nms_top_n=8, rpn_min_size=4,
                               scales=(2,), ratios=(0.5, 1, 2), feature_stride=16, output_score=True)
    assert r.shape == (8, 5) and s.shape == (8, 1)
    rb = r.asnumpy()
    assert (rb[:, 0] == 0).all() and (rb[:, 1:] >= 0).all() and (rb[:, 1:] <= 63).all()
    assert (np.diff(s.asnumpy()[:, 0]) <= 1e-6).all() or len(np.unique(rb, axis=0)) < 8       # sorted unless padded by repetition


def test_fft_sketch_quantize():
    x = torch.randn(3, 8)
    f = nd.contrib.fft(nd.array(x))
    assert f.shape == (3, 16)
    ref = torch.fft.fft(x)
    np.testing.assert_allclose(f.asnumpy()[:, 0::2], ref.real.numpy(), atol=1e-5)
    np.testing.assert_allclose(f.asnumpy()[:, 1::2], ref.imag.numpy(), atol=1e-5)
    np.testing.assert_allclose(nd.contrib.ifft(f).asnumpy() / 8, x.numpy(), atol=1e-5)
    h = nd.array([0, 2, 2, 1]); s = nd.array([1, -1, 1, 1])
    cs = nd.contrib.count_sketch(nd.array([[1.0, 2.0, 3.0, 4.0]]), h, s, 3).asnumpy()
    np.testing.assert_allclose(cs, [[1.0, 4.0, 1.0]])
    v = nd.array(torch.linspace(-3, 5, 101))
    q, lo, hi = nd.contrib.quantize(v, nd.array([-3.0]), nd.array([5.0]), out_type="uint8")
    assert q.dtype == np.uint8
    np.testing.assert_allclose(nd.contrib.dequantize(q, lo, hi).asnumpy(), v.asnumpy(), atol=8.0 / 255 / 2 + 1e-6)
    q8, lo8, hi8 = nd.contrib.quantize(v, nd.array([-3.0]), nd.array([5.0]), out_type="int8")
    assert q8.dtype == np.int8 and float(hi8.asnumpy()[0]) == 5.0
    np.testing.assert_allclose(nd.contrib.dequantize(q8, lo8, hi8).asnumpy(), v.asnumpy(), atol=5.0 / 127 / 2 + 1e-6)
    acc = nd.array(torch.tensor([1000, -2000, 30000], dtype=torch.int32))
    rq, rlo, rhi = nd.contrib.requantize(acc, nd.array([-1.0]), nd.array([1.0]))
    back = nd.contrib.dequantize(rq, rlo, rhi).asnumpy()
    np.testing.assert_allclose(back, np.array([1000, -2000, 30000]) / 2147483647.0, atol=float(rhi.asnumpy()[0]) / 127)


def test_control_flow_and_misc():
    data = nd.array(np.arange(12, dtype=np.float32).reshape(4, 3))
    outs, st = nd.contrib.foreach(lambda x, s: (x + s, x + s), data, nd.zeros((3,)))
    np.testing.assert_allclose(outs.asnumpy(), np.cumsum(data.asnumpy(), 0))
    np.testing.assert_allclose(st.asnumpy(), data.asnumpy().sum(0))
    outs, vars_ = nd.contrib.while_loop(lambda i, s: i < 3, lambda i, s: (s + i, [i + 1, s + i]), [nd.array([0.0]), nd.array([10.0])], max_iterations=5)
    assert outs[0].shape == (5, 1) and outs[0].asnumpy()[:, 0].tolist() == [10, 11, 13, 0, 0]
    assert float(vars_[0].asscalar()) == 3 and float(vars_[1].asscalar()) == 13
    assert float(nd.contrib.cond(nd.array([1.0]), lambda: nd.array([5.0]), lambda: nd.array([7.0])).asscalar()) == 5
    m = nd.contrib.boolean_mask(data, nd.array([1, 0, 0, 1])).asnumpy()
    np.testing.assert_allclose(m, data.asnumpy()[[0, 3]])
    ia = nd.contrib.index_array(nd.zeros((2, 3))).asnumpy()
    assert ia.shape == (2, 3, 2) and ia[1, 2].tolist() == [1, 2]
    x = nd.array([1.0, 2.0]); x.attach_grad()
    with mx.autograd.record():
        y = (nd.contrib.gradientmultiplier(x, -0.5) * 3).sum()
    y.backward()
    np.testing.assert_allclose(x.grad.asnumpy(), [-1.5, -1.5])
    assert nd.contrib.isnan(nd.array([float("nan"), 1.0])).asnumpy().tolist() == [1, 0]


def test_spatial_transformer_correlation_histogram_ravel():
    x = nd.array(torch.randn(2, 3, 5, 7))
    ident = nd.array([[1, 0, 0, 0, 1, 0]] * 2)
    y = nd.SpatialTransformer(x, ident, target_shape=(5, 7))
    np.testing.assert_allclose(y.asnumpy(), x.asnumpy(), atol=1e-5)
    flow = nd.zeros((2, 2, 5, 7))
    np.testing.assert_allclose(nd.BilinearSampler(x, nd.GridGenerator(flow, "warp")).asnumpy(), x.asnumpy(), atol=1e-5)
    a = torch.randn(1, 2, 6, 6); b = torch.randn(1, 2, 6, 6)
    c = nd.Correlation(nd.array(a), nd.array(b), kernel_size=1, max_displacement=1, stride1=1, stride2=1, pad_size=1).asnumpy()
    assert c.shape == (1, 9, 6, 6)
    ap = torch.nn.functional.pad(a, (1,) * 4); bp = torch.nn.functional.pad(b, (1,) * 4)
    # centre displacement (index 4) is the channel-mean product; displacement (dy=-1, dx=+1) is index 2
    np.testing.assert_allclose(c[0, 4], (a * b).mean(1)[0].numpy(), atol=1e-6)
    ref = (ap[:, :, 1:7, 1:7] * bp[:, :, 0:6, 2:8]).mean(1)[0].numpy()
    np.testing.assert_allclose(c[0, 2], ref, atol=1e-6)
    v = np.random.RandomState(0).rand(100).astype(np.float32)
    cnt, edges = nd.histogram(nd.array(v), bins=5, range=(0, 1))
    rc, re = np.histogram(v, bins=5, range=(0, 1))
    assert cnt.asnumpy().tolist() == rc.tolist(); np.testing.assert_allclose(edges.asnumpy(), re, atol=1e-6)
    idx = np.array([[1, 0, 2], [3, 1, 0]])
    flat = nd.ravel_multi_index(nd.array(idx), shape=(3, 4)).asnumpy()
    assert flat.tolist() == np.ravel_multi_index(idx, (3, 4)).tolist()
    assert nd.unravel_index(nd.array(flat), shape=(3, 4)).asnumpy().tolist() == idx.tolist()


def test_optimizer_update_ops_match_optimizer_classes():
    rs = np.random.RandomState(0)
    w0 = rs.randn(50).astype(np.float32); g = rs.randn(50).astype(np.float32)
    # SGD + momentum against the Optimizer class
    opt = mx.optimizer.SGD(learning_rate=0.1, momentum=0.9, wd=0.01, rescale_grad=0.5)
    w_ref = nd.array(w0); st = opt.create_state(0, w_ref)
    w = nd.array(w0); mom = nd.zeros((50,))
    for _ in range(3):
        opt.update(0, w_ref, nd.array(g), st)
        nd.sgd_mom_update(w, nd.array(g), mom, lr=0.1, momentum=0.9, wd=0.01, rescale_grad=0.5)
    np.testing.assert_allclose(w.asnumpy(), w_ref.asnumpy(), rtol=1e-5, atol=1e-6)
    # adam_update has no bias correction: compare with the closed form
    w = nd.array(w0); m = nd.zeros((50,)); v = nd.zeros((50,))
    nd.adam_update(w, nd.array(g), m, v, lr=0.01, beta1=0.9, beta2=0.999, epsilon=1e-8)
    mm = 0.1 * g; vv = 0.001 * g * g
    np.testing.assert_allclose(w.asnumpy(), w0 - 0.01 * mm / (np.sqrt(vv) + 1e-8), rtol=1e-5, atol=1e-6)
    # plain sgd writes into `out` and leaves weight alone
    w = nd.array(w0); out = nd.zeros((50,))
    nd.sgd_update(w, nd.array(g), lr=0.5, out=out)
    np.testing.assert_allclose(out.asnumpy(), w0 - 0.5 * g, rtol=1e-6); np.testing.assert_allclose(w.asnumpy(), w0)
    # multi-precision: fp16 weight follows the fp32 master
    w16 = nd.array(w0).astype("float16"); w32 = nd.array(w0)
    nd.mp_sgd_update(w16, nd.array(g).astype("float16"), w32, lr=0.1)
    np.testing.assert_allclose(w32.asnumpy(), w0 - 0.1 * g.astype(np.float16).astype(np.float32), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(w16.asnumpy().astype(np.float32), w32.asnumpy(), atol=2e-3)
    for fn, extra in ((nd.rmsprop_update, (nd.zeros((50,)),)), (nd.signsgd_update, ()), (nd.signum_update, (nd.zeros((50,)),)),
                      (nd.ftrl_update, (nd.zeros((50,)), nd.zeros((50,))))):
        w = nd.array(w0)
        fn(w, nd.array(g), *extra, lr=0.01)
        assert np.isfinite(w.asnumpy()).all() and not np.allclose(w.asnumpy(), w0)


def test_sampling_families_have_the_right_moments():
    mx.random.seed(7)
    n = 40000
    assert abs(float(nd.random.exponential(2.0, shape=(n,)).asnumpy().mean()) - 2.0) < 0.05
    gm = nd.random.gamma(2.0, 3.0, shape=(n,)).asnumpy()
    assert abs(gm.mean() - 6.0) < 0.15 and abs(gm.var() - 18.0) < 1.5
    assert abs(float(nd.random.poisson(4.0, shape=(n,)).asnumpy().mean()) - 4.0) < 0.06
    nb = nd.random.negative_binomial(3, 0.4, shape=(n,)).asnumpy()
    assert abs(nb.mean() - 4.5) < 0.15
    gnb = nd.random.generalized_negative_binomial(2.0, 0.5, shape=(n,)).asnumpy()
    assert abs(gnb.mean() - 2.0) < 0.08 and abs(gnb.var() - (2.0 + 0.5 * 4.0)) < 0.4
    per = nd.random.gamma(nd.array([1.0, 5.0]), nd.array([1.0, 2.0]), shape=(2000,)).asnumpy()
    assert per.shape == (2, 2000) and abs(per[0].mean() - 1.0) < 0.1 and abs(per[1].mean() - 10.0) < 0.5
    idx, logp = nd.random.multinomial(nd.array([[0.2, 0.8], [1.0, 0.0]]), shape=(4000,), get_prob=True)
    assert abs(idx.asnumpy()[0].mean() - 0.8) < 0.03 and (idx.asnumpy()[1] == 0).all()
    assert np.allclose(np.exp(logp.asnumpy()[1]), 1.0)
    assert nd.sample_multinomial(nd.array([0.5, 0.5])).shape in ((), (1,))


def test_fused_rnn_op_linalg_factorisations_and_graph_helpers():
    T, N, C, H, L = 5, 3, 4, 6, 2
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(C, H, L, bidirectional=True)
    ws, bs = [], []
    for l in range(L):
        for sfx in ("", "_reverse"):
            ws += [getattr(lstm, "weight_ih_l%d%s" % (l, sfx)).detach().reshape(-1), getattr(lstm, "weight_hh_l%d%s" % (l, sfx)).detach().reshape(-1)]
            bs += [getattr(lstm, "bias_ih_l%d%s" % (l, sfx)).detach(), getattr(lstm, "bias_hh_l%d%s" % (l, sfx)).detach()]
    flat = torch.cat(ws + bs)
    x = torch.randn(T, N, C); h0 = torch.randn(L * 2, N, H); c0 = torch.randn(L * 2, N, H)
    with torch.no_grad():
        ref, (hn, cn) = lstm(x, (h0, c0))
    out, h, c = nd.RNN(nd.array(x), nd.array(flat), nd.array(h0), nd.array(c0), state_size=H, num_layers=L, bidirectional=True, mode="lstm", state_outputs=True)
    for a, b in ((out, ref), (h, hn), (c, cn)):
        np.testing.assert_allclose(a.asnumpy(), b.numpy(), atol=1e-5)
    gru = torch.nn.GRU(C, H, 1)
    gflat = torch.cat([gru.weight_ih_l0.detach().reshape(-1), gru.weight_hh_l0.detach().reshape(-1), gru.bias_ih_l0.detach(), gru.bias_hh_l0.detach()])
    with torch.no_grad():
        gref = gru(x, h0[:1])[0]
    np.testing.assert_allclose(nd.RNN(nd.array(x), nd.array(gflat), nd.array(h0[:1]), state_size=H, mode="gru").asnumpy(), gref.numpy(), atol=1e-5)
    # the flat vector produced by the FusedRNN initializer has exactly the size this operator expects
    f = nd.zeros((int(flat.numel()),)); mx.init.FusedRNN(mx.init.Uniform(0.1), H, L, "lstm", bidirectional=True)(mx.init.InitDesc("p"), f)
    assert nd.RNN(nd.array(x), f, nd.array(h0), nd.array(c0), state_size=H, num_layers=L, bidirectional=True, mode="lstm").shape == (T, N, 2 * H)
    try:
        nd.RNN(nd.array(x), f[:-1], nd.array(h0), nd.array(c0), state_size=H, num_layers=L, bidirectional=True, mode="lstm")
        raise AssertionError("a short parameter vector must be rejected")
    except ValueError:
        pass

    A = torch.randn(3, 5)
    q, l = nd.linalg_gelqf(nd.array(A))
    np.testing.assert_allclose((l._t @ q._t).numpy(), A.numpy(), atol=1e-5)
    np.testing.assert_allclose((q._t @ q._t.T).numpy(), np.eye(3), atol=1e-5)
    assert (np.diag(l.asnumpy()) >= 0).all() and np.allclose(np.triu(l.asnumpy(), 1), 0)
    S = torch.randn(4, 4); S = S + S.T
    u, w = nd.linalg_syevd(nd.array(S))
    np.testing.assert_allclose((u._t.T @ torch.diag(w._t) @ u._t).numpy(), S.numpy(), atol=1e-4)
    assert (np.diff(w.asnumpy()) >= 0).all()

    assert nd.random.normal_like(nd.zeros((2, 3))).shape == (2, 3) and nd.BatchNorm_v1 is not None and nd.cast_storage(nd.ones((2, 2)), "csr").stype == "csr"
    samp, trials = nd.random.unique_zipfian(50, shape=(2, 8))
    s = samp.asnumpy()
    assert s.shape == (2, 8) and all(len(set(r)) == 8 for r in s) and s.min() >= 0 and s.max() < 50 and (trials.asnumpy() >= 8).all()

    g = nd.sparse.csr_matrix((np.arange(1, 6, dtype=np.float32), np.array([1, 2, 0, 2, 0]), np.array([0, 2, 4, 5])), shape=(3, 3))
    assert nd.contrib.edge_id(g, nd.array([0, 1, 2]), nd.array([2, 1, 0])).asnumpy().tolist() == [2, -1, 5]
    assert nd.contrib.dgl_adjacency(g).asnumpy().tolist() == [[0, 1, 1], [1, 0, 1], [1, 0, 0]]
    sg, mp = nd.contrib.dgl_subgraph(g, nd.array([0, 2]), return_mapping=True)
    assert sg.asnumpy().tolist() == [[0, 0], [1, 0]] and mp.asnumpy().tolist() == [[0, 2], [5, 0]]
    w_ = nd.ones((3, 2)); hist = nd.zeros((3,))
    nd.contrib.group_adagrad_update(w_, nd.array([[1.0, 1.0], [2.0, 2.0], [0.0, 0.0]]), hist, lr=0.1)
    assert hist.asnumpy().tolist() == [1, 4, 0] and np.allclose(w_.asnumpy()[:2], 0.9, atol=1e-4) and w_.asnumpy()[2].tolist() == [1, 1]

    # simulated int8 FC: accumulators times (scale_data * scale_weight) reproduce the float product
    xd = torch.randn(4, 6); wd = torch.randn(3, 6)
    qx, lox, hix = nd.contrib.quantize(nd.array(xd), nd.array([float(xd.min())]), nd.array([float(xd.max())]), "int8")
    qw, low, hiw = nd.contrib.quantize(nd.array(wd), nd.array([float(wd.min())]), nd.array([float(wd.max())]), "int8")
    acc, lo, hi = nd.contrib.quantized_fully_connected(qx, qw, None, lox, hix, low, hiw, no_bias=True)
    real = nd.contrib.dequantize(acc, lo, hi).asnumpy()
    assert np.abs(real - (xd @ wd.T).numpy()).max() < 0.15
    p, plo, phi = nd.contrib.quantized_pooling(nd.array(torch.randint(-100, 100, (1, 2, 4, 4)).to(torch.int8)), lox, hix, kernel=(2, 2), pool_type="max")
    assert p.shape == (1, 2, 2, 2) and p.dtype == np.int8


def test_deformable_psroi_pooling():
    G = P = 2; D = 3
    rois = nd.array([[0, 1.0, 1.0, 6.0, 6.0]])
    xc = torch.arange(D * G * G).float().view(1, -1, 1, 1).expand(1, D * G * G, 8, 8).contiguous()
    oc = nd.contrib.DeformablePSROIPooling(nd.array(xc), rois, spatial_scale=1.0, output_dim=D, group_size=G, pooled_size=P, sample_per_part=2, no_trans=True).asnumpy()
    assert oc[0].reshape(-1).tolist() == list(range(12))               # bin (c, gh, gw) reads channel (c*G + gh)*G + gw
    ramp = torch.arange(8.0).view(1, 1, 1, 8).expand(1, D * G * G, 8, 8).contiguous()
    kw = dict(spatial_scale=1.0, output_dim=D, group_size=G, pooled_size=P, sample_per_part=2, trans_std=0.1)
    t0 = nd.contrib.DeformablePSROIPooling(nd.array(ramp), rois, nd.zeros((1, 2, P, P)), **kw).asnumpy()
    shift = nd.array(torch.cat([torch.ones(1, 1, P, P), torch.zeros(1, 1, P, P)], 1))
    t1 = nd.contrib.DeformablePSROIPooling(nd.array(ramp), rois, shift, **kw).asnumpy()
    np.testing.assert_allclose(t0[0, 0], [[1.25, 4.25]] * 2, atol=1e-5)
    np.testing.assert_allclose(t1 - t0, 0.1 * 6.0, atol=1e-5)          # offset = trans * trans_std * roi width, in pixels of a unit ramp


def test_smoke_of_rarely_used_ops():
    """One call each, against a closed form where there is a cheap one."""
    x = nd.array(torch.randn(2, 8, 6, 6))
    rois = nd.array([[0, 0.0, 0.0, 4.0, 4.0], [1, 1.0, 1.0, 5.0, 5.0]])
    assert nd.contrib.PSROIPooling(x, rois, spatial_scale=1.0, output_dim=2, pooled_size=2).shape == (2, 2, 2, 2)
    w = nd.array(torch.randn(4, 8, 3, 3)); off = nd.zeros((2, 18, 6, 6))
    dc = nd.contrib.DeformableConvolution(x, off, w, kernel=(3, 3), pad=(1, 1), num_filter=4, no_bias=True)
    np.testing.assert_allclose(dc.asnumpy(), torch.nn.functional.conv2d(x._t, w._t, padding=1).numpy(), atol=1e-4)    # zero offsets = plain conv
    outs, st = nd.contrib.foreach(lambda xs, s: ([xs[0] + xs[1], xs[0] * s], s + 1), [nd.ones((3, 2)), nd.ones((3, 2)) * 2], nd.zeros((2,)))
    assert outs[0].shape == (3, 2) and outs[1].asnumpy()[2].tolist() == [2, 2] and st.asnumpy().tolist() == [3, 3]
    assert nd.SVMOutput(nd.ones((2, 3))).shape == (2, 3) and nd.Crop(x, h_w=(4, 4), center_crop=True).shape == (2, 8, 4, 4)
    assert nd.Crop(x, nd.zeros((1, 1, 3, 5)), num_args=2).shape == (2, 8, 3, 5)
    f = nd.fill_element_0index(nd.zeros((2, 3)), nd.array([5.0, 6.0]), nd.array([2, 0])).asnumpy()
    assert f.tolist() == [[0, 0, 5], [6, 0, 0]] and nd.choose_element_0index(nd.array(f), nd.array([2, 0])).asnumpy().tolist() == [5, 6]
    w0 = np.random.RandomState(0).randn(6).astype(np.float32); g = np.random.RandomState(1).randn(6).astype(np.float32)
    for fn, states in ((nd.rmspropalex_update, 3), (nd.ftml_update, 3)):
        wv = nd.array(w0); sts = [nd.zeros((6,)) for _ in range(states)]
        kw = {"t": 1} if fn is nd.ftml_update else {}
        fn(wv, nd.array(g), *sts, lr=0.01, **kw)
        assert np.isfinite(wv.asnumpy()).all() and not np.allclose(wv.asnumpy(), w0)
    w16 = nd.array(w0).astype("float16"); w32 = nd.array(w0); mom = nd.zeros((6,))
    nd.mp_sgd_mom_update(w16, nd.array(g).astype("float16"), mom, w32, lr=0.1, momentum=0.9)
    np.testing.assert_allclose(w32.asnumpy(), w0 - 0.1 * g.astype(np.float16).astype(np.float32), atol=1e-5)
    q = nd.array(torch.randint(-100, 100, (1, 2, 4, 4)).to(torch.int8)); lo, hi = nd.array([-1.0]), nd.array([1.0])
    assert nd.contrib.quantized_flatten(q, lo, hi)[0].shape == (1, 32)
    cat, clo, chi = nd.contrib.quantized_concat(q, q, lo, hi, nd.array([-2.0]), nd.array([2.0]), dim=1)
    assert cat.shape == (1, 4, 4, 4) and float(chi.asnumpy()[0]) == 2.0
    np.testing.assert_allclose(cat.asnumpy()[:, 2:], q.asnumpy())                               # the wider-range input is unchanged
    np.testing.assert_allclose(cat.asnumpy()[:, :2], np.round(q.asnumpy() / 2.0).clip(-127, 127))   # the narrower one re-scaled to it
    xd = torch.randn(1, 3, 5, 5); wd = torch.randn(2, 3, 3, 3)
    qx, lox, hix = nd.contrib.quantize(nd.array(xd), nd.array([float(xd.min())]), nd.array([float(xd.max())]), "int8")
    qw, low, hiw = nd.contrib.quantize(nd.array(wd), nd.array([float(wd.min())]), nd.array([float(wd.max())]), "int8")
    acc, alo, ahi = nd.contrib.quantized_conv(qx, qw, None, lox, hix, low, hiw, kernel=(3, 3), num_filter=2, no_bias=True)
    assert np.abs(nd.contrib.dequantize(acc, alo, ahi).asnumpy() - torch.nn.functional.conv2d(xd, wd).numpy()).max() < 0.3
    reg = nd.IdentityAttachKLSparseReg(nd.array([[0.2, 0.8]]))
    assert reg.asnumpy().tolist() == [[np.float32(0.2), np.float32(0.8)]]
    assert nd.getnnz if hasattr(nd, "getnnz") else nd.contrib.getnnz(nd.array([[0.0, 1.0], [2.0, 0.0]])).asnumpy().tolist() == [2]
