"""mx.nd operator library (ndarray/op_lib.py): spot checks of every family against numpy / closed forms, and differentiability."""
import numpy as np

import geomx_b200 as mx

nd = mx.nd


def test_shape_index_and_broadcast_ops():
    a = nd.array(np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    an = a.asnumpy()
    assert np.array_equal(nd.slice_axis(a, axis=2, begin=1, end=3).asnumpy(), an[:, :, 1:3])
    assert np.array_equal(nd.slice(a, (0, 1, 0), (2, 3, 4), (1, 1, 2)).asnumpy(), an[0:2, 1:3, 0:4:2])
    assert nd.swapaxes(a, 0, 2).shape == (4, 3, 2) and nd.expand_dims(a, 1).shape == (2, 1, 3, 4) and nd.squeeze(nd.expand_dims(a, 0)).shape == (2, 3, 4)
    assert np.array_equal(nd.take(a, nd.array([0, 2]), axis=1).asnumpy(), an[:, [0, 2]])
    assert np.array_equal(nd.flip(a, 2).asnumpy(), an[:, :, ::-1]) and np.array_equal(nd.repeat(a, 2, axis=0).asnumpy(), np.repeat(an, 2, 0))
    assert np.array_equal(nd.broadcast_add(a, nd.array(np.ones((1, 3, 1), dtype=np.float32))).asnumpy(), an + 1)
    assert nd.greater(a, 5).asnumpy().sum() == 18 and nd.logical_and(a, nd.zeros_like(a)).asnumpy().sum() == 0
    outs = nd.split(a, 3, axis=1)
    assert len(outs) == 3 and outs[0].shape == (2, 1, 4)
    idx = nd.array(np.array([[0, 1], [1, 2]]))
    assert np.array_equal(nd.gather_nd(a, idx).asnumpy(), an[[0, 1], [1, 2]])
    assert np.array_equal(nd.scatter_nd(nd.array([5.0, 7.0]), idx, (2, 3)).asnumpy(), np.array([[0, 5, 0], [0, 0, 7]], dtype=np.float32))
    assert nd.eye(3).asnumpy().trace() == 3 and nd.linspace(0, 1, 5).shape == (5,) and np.allclose(nd.prod(a + 1, axis=(1, 2)).asnumpy()[0], np.prod(an[0] + 1))
    p = nd.Pad(nd.array(np.ones((1, 1, 2, 2), dtype=np.float32)), mode="constant", pad_width=(0, 0, 0, 0, 1, 1, 2, 2), constant_value=3.0)
    assert p.shape == (1, 1, 4, 6) and p.asnumpy()[0, 0, 0, 0] == 3.0


def test_nn_sequence_and_linalg_ops():
    rng = np.random.RandomState(0)
    x = nd.array(rng.randn(2, 3, 8, 8).astype(np.float32)); w = nd.array(rng.randn(4, 3, 3, 3).astype(np.float32))
    assert nd.Convolution(x, w, None, kernel=(3, 3), num_filter=4, no_bias=True).shape == (2, 4, 6, 6)
    assert nd.Deconvolution(nd.Convolution(x, w, None, kernel=(3, 3), num_filter=4, no_bias=True), w, kernel=(3, 3), num_filter=3).shape == (2, 3, 8, 8)
    assert nd.Pooling(x, kernel=(2, 2), stride=(2, 2), pool_type="avg").shape == (2, 3, 4, 4) and nd.Pooling(x, global_pool=True).shape == (2, 3, 1, 1)
    assert np.allclose(nd.L2Normalization(x).asnumpy().reshape(2, -1).__pow__(2).sum(1), 1.0, atol=1e-4)
    s = nd.SequenceMask(nd.array(np.ones((4, 2, 3), dtype=np.float32)), nd.array([2, 3]), True).asnumpy()
    assert s[:, 0].sum() == 6 and s[:, 1].sum() == 9
    seq = nd.array(np.arange(8, dtype=np.float32).reshape(4, 2, 1))
    assert nd.SequenceLast(seq, nd.array([2, 4]), True).asnumpy().reshape(-1).tolist() == [2.0, 7.0]
    assert nd.SequenceReverse(seq, nd.array([2, 4]), True).asnumpy()[:, 0, 0].tolist() == [2.0, 0.0, 4.0, 6.0]
    a = nd.array(rng.randn(2, 3, 4).astype(np.float32))
    assert np.allclose(nd.batch_dot(a, a, transpose_b=True).asnumpy(), np.einsum("bij,bkj->bik", a.asnumpy(), a.asnumpy()), atol=1e-5)
    A = rng.randn(3, 3).astype(np.float32); spd = A @ A.T + 3 * np.eye(3, dtype=np.float32)
    L = nd.linalg_potrf(nd.array(spd))
    assert np.allclose(L.asnumpy() @ L.asnumpy().T, spd, atol=1e-4)
    assert abs(float(nd.linalg_sumlogdiag(L).asnumpy()) - np.log(np.diag(L.asnumpy())).sum()) < 1e-5
    B = nd.array(rng.randn(3, 2).astype(np.float32))
    X = nd.linalg_trsm(L, B)
    assert np.allclose(L.asnumpy() @ X.asnumpy(), B.asnumpy(), atol=1e-4)
    assert np.allclose(nd.linalg_syrk(nd.array(A)).asnumpy(), A @ A.T, atol=1e-5)


def test_op_lib_is_differentiable():
    z = nd.array(np.random.RandomState(1).randn(5).astype(np.float32)); z.attach_grad()
    with mx.autograd.record():
        y = nd.sin(z) * nd.expm1(z) + nd.smooth_l1(z, scalar=1.0)
    y.backward()
    zn = z.asnumpy()
    ref = np.cos(zn) * np.expm1(zn) + np.sin(zn) * np.exp(zn) + np.where(np.abs(zn) < 1, zn, np.sign(zn))
    assert np.allclose(z.grad.asnumpy(), ref, atol=1e-5)
    mx.test_utils.check_numeric_gradient(lambda a: nd.LeakyReLU(nd.broadcast_mul(a, a), act_type="elu", slope=1.0), [np.random.randn(3, 4)])
