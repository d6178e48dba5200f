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
