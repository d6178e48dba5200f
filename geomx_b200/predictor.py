"""Deployment predictor over the native C predict API (csrc/runtime/{predict.h,c_predict_api.cc}): symbol JSON + ``.params`` bytes in,
outputs as numpy arrays out — no PyTorch tensor is created, the graph runs in the C++ runtime with its own memory plan.

    pred = Predictor(open("lenet-symbol.json").read(), open("lenet-0010.params", "rb").read(), {"data": (1, 1, 28, 28)})
    pred.forward(data=x)                    # numpy array
    prob = pred.get_output(0)

Parity: the ctypes wrapper the reference ships next to its predict ABI (include/mxnet/c_predict_api.h; ``Predictor.forward / reshape /
get_output`` and ``load_ndarray_file``).  ``dev_type='gpu'`` is served by the Python ``Executor`` on a CUDA device instead (same class
interface) because device tensors belong to PyTorch in this framework."""
import ctypes
import glob
import os

import numpy as np

from .base import MXNetError

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        from . import build
        paths = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "_C*.so"))
        if not paths:
            build.build_runtime()
            paths = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "_C*.so"))
        _LIB = ctypes.CDLL(paths[0])
        _LIB.GXRTGetLastError.restype = ctypes.c_char_p
    return _LIB


def _ck(rc):
    if rc != 0:
        raise MXNetError(_lib().GXRTGetLastError().decode("utf-8", "replace"))


def _shape_args(input_shapes):
    keys = (ctypes.c_char_p * len(input_shapes))(*[k.encode() for k in input_shapes])
    indptr, data = [0], []
    for s in input_shapes.values():
        data.extend(int(d) for d in s)
        indptr.append(len(data))
    return keys, (ctypes.c_uint32 * len(indptr))(*indptr), (ctypes.c_uint32 * max(len(data), 1))(*data)


class _DevicePredictor:
    """dev_type='gpu': the same interface on the Python Executor."""

    def __init__(self, symbol_json, param_bytes, input_shapes, dev_id, output_keys):
        from . import symbol as sym_mod, ndarray as nd
        from .context import gpu
        from .ndarray.utils import load_frombuffer
        net = sym_mod.load_json(symbol_json)
        if output_keys:
            internals = net.get_internals()
            net = sym_mod.Group([internals[k if k.endswith("_output") else k + "_output"] for k in output_keys])
        self._sym, self._ctx = net, gpu(dev_id)
        self._params = {k.split(":", 1)[-1]: v for k, v in load_frombuffer(param_bytes).items()} if param_bytes else {}
        self._nd = nd
        self._bind(dict(input_shapes))

    def _bind(self, input_shapes):
        self._shapes = input_shapes
        arg_shapes, _, aux_shapes = self._sym.infer_shape(**input_shapes)
        args = {}
        for name, shp in zip(self._sym.list_arguments(), arg_shapes):
            args[name] = self._params[name].as_in_context(self._ctx) if name in self._params and name not in input_shapes else self._nd.zeros(shp, ctx=self._ctx)
        aux = {name: self._params[name].as_in_context(self._ctx) for name in self._sym.list_auxiliary_states()}
        self._exec = self._sym.bind(self._ctx, args, aux_states=aux, grad_req="null")

    def forward(self, **kwargs):
        for k, v in kwargs.items():
            self._exec.arg_dict[k][:] = self._nd.array(np.asarray(v, dtype=np.float32), ctx=self._ctx)
        self._out = self._exec.forward(is_train=False)

    def get_output(self, index):
        return self._out[index].asnumpy()

    def reshape(self, input_shapes):
        self._bind(dict(input_shapes))


class Predictor:
    """``Predictor(symbol_json, param_bytes, input_shapes, dev_type='cpu', dev_id=0, output_keys=None)``."""

    def __init__(self, symbol_json, param_raw_bytes, input_shapes, dev_type="cpu", dev_id=0, output_keys=None):
        if hasattr(symbol_json, "tojson"):
            symbol_json = symbol_json.tojson()
        self._dev = None
        self._h = None
        if dev_type in ("gpu", 2):
            self._dev = _DevicePredictor(symbol_json, param_raw_bytes, input_shapes, dev_id, output_keys)
            return
        lib = _lib()
        keys, indptr, data = _shape_args(input_shapes)
        h = ctypes.c_void_p()
        blob = bytes(param_raw_bytes or b"")
        if output_keys:
            ok = (ctypes.c_char_p * len(output_keys))(*[k.encode() for k in output_keys])
            _ck(lib.GXPredCreatePartialOut(symbol_json.encode(), blob, len(blob), 1, dev_id, len(input_shapes), keys, indptr, data, len(output_keys), ok, ctypes.byref(h)))
        else:
            _ck(lib.GXPredCreate(symbol_json.encode(), blob, len(blob), 1, dev_id, len(input_shapes), keys, indptr, data, ctypes.byref(h)))
        self._h = h

    @classmethod
    def _wrap(cls, handle):
        self = cls.__new__(cls)
        self._dev, self._h = None, handle
        return self

    def __del__(self):
        if getattr(self, "_h", None) is not None and _LIB is not None:
            _LIB.GXPredFree(self._h)
            self._h = None

    def forward(self, **kwargs):
        """Set the named inputs (numpy arrays) and run the whole graph."""
        if self._dev:
            return self._dev.forward(**kwargs)
        lib = _lib()
        for k, v in kwargs.items():
            v = np.ascontiguousarray(v, dtype=np.float32)
            _ck(lib.GXPredSetInput(self._h, k.encode(), v.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), v.size))
        _ck(lib.GXPredForward(self._h))

    def partial_forward(self, step):
        """Run operator ``step`` only; returns how many are left (MXPredPartialForward)."""
        left = ctypes.c_int()
        _ck(_lib().GXPredPartialForward(self._h, int(step), ctypes.byref(left)))
        return left.value

    @property
    def num_outputs(self):
        n = ctypes.c_uint32()
        _ck(_lib().GXPredGetNumOutputs(self._h, ctypes.byref(n)))
        return n.value

    def get_output_shape(self, index):
        pdata, ndim = ctypes.POINTER(ctypes.c_uint32)(), ctypes.c_uint32()
        _ck(_lib().GXPredGetOutputShape(self._h, int(index), ctypes.byref(pdata), ctypes.byref(ndim)))
        return tuple(pdata[i] for i in range(ndim.value))

    def get_output(self, index):
        if self._dev:
            return self._dev.get_output(index)
        out = np.empty(self.get_output_shape(index), dtype=np.float32)
        _ck(_lib().GXPredGetOutput(self._h, int(index), out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), out.size))
        return out

    def reshape(self, input_shapes):
        """Re-plan for other input shapes; parameters are shared with the old plan, which this call replaces."""
        if self._dev:
            return self._dev.reshape(input_shapes)
        keys, indptr, data = _shape_args(input_shapes)
        h = ctypes.c_void_p()
        _ck(_lib().GXPredReshape(len(input_shapes), keys, indptr, data, self._h, ctypes.byref(h)))
        _lib().GXPredFree(self._h)
        self._h = h

    def plan(self):
        """(activation arena bytes, number of operators that run) of the native memory plan."""
        a, n = ctypes.c_uint64(), ctypes.c_uint32()
        _ck(_lib().GXPredGetPlan(self._h, ctypes.byref(a), ctypes.byref(n)))
        return a.value, n.value


def create_multi_thread(symbol_json, param_raw_bytes, input_shapes, num_threads):
    """``num_threads`` predictors over one copy of the graph and parameters, one per serving thread (MXPredCreateMultiThread)."""
    if hasattr(symbol_json, "tojson"):
        symbol_json = symbol_json.tojson()
    keys, indptr, data = _shape_args(input_shapes)
    hs = (ctypes.c_void_p * num_threads)()
    blob = bytes(param_raw_bytes or b"")
    _ck(_lib().GXPredCreateMultiThread(symbol_json.encode(), blob, len(blob), 1, 0, len(input_shapes), keys, indptr, data, num_threads, hs))
    return [Predictor._wrap(ctypes.c_void_p(h)) for h in hs]


def load_ndarray_file(nd_bytes):
    """Parse an NDArray-list file natively: dict name -> numpy array (list when the file has no names).  MXNDListCreate / Get / Free."""
    lib = _lib()
    h, n = ctypes.c_void_p(), ctypes.c_uint32()
    _ck(lib.GXNDListCreate(bytes(nd_bytes), len(nd_bytes), ctypes.byref(h), ctypes.byref(n)))
    try:
        names, arrays = [], []
        for i in range(n.value):
            key, pdata, pshape, ndim = ctypes.c_char_p(), ctypes.POINTER(ctypes.c_float)(), ctypes.POINTER(ctypes.c_uint32)(), ctypes.c_uint32()
            _ck(lib.GXNDListGet(h, i, ctypes.byref(key), ctypes.byref(pdata), ctypes.byref(pshape), ctypes.byref(ndim)))
            shape = tuple(pshape[j] for j in range(ndim.value))
            size = int(np.prod(shape)) if shape else 0
            arrays.append(np.ctypeslib.as_array(pdata, shape=(size,)).reshape(shape).copy() if size else np.zeros(shape, np.float32))
            names.append(key.value.decode(errors="replace") if key.value else "")
        return dict(zip(names, arrays)) if any(names) else arrays
    finally:
        lib.GXNDListFree(h)
