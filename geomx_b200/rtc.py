"""``mx.rtc`` — compile and launch user CUDA kernels at run time (NVRTC → cubin for sm_100a → driver module).

Parity: ``python/mxnet/rtc.py`` (``CudaModule(source, options, exports)``, ``get_kernel(name, signature)``, ``CudaKernel.launch(args, ctx,
grid_dims, block_dims, shared_mem)``) over ``src/common/rtc.cc``.  The signature string uses the reference's syntax
(``"const float *x, float *y, int n"``): pointer parameters take NDArrays (their device address is passed), scalars take Python numbers.
Compilation needs no GPU (the cubin is produced for ``sm_100a``); loading and launching do."""
from __future__ import annotations

import ctypes
import re

import numpy as np

from .base import MXNetError
from .ndarray import NDArray

__all__ = ["CudaModule", "CudaKernel"]

_CTYPES = {"float": ctypes.c_float, "double": ctypes.c_double, "int": ctypes.c_int, "int32_t": ctypes.c_int32, "uint32_t": ctypes.c_uint32,
           "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "long": ctypes.c_long, "char": ctypes.c_char, "int8_t": ctypes.c_int8,
           "uint8_t": ctypes.c_uint8, "bool": ctypes.c_bool, "size_t": ctypes.c_size_t}


def _nvrtc():
    try:
        from cuda.bindings import nvrtc
    except ImportError:  # older cuda-python layout
        from cuda import nvrtc
    return nvrtc


def _driver():
    try:
        from cuda.bindings import driver
    except ImportError:
        from cuda import cuda as driver
    return driver


class CudaModule:
    def __init__(self, source, options=(), exports=()):
        nvrtc = _nvrtc()
        err, prog = nvrtc.nvrtcCreateProgram(source.encode(), b"geomx_rtc.cu", 0, [], [])
        if int(err) != 0:
            raise MXNetError("nvrtcCreateProgram failed: %s" % err)
        opts = [b"--gpu-architecture=sm_100a", b"--std=c++17"] + [o.encode() if isinstance(o, str) else o for o in options]
        err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
        if int(err) != 0:
            _, n = nvrtc.nvrtcGetProgramLogSize(prog)
            log = b" " * n
            nvrtc.nvrtcGetProgramLog(prog, log)
            raise MXNetError("NVRTC compilation failed:\n" + log.decode(errors="replace"))
        _, size = nvrtc.nvrtcGetCUBINSize(prog)
        self.cubin = b" " * size
        err, = nvrtc.nvrtcGetCUBIN(prog, self.cubin)
        if int(err) != 0:
            raise MXNetError("nvrtcGetCUBIN failed: %s" % err)
        nvrtc.nvrtcDestroyProgram(prog)
        self.exports, self._module = tuple(exports), None

    def _load(self):
        if self._module is None:
            import torch
            if not torch.cuda.is_available():
                raise MXNetError("mx.rtc: launching needs a CUDA device (compilation does not)")
            torch.cuda.init(); torch.zeros(1, device="cuda")          # make sure the primary context is current
            drv = _driver()
            err, mod = drv.cuModuleLoadData(np.frombuffer(self.cubin, dtype=np.uint8).ctypes.data)
            if int(err) != 0:
                raise MXNetError("cuModuleLoadData failed: %s" % err)
            self._module = mod
        return self._module

    def get_kernel(self, name, signature):
        """``signature``: C parameter list, e.g. ``"const float *x, float *y, float alpha, int n"``."""
        params = []
        for part in [p.strip() for p in signature.split(",") if p.strip()]:
            m = re.match(r"^(const\s+)?([\w:]+(?:\s+[\w:]+)*?)\s*(\*?)\s*(\w+)?$", part)
            if not m:
                raise MXNetError("cannot parse kernel parameter %r" % part)
            is_ptr, typ = m.group(3) == "*", m.group(2).strip()
            if not is_ptr and typ not in _CTYPES:
                raise MXNetError("unsupported scalar type %r in kernel signature" % typ)
            params.append((is_ptr, typ))
        return CudaKernel(self, name, params)


class CudaKernel:
    def __init__(self, module, name, params):
        self._mod, self._name, self._params, self._fn = module, name, params, None

    def launch(self, args, ctx, grid_dims, block_dims, shared_mem=0):
        import torch
        drv = _driver()
        if self._fn is None:
            err, fn = drv.cuModuleGetFunction(self._mod._load(), self._name.encode())
            if int(err) != 0:
                raise MXNetError("kernel %s not found in module: %s" % (self._name, err))
            self._fn = fn
        if len(args) != len(self._params):
            raise MXNetError("kernel %s expects %d arguments, got %d" % (self._name, len(self._params), len(args)))
        holders = []
        for a, (is_ptr, typ) in zip(args, self._params):
            if is_ptr:
                if not isinstance(a, NDArray) or not a._t.is_cuda:
                    raise MXNetError("pointer arguments must be NDArrays on a GPU context")
                holders.append(ctypes.c_void_p(a._t.data_ptr()))
            else:
                holders.append(_CTYPES[typ](a))
        argv = (ctypes.c_void_p * len(holders))(*[ctypes.cast(ctypes.pointer(h), ctypes.c_void_p) for h in holders])
        g, b = tuple(grid_dims) + (1,) * (3 - len(grid_dims)), tuple(block_dims) + (1,) * (3 - len(block_dims))
        stream = torch.cuda.current_stream().cuda_stream
        err, = drv.cuLaunchKernel(self._fn, g[0], g[1], g[2], b[0], b[1], b[2], int(shared_mem), stream, ctypes.addressof(argv), 0)
        if int(err) != 0:
            raise MXNetError("cuLaunchKernel failed: %s" % err)
