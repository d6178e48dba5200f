"""ctypes helpers over the flat C ABI (`geomx_b200/lib/_C*.so`) shared by the C API tests: nothing here goes through the Python front end."""
import ctypes
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None

u32, vp, cp = ctypes.c_uint32, ctypes.c_void_p, ctypes.c_char_p


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(glob.glob(os.path.join(ROOT, "geomx_b200", "lib", "_C*.so"))[0])
        _lib.GXRTGetLastError.restype = cp
    return _lib


def err():
    return lib().GXRTGetLastError().decode(errors="replace")


def ck(rc):
    assert rc == 0, err()


def strs(items):
    return (cp * max(len(items), 1))(*[s.encode() if isinstance(s, str) else s for s in items])


def handles(items):
    return (vp * max(len(items), 1))(*[h if isinstance(h, vp) else vp(h) for h in items])


def str_list(fn, *args):
    n, out = u32(), ctypes.POINTER(cp)()
    ck(fn(*args, ctypes.byref(n), ctypes.byref(out)))
    return [out[i].decode(errors="replace") for i in range(n.value)]


# ---------------------------------------------------------------------------------------------------------------- NDArray
def nd_create(arr):
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    h = vp()
    shape = (u32 * max(arr.ndim, 1))(*arr.shape)
    ck(lib().GXNDArrayCreate(shape, arr.ndim, 0, ctypes.byref(h)))
    nd_set(h, arr)
    return h


def nd_set(h, arr):
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    ck(lib().GXNDArraySyncCopyFromCPU(h, arr.ctypes.data_as(vp), ctypes.c_size_t(arr.size)))


def nd_shape(h):
    nd, shp = u32(), ctypes.POINTER(u32)()
    ck(lib().GXNDArrayGetShape(h, ctypes.byref(nd), ctypes.byref(shp)))
    return tuple(shp[i] for i in range(nd.value))


def nd_get(h):
    shape = nd_shape(h)
    out = np.empty(shape, dtype=np.float32)
    ck(lib().GXNDArraySyncCopyToCPU(h, out.ctypes.data_as(vp), ctypes.c_size_t(out.size)))
    return out


def nd_free(h):
    ck(lib().GXNDArrayFree(h))


# ---------------------------------------------------------------------------------------------------------------- Symbol
def var(name):
    h = vp()
    ck(lib().GXSymbolCreateVariable(name.encode(), ctypes.byref(h)))
    return h


def op(opname, name, inputs=(), kwinputs=None, **attrs):
    """Atomic symbol + Compose: positional ``inputs`` or keyword ``kwinputs``; attributes are stringified like the front ends do."""
    h = vp()
    keys, vals = list(attrs.keys()), [str(v) for v in attrs.values()]
    ck(lib().GXSymbolCreateAtomicSymbolByName(opname.encode(), len(keys), strs(keys), strs(vals), ctypes.byref(h)))
    if kwinputs:
        ck(lib().GXSymbolCompose(h, name.encode() if name else None, len(kwinputs), strs(list(kwinputs.keys())), handles(list(kwinputs.values()))))
    else:
        ck(lib().GXSymbolCompose(h, name.encode() if name else None, len(inputs), None, handles(list(inputs))))
    return h


def sym_json(h):
    out = cp()
    ck(lib().GXSymbolSaveToJSON(h, ctypes.byref(out)))
    return out.value.decode(errors="replace")


def sym_from_json(js):
    h = vp()
    ck(lib().GXSymbolCreateFromJSON(js.encode(), ctypes.byref(h)))
    return h


def list_arguments(h):
    return str_list(lib().GXSymbolListArguments, h)


def list_outputs(h):
    return str_list(lib().GXSymbolListOutputs, h)


def list_aux(h):
    return str_list(lib().GXSymbolListAuxiliaryStates, h)


def infer_shape(h, partial=False, **shapes):
    keys = list(shapes.keys())
    ind, data = [0], []
    for k in keys:
        data += list(shapes[k]); ind.append(len(data))
    outs = []
    args = [h, len(keys), strs(keys), (u32 * len(ind))(*ind), (u32 * max(len(data), 1))(*data)]
    refs = []
    for _ in range(3):
        n, nd, d = u32(), ctypes.POINTER(u32)(), ctypes.POINTER(ctypes.POINTER(u32))()
        refs.append((n, nd, d)); args += [ctypes.byref(n), ctypes.byref(nd), ctypes.byref(d)]
    complete = ctypes.c_int()
    fn = lib().GXSymbolInferShapePartial if partial else lib().GXSymbolInferShape
    rc = fn(*args, ctypes.byref(complete))
    if rc != 0:
        raise RuntimeError(err())
    for n, nd, d in refs:
        outs.append([tuple(d[i][j] for j in range(nd[i])) for i in range(n.value)])
    return outs[0], outs[1], outs[2], bool(complete.value)


# ---------------------------------------------------------------------------------------------------------------- Executor
def simple_bind(sym, shapes, grad_req="write", no_grad=()):
    keys = list(shapes.keys())
    ind, data = [0], []
    for k in keys:
        data += list(shapes[k]); ind.append(len(data))
    ex, na, nx = vp(), u32(), u32()
    a, g, x = ctypes.POINTER(vp)(), ctypes.POINTER(vp)(), ctypes.POINTER(vp)()
    ck(lib().GXExecutorSimpleBind(sym, len(keys), strs(keys), (u32 * len(ind))(*ind), (u32 * len(data))(*data), grad_req.encode(), len(no_grad), strs(list(no_grad)),
                                  ctypes.byref(ex), ctypes.byref(na), ctypes.byref(a), ctypes.byref(g), ctypes.byref(nx), ctypes.byref(x)))
    names, auxn = list_arguments(sym), list_aux(sym)
    args = {names[i]: vp(a[i]) for i in range(na.value)}
    grads = {names[i]: vp(g[i]) for i in range(na.value) if g[i]}
    aux = {auxn[i]: vp(x[i]) for i in range(nx.value)}
    return ex, args, grads, aux


def exec_outputs(ex):
    n, out = u32(), ctypes.POINTER(vp)()
    ck(lib().GXExecutorOutputs(ex, ctypes.byref(n), ctypes.byref(out)))
    return [vp(out[i]) for i in range(n.value)]


def forward(ex, is_train):
    ck(lib().GXExecutorForward(ex, int(is_train)))
    return [nd_get(h) for h in exec_outputs(ex)]


def backward(ex, head_grads=()):
    ck(lib().GXExecutorBackward(ex, len(head_grads), handles(list(head_grads)) if head_grads else None))


# ---------------------------------------------------------------------------------------------------------------- imperative / autograd
def invoke(opname, inputs, **attrs):
    keys, vals = list(attrs.keys()), [str(v) for v in attrs.values()]
    n, out = ctypes.c_int(0), ctypes.POINTER(vp)()
    rc = lib().GXImperativeInvokeByName(opname.encode(), len(inputs), handles(list(inputs)), ctypes.byref(n), ctypes.byref(out), len(keys), strs(keys), strs(vals))
    if rc != 0:
        raise RuntimeError(err())
    return vp(out[0])


def mark_variables(variables, grads, req=1):
    ck(lib().GXAutogradMarkVariables(len(variables), handles(variables), (u32 * len(variables))(*[req] * len(variables)), handles(grads)))


class record:
    def __init__(self, train=True):
        self.train = train

    def __enter__(self):
        self.p1, self.p2 = ctypes.c_int(), ctypes.c_int()
        ck(lib().GXAutogradSetIsRecording(1, ctypes.byref(self.p1))); ck(lib().GXAutogradSetIsTraining(int(self.train), ctypes.byref(self.p2)))

    def __exit__(self, *a):
        ck(lib().GXAutogradSetIsRecording(self.p1.value, None)); ck(lib().GXAutogradSetIsTraining(self.p2.value, None))


def ag_backward(outputs, ograds=None, retain=False):
    rc = lib().GXAutogradBackward(len(outputs), handles(outputs), handles(ograds) if ograds else None, int(retain))
    if rc != 0:
        raise RuntimeError(err())
