"""Plain C ABI of the native runtime (csrc/runtime/c_api_runtime.cc) driven through ctypes only: host NDArray handles + byte-exact .params
save/load (read back by the Python frontend), dependency engine with C callbacks, profiler, pooled host storage."""
import ctypes
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_runtime_c_api(tmp_path):
    lib = ctypes.CDLL(glob.glob(os.path.join(ROOT, "geomx_b200", "lib", "_C*.so"))[0])
    lib.GXRTGetLastError.restype = ctypes.c_char_p

    def ck(rc):
        assert rc == 0, lib.GXRTGetLastError().decode()
    shape = (ctypes.c_uint32 * 2)(3, 4)
    h = ctypes.c_void_p()
    ck(lib.GXNDArrayCreate(shape, 2, 0, ctypes.byref(h)))
    data = (ctypes.c_float * 12)(*[float(i) for i in range(12)])
    ck(lib.GXNDArraySyncCopyFromCPU(h, data, ctypes.c_size_t(12)))
    assert lib.GXNDArraySyncCopyFromCPU(h, data, ctypes.c_size_t(11)) == -1 and b"size" in lib.GXRTGetLastError()
    fname = str(tmp_path / "capi.params").encode()
    ck(lib.GXNDArraySave(fname, 1, (ctypes.c_void_p * 1)(h), (ctypes.c_char_p * 1)(b"arg:w")))
    n, hn = ctypes.c_uint32(), ctypes.c_uint32()
    outh = ctypes.POINTER(ctypes.c_void_p)(); outn = ctypes.POINTER(ctypes.c_char_p)()
    ck(lib.GXNDArrayLoad(fname, ctypes.byref(n), ctypes.byref(outh), ctypes.byref(hn), ctypes.byref(outn)))
    nd, shp, dt = ctypes.c_uint32(), ctypes.POINTER(ctypes.c_uint32)(), ctypes.c_int()
    ck(lib.GXNDArrayGetShape(ctypes.c_void_p(outh[0]), ctypes.byref(nd), ctypes.byref(shp)))
    ck(lib.GXNDArrayGetDType(ctypes.c_void_p(outh[0]), ctypes.byref(dt)))
    back = (ctypes.c_float * 12)()
    ck(lib.GXNDArraySyncCopyToCPU(ctypes.c_void_p(outh[0]), back, ctypes.c_size_t(12)))
    assert (n.value, hn.value, outn[0], nd.value, shp[0], shp[1], dt.value) == (1, 1, b"arg:w", 2, 3, 4, 0) and list(back) == [float(i) for i in range(12)]
    import geomx_b200 as mx
    d = mx.nd.load(fname.decode())                       # the Python frontend reads what the C API wrote
    assert tuple(d["arg:w"].shape) == (3, 4) and float(d["arg:w"].asnumpy()[2, 3]) == 11.0
    ck(lib.GXNDArrayFree(h)); ck(lib.GXNDArrayFree(ctypes.c_void_p(outh[0])))
    # engine: five writers of one variable run in push order
    FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p)
    log = []
    cbs = [FN(lambda a, i=i: log.append(i)) for i in range(5)]
    v = ctypes.c_int(); ck(lib.GXEngineNewVariable(ctypes.byref(v)))
    mv = (ctypes.c_int * 1)(v.value)
    for cb in cbs:
        ck(lib.GXEnginePushAsync(cb, None, None, 0, mv, 1, 0, b"w"))
    ck(lib.GXEngineWaitForVar(v.value)); ck(lib.GXEngineWaitAll())
    assert log == [0, 1, 2, 3, 4]
    # ... also on a device's copy pool (GXEnginePushAsyncEx: exec device + FnProperty), then the variable is deleted
    cb2 = FN(lambda a: log.append("copy"))
    ck(lib.GXEnginePushAsyncEx(cb2, None, None, 0, mv, 1, 0, b"c", 3, 1))
    ck(lib.GXEngineDeleteVariable(v.value)); ck(lib.GXEngineWaitAll())
    assert log[-1] == "copy"
    # profiler
    prof = str(tmp_path / "prof.json").encode()
    ck(lib.GXSetProfilerConfig(1, (ctypes.c_char_p * 1)(b"filename"), (ctypes.c_char_p * 1)(prof)))
    ck(lib.GXSetProfilerState(1)); ck(lib.GXProfileSetMarker(b"c_api_marker", b"test")); ck(lib.GXDumpProfile(1)); ck(lib.GXSetProfilerState(0))
    assert "c_api_marker" in open(prof.decode()).read()
    p = ctypes.c_void_p(); ck(lib.GXStorageAlloc(ctypes.c_size_t(1 << 20), ctypes.byref(p))); assert p.value; ck(lib.GXStorageFree(p))
