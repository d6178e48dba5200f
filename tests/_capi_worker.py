"""A parameter-server worker written against the plain C API only (ctypes over lib/_C*.so): init / push / pull / barrier through GXKVStore*."""
import ctypes
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(glob.glob(os.path.join(ROOT, "geomx_b200", "lib", "_C*.so"))[0])
lib.GXGetLastError.restype = ctypes.c_char_p


def ck(rc):
    if rc != 0:
        raise RuntimeError(lib.GXGetLastError().decode())


h = ctypes.c_void_p()
ck(lib.GXKVStoreCreate(b"dist_sync", ctypes.byref(h)))
rank, nw = ctypes.c_int(), ctypes.c_int()
ck(lib.GXKVStoreGetRank(h, ctypes.byref(rank))); ck(lib.GXKVStoreGetGroupSize(h, ctypes.byref(nw)))
if rank.value == 0:   # optimizer spec (command 7): SGD lr 0.1, executed natively on the server
    ck(lib.GXKVStoreSendCommmandToServers(h, 7, b"name=sgd;lr=0.1;wd=0.0;rescale_grad=1.0;clip_gradient=-1.0;momentum=0.0"))
N = 6
w = (ctypes.c_float * N)(*[1.0] * N)
ck(lib.GXKVStoreInit(h, 3, w, ctypes.c_size_t(N), 0))
vals = []
for step in range(2):
    g = (ctypes.c_float * N)(*[0.5 * (rank.value + 1)] * N)
    hp, hl = ctypes.c_int(), ctypes.c_int()
    ck(lib.GXKVStorePush(h, 3, g, ctypes.c_size_t(N), 0, 0, ctypes.byref(hp)))
    ck(lib.GXKVStorePull(h, 3, w, ctypes.c_size_t(N), 0, 0, ctypes.byref(hl)))
    ck(lib.GXKVStoreWait(h, hl))
    vals.append(float(w[0]))
print("RESULT " + json.dumps({"rank": rank.value, "num_workers": nw.value, "vals": vals}), flush=True)
ck(lib.GXKVStoreFree(h))
