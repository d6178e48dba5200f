"""Worker body for the multi-process HiPS tests (CPU).  Prints RESULT lines parsed by the test."""
import faulthandler
import json
import os
import signal
import sys
import time

faulthandler.register(signal.SIGUSR1, all_threads=True)      # the test harness asks for a traceback before it kills a hung run

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GEOMX_SYNTHETIC_SIZE", "256")
import numpy as np  # noqa: E402

import geomx_b200 as mx  # noqa: E402

mode = os.environ.get("TEST_MODE", "sgd")
steps = int(os.environ.get("TEST_STEPS", "3"))
gid = int(os.environ.get("TEST_WORKER_GID", os.environ.get("GEOMX_WORKER_INDEX", "0")))       # global worker index (for deterministic gradients)
shapes = [(4, 5), (7,), (300,), (3, 2)] if mode != "big" else [(4, 5), (2500,)]

kv = mx.kv.create(os.environ.get("TEST_KV", "dist_sync"))
master = kv.is_master_worker
if master or (os.environ.get("TEST_STANDALONE") == "1" and kv.rank == 0):
    if mode in ("sgd", "big", "p3", "2bit", "async", "fp16", "rowsparse", "recovery") and os.environ.get("TEST_RECOVERED") != "1":
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1, multi_precision=(mode == "fp16")))
    elif mode == "adam_py":
        os.environ["GEOMX_PY_UPDATER"] = "1"
        kv.set_optimizer(mx.optimizer.Adam(learning_rate=0.01))
    if mode == "bsc_async":
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.1))
    if mode in ("bsc", "bsc_async"):
        kv.set_gradient_compression({"type": "bsc", "threshold": 0.1})
if mode == "2bit":      # every worker (master included, it configures the global servers), as the reference scripts do
    kv.set_gradient_compression({"type": "2bit", "threshold": 0.5})
time.sleep(0.5)
dt = "float16" if mode == "fp16" else "float32"
params = [mx.nd.array(np.full(s, 1.0 + i, dtype=np.float32)).astype(dt) for i, s in enumerate(shapes)]
for i, p in enumerate(params):
    kv.init(i, p)
    if master:
        continue
    kv.pull(i, p)
mx.nd.waitall()
if master:
    print("RESULT master done", flush=True)
    kv.close()
    sys.exit(0)
out = {"rank": kv.rank, "num_workers": kv.num_workers, "num_all_workers": kv.num_all_workers, "vals": []}
if mode == "recovery":
    # elastic recovery (van.cc:176-192): gid 1 crashes after round 1; the harness starts a replacement (TEST_RECOVERED=1) that takes over the
    # dead id, skips barriers / init, pulls the current parameters and takes part in round 2.
    recovered = os.environ.get("TEST_RECOVERED") == "1"
    w = mx.nd.array(np.full((6,), 1.0, dtype=np.float32))
    kv.init(5, w)
    kv.pull(5, w); mx.nd.waitall()
    vals = [float(w.asnumpy()[0])]
    for rnd in range(2):
        if recovered and rnd == 0:
            continue                               # round 1 happened before this process existed
        kv.push(5, mx.nd.array(np.full((6,), 0.5 * (gid + 1), dtype=np.float32)))
        kv.pull(5, w); mx.nd.waitall()
        vals.append(float(w.asnumpy()[0]))
        if rnd == 0 and gid == 1 and not recovered:
            print("RESULT " + json.dumps({"gid": gid, "crashed": True, "vals": vals}), flush=True)
            os._exit(17)                           # crash
    print("RESULT " + json.dumps({"gid": gid, "recovered": recovered, "is_recovery": bool(kv.is_recovery), "vals": vals}), flush=True)
    kv.close()
    sys.exit(0)
if mode == "heartbeat":
    # failure detection (postoffice.cc GetDeadNodes, kvstore_dist.h:225-234): a worker's view covers the scheduler it heart-beats with.
    # Signal readiness through a file, then wait until the (killed) scheduler is reported dead.
    open(os.environ["TEST_READY_FILE"] + str(kv.rank), "w").write("ready")
    alive_first = kv.get_num_dead_node(1, timeout=2)
    dead = 0
    for _ in range(60):
        dead = kv.get_num_dead_node(1, timeout=2)
        if dead:
            break
        time.sleep(0.5)
    print("RESULT " + json.dumps({"rank": kv.rank, "alive_first": alive_first, "dead": dead}), flush=True)
    os._exit(0)                                  # no orderly shutdown is possible without a scheduler
if mode == "rowsparse":
    # embedding-style key: every worker pushes two rows (one shared, one private), then pulls a few rows back through the sparse wire
    emb = mx.nd.array(np.zeros((10, 4), dtype=np.float32))
    kv.init(100, emb)
    kv.pull(100, emb); mx.nd.waitall()
    for step in range(steps):
        g = mx.nd.sparse.row_sparse_array((np.full((2, 4), 1.0 + gid, dtype=np.float32), [3, 5 + gid]), shape=(10, 4))
        kv.push(100, g)
        got = mx.nd.sparse.zeros("row_sparse", (10, 4))
        kv.row_sparse_pull(100, out=got, row_ids=mx.nd.array([5, 3, 6, 3], dtype="int64"))
        out["vals"].append({"ids": got.indices.asnumpy().tolist(), "rows": got.data.asnumpy()[:, 0].tolist()})
    print("RESULT " + json.dumps(out), flush=True)
    kv.close()
    sys.exit(0)
prof_path = os.environ.get("TEST_SERVER_PROFILE")
if prof_path and kv.rank == 0:
    # remote server profiling (profiler.py:28-33 -> kSetProfilerParams): the server prefixes the file name with rank<r>_
    mx.profiler.set_config(profile_process="server", filename=prof_path, profile_all=True, aggregate_stats=True)
    mx.profiler.set_state("run", profile_process="server")
for step in range(steps):
    for i, p in enumerate(params):
        if mode == "hfa":
            local = mx.nd.array(np.full(shapes[i], float(step + 1) * (gid + 1), dtype=np.float32))
            kv.push(i, local / kv.num_workers, priority=-i)
        else:
            g = np.full(shapes[i], 0.5 * (gid + 1) * (1 if mode != "2bit" else 2), dtype=np.float32)
            if os.environ.get("TEST_NONUNIFORM") == "1":          # ramp: the largest entries sit at the END of the tensor
                n_el = int(np.prod(shapes[i]))
                g = (g.reshape(-1) * (np.arange(n_el, dtype=np.float32) + 1.0) / n_el).reshape(shapes[i])
            kv.push(i, mx.nd.array(g).astype(dt), priority=-i)
        kv.pull(i, p, priority=-i)
    mx.nd.waitall()
    out["vals"].append([float(p.astype("float32").asnumpy().reshape(-1)[0]) for p in params])
    out.setdefault("last", [float(p.astype("float32").asnumpy().reshape(-1)[-1]) for p in params])
    out.setdefault("nonzeros", [int(np.count_nonzero(p.astype("float32").asnumpy())) for p in params])
if prof_path and kv.rank == 0:
    mx.profiler.pause(profile_process="server"); mx.profiler.resume(profile_process="server")
    mx.profiler.dump(profile_process="server")
if hasattr(getattr(kv, "_kv", None), "ts_stats"):
    out["ts_stats"] = list(kv._kv.ts_stats())
print("RESULT " + json.dumps(out), flush=True)
kv.close()
