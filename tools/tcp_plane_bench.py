"""Push+pull round time of the native TCP HiPS plane on localhost — BASELINE config 1 (1 local PS + 2 workers, CPU) and the 2x2 two-tier
topology — for the ten keys of the demo CNN (178 762 fp32 parameters, 715 KB), SGD on the (global) server.

    python tools/tcp_plane_bench.py [--rounds 200] [--two-tier] [--workers 2]

Every worker pushes all keys (priority -idx) and pulls them back, then waits: one "round" = what a training step pays for synchronisation on
this plane.  Reports the median / p90 round time of worker 0 and the payload rate.  Role processes are the usual ones (``import geomx_b200``)."""
import argparse
import json
import os
import socket
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(16, 1, 5, 5), (16,), (32, 16, 5, 5), (32,), (256, 512), (256,), (128, 256), (128,), (10, 128), (10,)]
if os.environ.get("BENCH_SHAPES"):            # e.g. BENCH_SHAPES="10" (one tiny key: pure message latency) or "131072;131072"
    SHAPES = [tuple(int(d) for d in s.split("x")) for s in os.environ["BENCH_SHAPES"].split(";")]


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def worker_main():
    sys.path.insert(0, ROOT)
    import numpy as np
    import geomx_b200 as mx
    rounds = int(os.environ["BENCH_ROUNDS"])
    kv = mx.kv.create("dist_sync")
    master = kv.is_master_worker
    if master or (os.environ.get("BENCH_STANDALONE") == "1" and kv.rank == 0):
        kv.set_optimizer(mx.optimizer.SGD(learning_rate=0.01))
    time.sleep(0.5)
    params = [mx.nd.array(np.full(s, 0.1, dtype=np.float32)) for s in SHAPES]
    grads = [mx.nd.array(np.full(s, 0.01, dtype=np.float32)) for s in SHAPES]
    for i, p in enumerate(params):
        kv.init(i, p)
        if not master:
            kv.pull(i, p)
    mx.nd.waitall()
    if master:
        kv.close(); return
    if os.environ.get("BENCH_SPLIT") == "1":          # push and pull timed separately (each followed by a wait): where a round's time goes
        tp, tq = [], []
        for r in range(rounds + 5):
            t0 = time.perf_counter()
            for i, g in enumerate(grads):
                kv.push(i, g, priority=-i)
            mx.nd.waitall(); t1 = time.perf_counter()
            for i, p in enumerate(params):
                kv.pull(i, p, priority=-i)
            mx.nd.waitall(); t2 = time.perf_counter()
            if r >= 5:
                tp.append((t1 - t0) * 1e3); tq.append((t2 - t1) * 1e3)
        if os.environ.get("BENCH_REPORT") == "1":
            tp.sort(); tq.sort()
            print("RESULT " + json.dumps({"rounds": rounds, "push_median_ms": round(tp[len(tp) // 2], 3), "pull_median_ms": round(tq[len(tq) // 2], 3),
                                          "payload_bytes_each_way": sum(int(np.prod(s)) for s in SHAPES) * 4}), flush=True)
        kv.close(); return
    times = []
    for r in range(rounds + 10):
        t0 = time.perf_counter()
        for i, (p, g) in enumerate(zip(params, grads)):
            kv.push(i, g, priority=-i)
            kv.pull(i, p, priority=-i)
        mx.nd.waitall()
        if r >= 10:
            times.append((time.perf_counter() - t0) * 1e3)
    if os.environ.get("BENCH_REPORT") == "1":
        times.sort()
        nbytes = sum(int(np.prod(s)) for s in SHAPES) * 4
        med = times[len(times) // 2]
        print("RESULT " + json.dumps({"rounds": rounds, "median_ms": round(med, 3), "p90_ms": round(times[int(len(times) * 0.9)], 3), "min_ms": round(times[0], 3),
                                      "payload_bytes_each_way": nbytes, "MB_per_s_each_way": round(nbytes / med / 1e3, 1)}), flush=True)
    kv.close()


def spawn(env, worker=False):
    e = dict(os.environ, OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update({k: str(v) for k, v in env.items()})
    cmd = [sys.executable, os.path.abspath(__file__), "--as-worker"] if worker else [sys.executable, "-c", "import sys; sys.path.insert(0, %r); import geomx_b200" % ROOT]
    return subprocess.Popen(cmd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=200)
    ap.add_argument("--workers", type=int, default=2, help="workers per party")
    ap.add_argument("--two-tier", action="store_true", help="2 parties + global server (the 12-process demo topology)")
    ap.add_argument("--as-worker", action="store_true")
    a = ap.parse_args()
    if a.as_worker:
        return worker_main()
    procs = []
    common = {"BENCH_ROUNDS": a.rounds}
    if not a.two_tier:
        base = dict(common, DMLC_PS_ROOT_URI="127.0.0.1", DMLC_PS_ROOT_PORT=free_port(), DMLC_NUM_SERVER=1, DMLC_NUM_WORKER=a.workers, DMLC_NUM_ALL_WORKER=a.workers,
                    BENCH_STANDALONE=1)
        procs += [spawn(dict(base, DMLC_ROLE="scheduler")), spawn(dict(base, DMLC_ROLE="server"))]
        procs += [spawn(dict(base, DMLC_ROLE="worker", BENCH_REPORT=int(i == 0)), worker=True) for i in range(a.workers)]
    else:
        parties, allw = 2, 2 * a.workers
        g = dict(common, DMLC_PS_GLOBAL_ROOT_URI="127.0.0.1", DMLC_PS_GLOBAL_ROOT_PORT=free_port(), DMLC_NUM_GLOBAL_SERVER=1, DMLC_NUM_GLOBAL_WORKER=parties)
        procs.append(spawn(dict(g, DMLC_ROLE_GLOBAL="global_scheduler")))
        central = dict(common, DMLC_PS_ROOT_URI="127.0.0.1", DMLC_PS_ROOT_PORT=free_port(), DMLC_NUM_SERVER=1, DMLC_NUM_WORKER=1, DMLC_NUM_ALL_WORKER=allw)
        procs.append(spawn(dict(g, **central, DMLC_ROLE_GLOBAL="global_server", DMLC_ROLE="server", DMLC_ENABLE_CENTRAL_WORKER=0)))
        procs.append(spawn(dict(central, DMLC_ROLE="scheduler")))
        procs.append(spawn(dict(central, DMLC_ROLE="worker", DMLC_ROLE_MASTER_WORKER=1), worker=True))
        first = True
        for _ in range(parties):
            party = dict(common, DMLC_PS_ROOT_URI="127.0.0.1", DMLC_PS_ROOT_PORT=free_port(), DMLC_NUM_SERVER=1, DMLC_NUM_WORKER=a.workers, DMLC_NUM_ALL_WORKER=allw)
            procs.append(spawn(dict(party, DMLC_ROLE="scheduler")))
            procs.append(spawn(dict(g, **party, DMLC_ROLE="server")))
            for _ in range(a.workers):
                procs.append(spawn(dict(party, DMLC_ROLE="worker", BENCH_REPORT=int(first)), worker=True)); first = False
    deadline = time.time() + 600
    while any(p.poll() is None for p in procs) and time.time() < deadline:
        if any(p.poll() not in (None, 0) for p in procs):
            break
        time.sleep(0.2)
    ok = all(p.poll() == 0 for p in procs)
    for p in procs:
        if p.poll() is None:
            p.kill()
    outs = [p.communicate()[0] for p in procs]
    for o in outs:
        for line in (o or "").splitlines():
            if line.startswith("RESULT "):
                d = json.loads(line[7:]); d["topology"] = "2 parties x %d workers + global server" % a.workers if a.two_tier else "1 server + %d workers" % a.workers
                print(json.dumps(d))
    if not ok:
        print("\n".join((o or "")[-800:] for o in outs), file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
