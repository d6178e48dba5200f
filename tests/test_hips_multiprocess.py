"""Multi-process HiPS on one box over the native TCP transport (CPU; the reference fakes multi-node the same way:
3rdparty/ps-lite/tests/local.sh:17-35 and scripts/cpu/run_vanilla_hips.sh).  Covers BASELINE config 1 (1 local PS + 2 workers) and the
full two-tier topology (global scheduler + global server + master worker + central scheduler + 2 x (scheduler, server, 2 workers))."""
import json
import os
import socket
import subprocess
import sys

import pytest

from geomx_b200 import runtime

pytestmark = pytest.mark.skipif(not runtime.available(), reason="native runtime not built")
HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "_hips_worker.py")
BOOT = "import sys; sys.path.insert(0, %r); import geomx_b200" % os.path.dirname(HERE)


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def spawn(env, worker=False, extra=None):
    e = dict(os.environ); e.update({k: str(v) for k, v in env.items()}); e.update(extra or {})
    e.pop("RANK", None); e.pop("WORLD_SIZE", None)
    cmd = [sys.executable, WORKER] if worker else [sys.executable, "-c", BOOT]
    return subprocess.Popen(cmd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def collect(procs, timeout=120):
    """Wait for every process.  Fails as soon as ONE exits non-zero (the rest would wait at a rendezvous forever) or when the deadline
    passes; in both cases the workers are asked for their Python stacks (SIGUSR1) before everything is killed."""
    import signal
    import time
    deadline = time.time() + timeout
    failed = None
    while True:
        states = [p.poll() for p in procs]
        failed = next((p for p, st in zip(procs, states) if st not in (None, 0)), None)
        if failed is not None or all(st == 0 for st in states) or time.time() > deadline:
            break
        time.sleep(0.1)
    if failed is None and all(p.poll() == 0 for p in procs):
        return [p.communicate()[0] for p in procs]
    for q in procs:
        if q.poll() is None and str(q.args[-1]).endswith("_hips_worker.py"):
            q.send_signal(signal.SIGUSR1)
    time.sleep(1.0)
    for q in procs:
        if q.poll() is None:
            q.kill()
    parts = []
    for q in procs:
        try:
            parts.append((q.communicate(timeout=5)[0] or "")[-1500:])
        except Exception:
            parts.append("")
    why = "process exited with code %s" % failed.returncode if failed is not None else "timeout after %d s" % timeout
    raise AssertionError("%s; partial output:\n" % why + "\n-----\n".join(parts))


def results(outs):
    res = []
    for o in outs:
        for line in o.splitlines():
            if line.startswith("RESULT {"):
                res.append(json.loads(line[7:]))
    return res


def launch_single_tier(extra, workers=2, port=None):
    port = port or free_port()
    base = {"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": port, "DMLC_NUM_SERVER": 1, "DMLC_NUM_WORKER": workers, "DMLC_NUM_ALL_WORKER": workers,
            "TEST_STANDALONE": 1}
    procs = [spawn(dict(base, DMLC_ROLE="scheduler"), extra=extra), spawn(dict(base, DMLC_ROLE="server"), extra=extra)]
    ws = [spawn(dict(base, DMLC_ROLE="worker", TEST_WORKER_GID=i), worker=True, extra=extra) for i in range(workers)]
    outs = collect(procs + ws, timeout=int(__import__("os").environ.get("T_TIMEOUT", "120")))
    return results(outs)


def launch_hips(extra, parties=2, wpp=2, global_servers=1):
    gport = free_port()
    g = {"DMLC_PS_GLOBAL_ROOT_URI": "127.0.0.1", "DMLC_PS_GLOBAL_ROOT_PORT": gport, "DMLC_NUM_GLOBAL_SERVER": global_servers, "DMLC_NUM_GLOBAL_WORKER": parties}
    allw = parties * wpp
    procs = [spawn(dict(g, DMLC_ROLE_GLOBAL="global_scheduler"), extra=extra)]
    cport = free_port()
    central = {"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": cport, "DMLC_NUM_SERVER": global_servers, "DMLC_NUM_WORKER": 1, "DMLC_NUM_ALL_WORKER": allw}
    for _ in range(global_servers):
        procs.append(spawn(dict(g, **central, DMLC_ROLE_GLOBAL="global_server", DMLC_ROLE="server", DMLC_ENABLE_CENTRAL_WORKER=0), extra=extra))
    procs.append(spawn(dict(central, DMLC_ROLE="scheduler"), extra=extra))
    ws = [spawn(dict(central, DMLC_ROLE="worker", DMLC_ROLE_MASTER_WORKER=1), worker=True, extra=extra)]
    gid = 0
    for _ in range(parties):
        port = free_port()
        party = {"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": port, "DMLC_NUM_SERVER": 1, "DMLC_NUM_WORKER": wpp, "DMLC_NUM_ALL_WORKER": allw}
        procs.append(spawn(dict(party, DMLC_ROLE="scheduler"), extra=extra))
        procs.append(spawn(dict(g, **party, DMLC_ROLE="server"), extra=extra))
        for _ in range(wpp):
            ws.append(spawn(dict(party, DMLC_ROLE="worker", TEST_WORKER_GID=gid), worker=True, extra=extra)); gid += 1
    outs = collect(procs + ws, timeout=int(os.environ.get("T_TIMEOUT", "180")))
    return results(outs)


def test_single_tier_dist_sync_sgd():
    """BASELINE config 1: 1 local PS + 2 workers.  w_t = w_0 - t * lr * sum_workers(grad)."""
    res = launch_single_tier({"TEST_MODE": "sgd"})
    assert len(res) == 2
    gsum = 0.5 * 1 + 0.5 * 2
    for r in res:
        assert r["num_workers"] == 2
        for t, vals in enumerate(r["vals"]):
            for i, v in enumerate(vals):
                assert abs(v - ((1.0 + i) - 0.1 * gsum * (t + 1))) < 1e-5


def test_foreign_connections_do_not_kill_a_node():
    """a port scan / health check / stale client talking to a HiPS port must cost that connection only.  While a job runs, junk
    is thrown at the scheduler's listening port: a wrong magic, a correct magic with an absurd header, a correct header with a meta block that
    does not parse, and a connection that stalls mid-frame.  The job still finishes with the right numbers."""
    import struct
    import threading
    import time
    port = free_port()
    stop = threading.Event()

    def scanner():
        magic = 0x48695053
        junk = [b"GET / HTTP/1.1\r\nHost: x\r\n\r\n", struct.pack("<III", magic, 0xFFFFFFF0, 7), struct.pack("<III", magic, 8, 0) + b"\xff" * 8,
                struct.pack("<III", magic, 64, 1)]                   # the last one announces a frame and then goes silent
        held = []
        while not stop.is_set():
            for j in junk:
                try:
                    c = socket.create_connection(("127.0.0.1", port), timeout=0.5)
                    c.sendall(j)
                    held.append(c) if j is junk[-1] and len(held) < 4 else c.close()
                except OSError:
                    pass
            time.sleep(0.05)
        for c in held:
            c.close()

    t = threading.Thread(target=scanner, daemon=True); t.start()
    try:
        res = launch_single_tier({"TEST_MODE": "sgd", "PS_RECV_TIMEOUT_MS": "300"}, port=port)
    finally:
        stop.set(); t.join(timeout=5)
    assert len(res) == 2
    for r in res:
        for step, vals in enumerate(r["vals"]):
            assert abs(vals[0] - (1.0 - 0.1 * 1.5 * (step + 1))) < 1e-5


def test_remote_server_profiling(tmp_path):
    """``profile_process='server'``: worker rank 0 configures, runs and dumps the profiler of the server process over the command channel."""
    import json
    res = launch_single_tier({"TEST_MODE": "sgd", "TEST_SERVER_PROFILE": str(tmp_path / "srv.json")})
    assert len(res) == 2
    tr = json.load(open(str(tmp_path / "rank0_srv.json")))
    names = [e["name"] for e in tr["traceEvents"]]
    assert names.count("KVStoreDistServerPush") >= 3 * 4 * 2 - 8 and "KVStoreDistServerPull" in names      # 3 steps x 4 keys x 2 workers (minus those before 'run')


def test_hips_two_tier_fsa():
    """2 parties x 2 workers, global server runs SGD: every worker sees w_0 - t*lr*sum_{4 workers} grad."""
    res = launch_hips({"TEST_MODE": "sgd"})
    assert len(res) == 4
    gsum = 0.5 * (1 + 2 + 3 + 4)
    for r in res:
        assert r["num_all_workers"] == 4 and r["num_workers"] == 2
        for t, vals in enumerate(r["vals"]):
            for i, v in enumerate(vals):
                assert abs(v - ((1.0 + i) - 0.1 * gsum * (t + 1))) < 1e-4, (t, i, v)


def test_fused_inter_tier_pull_is_transparent():
    """GEOMX_FUSED_TIER_PULL (default on): the global server answers a local server's dense push with the post-update value, saving the
    separate pull over the link between parties.  Same arithmetic with the fusion off (the reference's push-ack-then-pull exchange)."""
    gsum = 0.5 * (1 + 2 + 3 + 4)
    for flag in ("1", "0"):
        res = launch_hips({"TEST_MODE": "sgd", "GEOMX_FUSED_TIER_PULL": flag, "TEST_STEPS": "2"})
        assert len(res) == 4
        for r in res:
            for t, vals in enumerate(r["vals"]):
                for i, v in enumerate(vals):
                    assert abs(v - ((1.0 + i) - 0.1 * gsum * (t + 1))) < 1e-4, (flag, t, i, v)


def test_fused_tier_pull_halves_the_round_time_on_a_slow_inter_party_link(tmp_path):
    """``GEOMX_EMULATE_DELAY_MS`` delays every data message on the global plane (one-way latency between parties).  A synchronisation round
    then costs about ONE inter-party round trip with the fused tier pull and about TWO with the reference's push-ack-then-pull exchange."""
    import json
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "tcp_plane_bench.py")
    med = {}
    for flag in ("1", "0"):
        env = dict(os.environ, GEOMX_EMULATE_DELAY_MS="15", GEOMX_FUSED_TIER_PULL=flag, BENCH_SHAPES="64;64")
        for k in ("RANK", "WORLD_SIZE", "DMLC_ROLE"):
            env.pop(k, None)
        r = subprocess.run([sys.executable, tool, "--rounds", "12", "--two-tier"], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
        med[flag] = json.loads(r.stdout.strip().splitlines()[-1])["median_ms"]
    rtt = 2 * 15.0
    assert rtt * 0.95 < med["1"] < rtt * 1.5, med            # one round trip (+ local hops)
    assert 2 * rtt * 0.95 < med["0"] < 2 * rtt * 1.4, med    # two round trips
    assert med["0"] / med["1"] > 1.6, med


def test_hips_multigps_bigarray_python_updater():
    """MultiGPS: 2 global servers, big array partitioned across them; foreign (pickled Python) Adam executed through the Executor."""
    res = launch_hips({"TEST_MODE": "adam_py", "MXNET_KVSTORE_BIGARRAY_BOUND": "100", "TEST_STEPS": "2"}, global_servers=2)
    assert len(res) == 4
    # first Adam step moves every weight by exactly lr (|m/sqrt(v)| = 1), second by ~lr again
    for r in res:
        for i, v in enumerate(r["vals"][0]):
            assert abs(v - ((1.0 + i) - 0.01)) < 1e-4
        for i, v in enumerate(r["vals"][1]):
            assert abs(v - ((1.0 + i) - 0.02)) < 2e-4
        assert abs(r["last"][2] - r["vals"][0][2]) < 1e-6      # both halves of the partitioned 300-element key updated


def test_multigps_keeps_compressed_keys_on_one_server():
    """2 global servers + a key above MXNET_KVSTORE_BIGARRAY_BOUND: dense keys of that size are partitioned across the servers, but a
    Bi-Sparse / 2-bit key must live whole on its hashed server (init, push and pull use one plan) — with a ramp gradient the top-k entries
    all lie in the second half of the tensor, which a server holding only the first slice would silently drop."""
    env = {"MXNET_KVSTORE_BIGARRAY_BOUND": "100", "MXNET_KVSTORE_SIZE_LOWER_BOUND": "100", "TEST_STEPS": "1", "TEST_NONUNIFORM": "1"}
    res = launch_hips(dict(env, TEST_MODE="bsc"), global_servers=2)
    assert len(res) == 4
    for r in res:
        # key 2 (300 elements, threshold 0.1): every party sends its 30 largest entries = indices 270..299; the pull returns their sum
        assert r["nonzeros"][2] == 30, r["nonzeros"]
        assert abs(r["last"][2] - 0.5 * (1 + 2 + 3 + 4)) < 1e-5
    res = launch_hips(dict(env, TEST_MODE="2bit", TEST_STEPS="2"), global_servers=2)
    assert len(res) == 4                      # used to abort the local server with "pull response size mismatch"
    for r in res:
        assert r["nonzeros"][2] == 300


def test_hips_bsc_and_hfa():
    res = launch_hips({"TEST_MODE": "bsc", "MXNET_KVSTORE_SIZE_LOWER_BOUND": "100", "TEST_STEPS": "2"})
    assert len(res) == 4
    # no optimizer on the server: pulls return the aggregated (sparsified above the size bound) gradients
    gsum = 0.5 * (1 + 2 + 3 + 4)
    for r in res:
        assert abs(r["vals"][0][0] - gsum) < 1e-5 and abs(r["vals"][0][1] - gsum) < 1e-5      # small keys: dense path
    res = launch_hips({"TEST_MODE": "hfa", "MXNET_KVSTORE_USE_HFA": "1", "MXNET_KVSTORE_HFA_K1": "1", "MXNET_KVSTORE_HFA_K2": "2", "TEST_STEPS": "2"})
    assert len(res) == 4
    by_rank = sorted(res, key=lambda r: r["vals"][0][0])
    # step 1 (local sync only): each party sees its own average  (1*(g0+1)+1*(g1+1))/2 ; step 2 (global): milestone + mean of party deltas
    firsts = sorted(r["vals"][0][0] for r in res)
    assert firsts == pytest.approx([1.5, 1.5, 3.5, 3.5])
    w0 = 1.0
    expect2 = w0 + ((2 * 1.5 - w0) + (2 * 3.5 - w0)) / 2
    for r in res:
        assert r["vals"][1][0] == pytest.approx(expect2, abs=1e-4)


def test_periodic_server_checkpoint_and_resume(tmp_path):
    """``GEOMX_SERVER_CKPT_PREFIX`` + ``GEOMX_SERVER_CKPT_EVERY``: the server writes its state (weights, optimizer moments, compression
    residuals) every N rounds; a restarted job with ``GEOMX_SERVER_RESUME=1`` continues from it although its scripts call ``kv.init`` again.
    Single tier, then the two-tier topology (checkpoint taken by the global server)."""
    gsum = 0.5 * 1 + 0.5 * 2
    ck = {"TEST_MODE": "sgd", "GEOMX_SERVER_CKPT_PREFIX": str(tmp_path / "ck"), "GEOMX_SERVER_CKPT_EVERY": "1"}
    res = launch_single_tier(dict(ck, TEST_STEPS="3"))
    assert abs(res[0]["vals"][-1][0] - (1.0 - 0.1 * gsum * 3)) < 1e-5 and (tmp_path / "ck.server0l").exists()
    res = launch_single_tier(dict(ck, TEST_STEPS="2", GEOMX_SERVER_RESUME="1"))
    for r in res:
        for t, vals in enumerate(r["vals"]):
            for i, v in enumerate(vals):
                assert abs(v - ((1.0 + i) - 0.1 * gsum * (3 + t + 1))) < 1e-5, (t, i, v)
    # without the resume flag the same prefix is ignored: a fresh job starts from its own kv.init values
    res = launch_single_tier(dict(ck, TEST_STEPS="1"))
    assert abs(res[0]["vals"][0][0] - (1.0 - 0.1 * gsum)) < 1e-5
    g4 = 0.5 * (1 + 2 + 3 + 4)
    ck2 = {"TEST_MODE": "sgd", "GEOMX_SERVER_CKPT_PREFIX": str(tmp_path / "hips"), "GEOMX_SERVER_CKPT_EVERY": "2"}
    launch_hips(dict(ck2, TEST_STEPS="2"))
    assert (tmp_path / "hips.server0g").exists()
    res = launch_hips(dict(ck2, TEST_STEPS="1", GEOMX_SERVER_RESUME="1"))
    assert len(res) == 4
    for r in res:
        assert abs(r["vals"][0][0] - (1.0 - 0.1 * g4 * 3)) < 1e-4


def test_async_global_tier_with_bisparse():
    """MixedSync + Bi-Sparse: the reference leaves ``DataHandleAsyncBSCompressed`` empty (kvstore_dist_server.h:1700-1703); here every party's
    sparsified aggregate is applied by the global optimizer on arrival.  Constant gradients: dense keys move by exactly lr * party sum per
    arrival; the big key moves in the same direction (top-k of an all-equal tensor sends k entries per round, the rest accumulates)."""
    res = launch_hips({"TEST_MODE": "bsc_async", "TEST_KV": "dist_async", "MXNET_KVSTORE_SIZE_LOWER_BOUND": "100", "TEST_STEPS": "3"})
    assert len(res) == 4
    for r in res:
        last = r["vals"][-1]
        assert all(v == v and abs(v) < 1e3 for v in last)                            # finite
        assert last[0] < 1.0 and last[1] < 2.0                                         # dense keys descended
        # after 3 rounds both parties delivered at least twice: w0 <= 1 - 0.1 * 2 * min(party sums = 1.5, 3.5)
        assert last[0] <= 1.0 - 0.1 * 2 * 1.5 + 1e-4
        assert r["last"][2] <= 3.0 + 1e-6                                              # sparse key never moves against the gradient


def test_features_p3_2bit_fp16_async():
    res = launch_hips({"TEST_MODE": "p3", "ENABLE_P3": "1", "TEST_STEPS": "2"})
    gsum = 0.5 * (1 + 2 + 3 + 4)
    for r in res:
        assert abs(r["vals"][1][0] - (1.0 - 0.1 * gsum * 2)) < 1e-4
    res = launch_single_tier({"TEST_MODE": "2bit", "TEST_STEPS": "1"})
    for r in res:    # grads 1.0 and 2.0 with threshold 0.5 -> each worker contributes +0.5 -> w = 1 - 0.1*(0.5+0.5)
        assert abs(r["vals"][0][0] - (1.0 - 0.1 * 1.0)) < 1e-5
    res = launch_single_tier({"TEST_MODE": "fp16", "TEST_STEPS": "2"})
    for r in res:
        assert abs(r["vals"][1][0] - (1.0 - 0.1 * 1.5 * 2)) < 5e-3
    res = launch_hips({"TEST_MODE": "async", "TEST_KV": "dist_async", "TEST_STEPS": "2"})
    assert len(res) == 4
    finals = [r["vals"][1][0] for r in res]
    # MixedSync: every party's aggregate is applied once per step, in arrival order: after 2 steps all 4 updates of both steps landed
    assert min(finals) >= 1.0 - 0.1 * gsum * 2 - 1e-4 and max(finals) <= 1.0


def test_dgt_priority_channels_and_resend():
    res = launch_hips({"TEST_MODE": "big", "ENABLE_DGT": "2", "DGT_BLOCK_SIZE": "1024", "DMLC_K": "0.5", "TEST_STEPS": "2"})
    gsum = 0.5 * (1 + 2 + 3 + 4)
    for r in res:
        assert abs(r["vals"][1][1] - (2.0 - 0.1 * gsum * 2)) < 1e-4 and abs(r["last"][1] - r["vals"][0][1]) < 1e-6
    cfg = {"TEST_MODE": "sgd", "PS_RESEND": "1", "PS_RESEND_TIMEOUT": "200", "PS_DROP_MSG": "10", "TEST_STEPS": "2"}
    try:
        res = launch_single_tier(cfg)
    except AssertionError as e:
        # random message loss: about one run in several dozen (under a loaded machine) a role is still in its tear-down when the harness
        # deadline passes although every worker has already reported correct values; the arithmetic check below is what this test is about
        if "timeout after" not in str(e):
            raise
        res = launch_single_tier(cfg)
    for r in res:
        assert abs(r["vals"][1][0] - (1.0 - 0.1 * 1.5 * 2)) < 1e-5


def test_tsengine_intra_party_merge_and_relay():
    """ENABLE_INTRA_TS on one party of 4 workers: pushes are merged peer-to-peer on their way to the server, fresh parameters arrive through
    the scheduler-routed relay; the arithmetic must be exactly the plain dist_sync one and the overlay must actually have been used."""
    res = launch_single_tier({"TEST_MODE": "sgd", "ENABLE_INTRA_TS": "1", "TEST_STEPS": "4"}, workers=4)
    assert len(res) == 4
    gsum = 0.5 * (1 + 2 + 3 + 4)
    relays = merges = 0
    for r in res:
        for t, vals in enumerate(r["vals"]):
            for i, v in enumerate(vals):
                assert abs(v - ((1.0 + i) - 0.1 * gsum * (t + 1))) < 1e-4, (t, i, v)
        merges += r["ts_stats"][0]; relays += r["ts_stats"][1]
    # 4 keys x 4 rounds reach 4 workers each: the server sends at least one copy per (key, round), workers relay the rest
    assert relays + merges > 0, res


def test_tsengine_two_tier_intra_and_inter():
    """Full HiPS (2 parties x 2 workers) with both overlays: workers merge inside the party, local servers merge between the parties and the
    global server's fresh values return by relay."""
    res = launch_hips({"TEST_MODE": "sgd", "ENABLE_INTRA_TS": "1", "ENABLE_INTER_TS": "1"})
    assert len(res) == 4
    gsum = 0.5 * (1 + 2 + 3 + 4)
    for r in res:
        for t, vals in enumerate(r["vals"]):
            for i, v in enumerate(vals):
                assert abs(v - ((1.0 + i) - 0.1 * gsum * (t + 1))) < 1e-4, (t, i, v)


def test_row_sparse_push_pull_over_the_wire():
    """row_sparse gradients / row_sparse_pull through the native server (kvstore_dist.h PushRowSparse / PullRowSparse_): 2 workers, SGD lr 0.1.
    Row 3 receives both workers' rows, rows 5 / 6 one each; only the requested rows travel back."""
    res = launch_single_tier({"TEST_MODE": "rowsparse", "TEST_STEPS": "2"})
    assert len(res) == 2
    for r in res:
        for t, v in enumerate(r["vals"]):
            assert v["ids"] == [3, 5, 6]
            exp = [-0.1 * (1 + 2) * (t + 1), -0.1 * 1 * (t + 1), -0.1 * 2 * (t + 1)]
            assert all(abs(a - b) < 1e-5 for a, b in zip(v["rows"], exp)), (v, exp)


def test_tracker_cluster_backends_dry_run(tmp_path):
    """slurm / sge / kubernetes back-ends describe the same 12-process two-tier job (3rdparty/dmlc-core/tracker/dmlc_tracker/{slurm,sge,kubernetes}.py)."""
    from geomx_b200.tracker.launch import HipsJob, launch, kubernetes_manifest, _sge_script
    job = HipsJob(2, 1, 2, 1, 9092, ["c0", "h1", "h2"], {"ENABLE_P3": "1"})
    sl = launch(job, ["python", "examples/cnn.py"], "slurm", dry_run=True)
    assert len(sl) == 12 and all(c[0] == "srun" for _, c, _ in sl)
    w = [c for n, c, _ in sl if n == "party1_worker1"][0]
    assert w[-2:] == ["python", "examples/cnn.py"] and "h2" in w and any("DMLC_ROLE=worker" in a and "ENABLE_P3=1" in a for a in w)
    sg = launch(job, ["python", "examples/cnn.py"], "sge", log_dir=str(tmp_path), dry_run=True)
    assert len(sg) == 12 and sg[0][1][:3] == ["qsub", "-sync", "y"]
    procs = {p.name: (p, a, e) for p, a, e in job.command_lines(["python", "examples/cnn.py"])}
    script = _sge_script(*procs["party0_server"], "/work")
    assert "#$ -l hostname=h1" in script and "export DMLC_ROLE=server" in script and "export DMLC_PS_GLOBAL_ROOT_URI=c0" in script and "import geomx_b200" in script
    m = kubernetes_manifest(job, ["python", "examples/cnn.py"])
    assert m.count("kind: Pod") == 12 and m.count("kind: Service") == 12
    pod = m[m.index("name: hips-party0-server\n  namespace"):]
    pod = pod[:pod.index("---")]
    # rendezvous addresses point at the schedulers' Services, not at host names
    assert 'DMLC_PS_GLOBAL_ROOT_URI, value: "hips-global-scheduler"' in pod and 'DMLC_PS_ROOT_URI, value: "hips-party0-scheduler"' in pod
    assert "nvidia.com/gpu" not in pod and "nvidia.com/gpu" in m[m.index("name: hips-party1-worker0\n  namespace"):][:1500]


def test_tracker_launcher_single_tier_and_layout(tmp_path):
    """geomx_b200.tracker: the two-tier layout is the reference's 12 processes, and a locally launched single-tier job trains correctly."""
    from geomx_b200.tracker import HipsJob, launch
    names = [p.name for p in HipsJob(2, 1, parties=2).processes()]
    assert len(names) == 12 and names[:4] == ["global_scheduler", "global_server0", "central_scheduler", "master_worker"]
    plan = launch(HipsJob(2, 1, parties=2, hosts=["10.0.0.1", "10.0.0.2", "10.0.0.3"]), ["python", "train.py"], launcher="ssh", dry_run=True)
    assert plan[0][1][0] == "ssh" and "10.0.0.1" in plan[0][1] and any("10.0.0.3" in c for _, c, _ in plan)
    env = {"TEST_MODE": "sgd", "TEST_STANDALONE": "1", "TEST_STEPS": "2", "PYTHONPATH": os.path.dirname(HERE)}
    job = HipsJob(2, 1, base_port=free_port(), extra_env=env)
    rc = launch(job, [sys.executable, WORKER], log_dir=str(tmp_path), timeout=120)
    assert rc == 0
    outs = [open(os.path.join(str(tmp_path), "worker%d.log" % i)).read() for i in range(2)]
    res = results(outs)
    assert len(res) == 2 and all(abs(r["vals"][1][0] - (1.0 - 0.1 * 1.5 * 2)) < 1e-5 for r in res)


def test_dgt_udp_datagram_channel():
    """ENABLE_DGT=1: unimportant blocks leave the local servers as UDP datagrams (per-message TOS) and are reassembled next to the TCP-borne
    important blocks on the global server; with 20 % emulated loss the tensor is still delivered (zero-filled gaps) and training proceeds."""
    res = launch_hips({"TEST_MODE": "big", "ENABLE_DGT": "1", "DGT_BLOCK_SIZE": "1024", "DMLC_K": "0.3", "TEST_STEPS": "2", "GEOMX_NET_STATS": "1"})
    workers = [r for r in res if "vals" in r]
    stats = [r["net_stats"] for r in res if "net_stats" in r and r["net_stats"]["plane"] == 1]
    assert len(workers) == 4
    gsum = 0.5 * (1 + 2 + 3 + 4)
    sent, recvd = sum(s["udp_sent"] for s in stats), sum(s["udp_received"] for s in stats)
    assert sent > 0 and 0 < recvd <= sent
    if recvd == sent:                                           # nothing was lost on the loopback: exact arithmetic
        for r in workers:
            assert abs(r["vals"][1][1] - (2.0 - 0.1 * gsum * 2)) < 1e-4
    res = launch_hips({"TEST_MODE": "big", "ENABLE_DGT": "1", "DGT_BLOCK_SIZE": "1024", "DMLC_K": "0.3", "DGT_UDP_LOSS": "20", "TEST_STEPS": "2"})
    assert len([r for r in res if "vals" in r]) == 4            # lossy: values are not exact, but every round completes


def test_plain_c_api_worker():
    """GXKVStore* (csrc/hips/c_api.cc): two workers that use only the C ABI train against the standard scheduler / server processes."""
    port = free_port()
    base = {"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": port, "DMLC_NUM_SERVER": 1, "DMLC_NUM_WORKER": 2, "DMLC_NUM_ALL_WORKER": 2}
    procs = [spawn(dict(base, DMLC_ROLE="scheduler")), spawn(dict(base, DMLC_ROLE="server"))]
    capi = os.path.join(HERE, "_capi_worker.py")
    for i in range(2):
        e = dict(os.environ); e.update({k: str(v) for k, v in dict(base, DMLC_ROLE="worker").items()}); e.pop("RANK", None); e.pop("WORLD_SIZE", None)
        procs.append(subprocess.Popen([sys.executable, capi], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    res = results(collect(procs))
    assert len(res) == 2
    for r in res:
        assert r["num_workers"] == 2 and abs(r["vals"][0] - (1.0 - 0.1 * 1.5)) < 1e-6 and abs(r["vals"][1] - (1.0 - 0.2 * 1.5)) < 1e-6


def test_hybrid_two_boxes_over_tcp():
    """KVStoreHybrid: 2 "boxes" x 2 ranks (gloo inside a box — NCCL on real GPUs), one TCP endpoint per box against scheduler + server.
    Every rank of every box must see  w_0 - t * lr * sum over all 4 ranks of the gradients."""
    port = free_port()
    base = {"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": port, "DMLC_NUM_SERVER": 1, "DMLC_NUM_WORKER": 2, "DMLC_NUM_ALL_WORKER": 4,
            "TEST_STANDALONE": 1, "TEST_MODE": "sgd", "TEST_STEPS": 2}
    procs = [spawn(dict(base, DMLC_ROLE="scheduler")), spawn(dict(base, DMLC_ROLE="server"))]
    gid = 0
    for box in range(2):
        mport = free_port()
        for r in range(2):
            e = dict(os.environ); e.update({k: str(v) for k, v in base.items()})
            e.update({"DMLC_ROLE": "worker", "RANK": str(r), "LOCAL_RANK": str(r), "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(mport),
                      "TEST_WORKER_GID": str(gid), "CUDA_VISIBLE_DEVICES": ""})
            gid += 1
            procs.append(subprocess.Popen([sys.executable, WORKER], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    res = results(collect(procs))
    assert len(res) == 4
    gsum = 0.5 * (1 + 2 + 3 + 4)
    for r in res:
        assert r["num_workers"] == 2 and r["num_all_workers"] == 4
        for t, vals in enumerate(r["vals"]):
            for i, v in enumerate(vals):
                assert abs(v - ((1.0 + i) - 0.1 * gsum * (t + 1))) < 1e-4, (t, i, v)


def test_heartbeat_detects_dead_scheduler(tmp_path):
    """PS_HEARTBEAT_INTERVAL / get_num_dead_node: workers heart-beat with their scheduler; after the scheduler is killed they report it dead
    within the timeout (reference: van.cc:242-257,1128-1140; postoffice.cc:284-303)."""
    import time
    port = free_port()
    ready = str(tmp_path / "ready")
    base = {"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": port, "DMLC_NUM_SERVER": 1, "DMLC_NUM_WORKER": 2, "DMLC_NUM_ALL_WORKER": 2,
            "TEST_STANDALONE": 1, "TEST_MODE": "heartbeat", "PS_HEARTBEAT_INTERVAL": 1, "PS_HEARTBEAT_TIMEOUT": 2, "TEST_READY_FILE": ready}
    sched = spawn(dict(base, DMLC_ROLE="scheduler"))
    server = spawn(dict(base, DMLC_ROLE="server"))
    ws = [spawn(dict(base, DMLC_ROLE="worker", TEST_WORKER_GID=i), worker=True) for i in range(2)]
    try:
        t0 = time.time()
        while not (os.path.exists(ready + "0") and os.path.exists(ready + "1")):
            assert time.time() - t0 < 90, "workers never became ready"
            assert all(w.poll() is None for w in ws), "a worker died early"
            time.sleep(0.2)
        time.sleep(1.5)                               # a few heart-beats while everything is alive
        sched.kill()
        outs = [w.communicate(timeout=60)[0] for w in ws]
    finally:
        for p in (sched, server, *ws):
            if p.poll() is None:
                p.kill()
    res = results(outs)
    assert len(res) == 2 and all(r["alive_first"] == 0 and r["dead"] == 1 for r in res), res


def test_dead_worker_is_replaced():
    """Elastic recovery on the local tier: a worker crashes after round 1, heart-beats expire, a restarted worker registers, receives the dead
    node's id (is_recovery), skips the start-up barriers / key initialisation, and round 2 completes with it (reference van.cc:90-111,176-192)."""
    import time
    port = free_port()
    base = {"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": port, "DMLC_NUM_SERVER": 1, "DMLC_NUM_WORKER": 2, "DMLC_NUM_ALL_WORKER": 2,
            "TEST_STANDALONE": 1, "TEST_MODE": "recovery", "PS_HEARTBEAT_INTERVAL": 1, "PS_HEARTBEAT_TIMEOUT": 2}
    procs = [spawn(dict(base, DMLC_ROLE="scheduler")), spawn(dict(base, DMLC_ROLE="server"))]
    w0 = spawn(dict(base, DMLC_ROLE="worker", TEST_WORKER_GID=0), worker=True)
    w1 = spawn(dict(base, DMLC_ROLE="worker", TEST_WORKER_GID=1), worker=True)
    w1b = None
    try:
        out1, _ = w1.communicate(timeout=90)
        assert w1.returncode == 17, out1
        time.sleep(4.0)                                # heart-beat timeout: the scheduler now considers the crashed worker dead
        w1b = spawn(dict(base, DMLC_ROLE="worker", TEST_WORKER_GID=1, TEST_RECOVERED=1), worker=True)
        outs = [out1] + collect(procs + [w0, w1b], timeout=120)
    finally:
        for p in procs + [w0, w1] + ([w1b] if w1b else []):
            if p.poll() is None:
                p.kill()
    res = {(r["gid"], bool(r.get("recovered"))): r for r in results(outs) if "gid" in r}
    assert res[(1, False)]["crashed"] and abs(res[(1, False)]["vals"][1] - 0.85) < 1e-5
    assert res[(1, True)]["is_recovery"] and abs(res[(1, True)]["vals"][0] - 0.85) < 1e-5 and abs(res[(1, True)]["vals"][1] - 0.70) < 1e-5
    assert abs(res[(0, False)]["vals"][1] - 0.85) < 1e-5 and abs(res[(0, False)]["vals"][2] - 0.70) < 1e-5
