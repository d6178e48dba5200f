"""Convergence of the demo scenarios: each test starts the reference's 12-process topology on localhost (scripts/cpu/run_*.sh, i.e. the
reference's own "integration tests") on the learnable synthetic MNIST stand-in and checks that the test accuracy printed by a
worker rises well above chance.  All ten demo scenarios are covered (about 15 s each with one BLAS thread per process)."""
import os
import re
import socket
import subprocess

import pytest

from geomx_b200 import runtime

pytestmark = pytest.mark.skipif(not runtime.available(), reason="native runtime not built")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_base_port(span=8):
    """A base port whose whole block (global scheduler, central party, one per party) is currently free."""
    for _ in range(200):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
        if p + span >= 60000:
            continue
        ok = True
        for q in range(p, p + span):
            t = socket.socket()
            try:
                t.bind(("127.0.0.1", q))
            except OSError:
                ok = False
            finally:
                t.close()
            if not ok:
                break
        if ok:
            return p
    return 23456


def _run_group(cmd, env, timeout):
    """Run a launcher script in its own process group; on timeout the whole group (all 12+ roles) is killed, not just the shell."""
    import signal
    p = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, start_new_session=True)
    try:
        out, _ = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        out, _ = p.communicate()
        out = (out or "") + "\n[harness] timed out after %d s, process group killed" % timeout
        return subprocess.CompletedProcess(cmd, 124, out)
    return subprocess.CompletedProcess(cmd, p.returncode, out)


def run_scenario(script, tmp_path, iters=31, extra_env=None, args=()):
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    env.update({"LOG_DIR": str(tmp_path), "BASE_PORT": str(_free_base_port()), "GEOMX_SYNTHETIC_SIZE": "2048", "GEOMX_MAX_ITERS": str(iters),
                "GEOMX_EVAL_EVERY": "10", "OMP_NUM_THREADS": "1", "MKL_NUM_THREADS": "1", "GEOMX_SEED": "11"})
    env.update(extra_env or {})
    for attempt in range(3):
        r = _run_group(["bash", os.path.join(ROOT, "scripts", "cpu", script), "-ep", "8"] + list(args), env, 240)
        logs = {f: open(os.path.join(str(tmp_path), f)).read() for f in os.listdir(str(tmp_path)) if f.endswith(".log")}
        if r.returncode == 0 or not any("bind failed" in v for v in list(logs.values()) + [r.stdout]):
            break
        # a port of the freshly probed block was taken by somebody's ephemeral socket in the meantime: pick another block
        env["BASE_PORT"] = str(_free_base_port())
    assert r.returncode == 0, r.stdout[-2000:] + "\n".join("%s:\n%s" % (k, v[-600:]) for k, v in logs.items())
    accs = [float(x) for x in re.findall(r"Test Acc ([0-9.]+)", logs["party1_worker1.log"])]
    assert accs, logs["party1_worker1.log"][-1500:]
    return accs


def test_fsa_vanilla_hips_converges(tmp_path):
    accs = run_scenario("run_vanilla_hips.sh", tmp_path, iters=41)      # unseeded Xavier init: after 30 steps some draws are still at ~0.45
    assert max(accs) > 0.5, accs


def test_bisparse_compression_converges(tmp_path):
    # sparsified (1 %) gradients + residual bursts make the curve noisy: judge the best evaluation, not the last one
    accs = run_scenario("run_bisparse_compression.sh", tmp_path, iters=61)
    assert max(accs) > 0.5, accs


def test_hfa_converges(tmp_path):
    env = {"EXTRA_SERVER_ENV": "MXNET_KVSTORE_USE_HFA=1 MXNET_KVSTORE_HFA_K1=2 MXNET_KVSTORE_HFA_K2=2",
           "EXTRA_WORKER_ENV": "MXNET_KVSTORE_USE_HFA=1 MXNET_KVSTORE_HFA_K1=2 MXNET_KVSTORE_HFA_K2=2"}
    env2 = dict(os.environ); env2.update(env)
    # the scenario script hard-codes K1=20/K2=10 (the reference's values); call the launcher directly with short periods
    e = dict(os.environ); e.pop("RANK", None); e.pop("WORLD_SIZE", None)
    e.update({"LOG_DIR": str(tmp_path), "BASE_PORT": str(_free_base_port()), "GEOMX_SYNTHETIC_SIZE": "2048", "GEOMX_MAX_ITERS": "41",
              "GEOMX_EVAL_EVERY": "10", "OMP_NUM_THREADS": "1", "N_GS": "1", "GEOMX_SEED": "11"})
    e.update(env)
    r = _run_group(["bash", os.path.join(ROOT, "scripts", "hips_launch.sh"), "cpu", os.path.join(ROOT, "examples", "cnn_hfa.py"), "-ep", "8"], e, 240)
    log = open(os.path.join(str(tmp_path), "party1_worker1.log")).read()
    assert r.returncode == 0, r.stdout[-1500:] + log[-1500:]
    accs = [float(x) for x in re.findall(r"Test Acc ([0-9.]+)", log)]
    assert accs and max(accs) > 0.5, (accs, log[-800:])


@pytest.mark.parametrize("script", ["run_mixed_sync.sh", "run_fp16.sh", "run_mixed_precision.sh", "run_dgt.sh", "run_p3.sh", "run_tsengine.sh",
                                    "run_multi_gps.sh"])
def test_remaining_demo_scenarios_converge(script, tmp_path):
    # MixedSync applies every party's (stale) aggregate on arrival: with Adam the effective step doubles, so it gets the smaller learning
    # rate such a schedule needs (at 0.01 the outcome depends on the arrival order of the first few updates)
    args = ("-lr", "0.003") if script == "run_mixed_sync.sh" else ()
    accs = run_scenario(script, tmp_path, iters=41, args=args)
    assert max(accs) > 0.5, (script, accs)      # reduced-precision / asynchronous scenarios are noisy on 40 iterations: judge the best evaluation
