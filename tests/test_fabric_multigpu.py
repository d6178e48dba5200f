"""Multi-GPU checks of the fused HiPS kernels (skipped on boxes with < 2 GPUs): each test launches one rank per GPU with torchrun and
runs a tool from tools/ that compares the in-kernel collectives with NCCL / plain PyTorch fp32 oracles."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _ngpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script, *args, port=29611, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script)] + list(args)
    e = dict(os.environ); e.update(env or {})
    r = subprocess.run(cmd, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    return r.returncode, r.stdout


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("parties", [1, 2])
@pytest.mark.parametrize("overlap", ["1", "0"])
def test_fused_step_matches_nccl_allreduce(parties, overlap):
    """overlap=1: per-key-group channels of the direct protocol (dense keys two-hop sharded underneath the conv backward pass, conv keys one-hop
    replicated); overlap=0: the single 3-hop LL exchange at the end of the step.  Both must equal w - lr * NCCL-all-reduce(grad)/B."""
    rc, out = _torchrun(2, "fabric_check.py", "--parties", str(parties), port=29611 + parties + 4 * int(overlap), env={"GEOMX_STEP_OVERLAP": overlap})
    assert rc == 0 and "FABRIC_CHECK PASS" in out, out[-3000:]


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_fused_step_flag_protocol():
    rc, out = _torchrun(2, "fabric_check.py", "--parties", "2", port=29621, env={"GEOMX_FABRIC_PROTOCOL": "bulk"})
    assert rc == 0 and "FABRIC_CHECK PASS" in out, out[-3000:]


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("parties", [1, 2])
def test_wire_formats_fp16_and_bisparse(parties):
    rc, out = _torchrun(2, "fabric_formats_check.py", str(parties), port=29631 + parties)
    assert rc == 0 and "FORMATS_CHECK PASS" in out, out[-3000:]


@pytest.mark.skipif(_ngpus() < 4, reason="needs >= 4 GPUs")
def test_two_parties_two_global_servers():
    rc, out = _torchrun(4, "fabric_check.py", "--parties", "2", "--gs", "2", port=29641, env={"GEOMX_STEP_OVERLAP": "0"})
    assert rc == 0 and "FABRIC_CHECK PASS" in out, out[-3000:]


@pytest.mark.skipif(_ngpus() < 4, reason="needs >= 4 GPUs")
def test_direct_channels_four_ranks():
    rc, out = _torchrun(4, "fabric_check.py", "--parties", "2", port=29645, env={"GEOMX_STEP_OVERLAP": "1"})
    assert rc == 0 and "FABRIC_CHECK PASS" in out, out[-3000:]


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_mixed_sync_one_sided_kernel():
    """dist_async on the fabric (per-tile locks in the global owner's HBM): every party applies its aggregate on arrival; the run must train
    and stay consistent inside a party."""
    rc, out = _torchrun(2, "fabric_check.py", "--parties", "2", "--mode", "dist_async", port=29651)
    assert rc == 0 and "FABRIC_CHECK PASS" in out, out[-3000:]


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["hostopt", "sched", "2bit", "rowsparse"])
def test_kvstore_api_surface_on_the_fabric(mode):
    """Python-executed optimizers (no native spec / lr_scheduler), priorities, row_sparse_pull and 2-bit compression through
    mx.kv.create('dist_sync') on the fabric."""
    rc, out = _torchrun(2, "fabric_api_check.py", mode, port=29681 + ["hostopt", "sched", "2bit", "rowsparse"].index(mode))
    assert rc == 0 and "API_CHECK %s PASS" % mode in out, out[-3000:]


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_group2ctx_places_stages_on_two_gpus():
    """Manual model parallelism on real devices: stage 1 on cuda:0, stage 2 on cuda:1, activations and gradients cross NVLink."""
    import numpy as np
    import geomx_b200 as mx
    with mx.AttrScope(ctx_group="stage1"):
        data = mx.sym.Variable("data")
        act = mx.sym.Activation(mx.sym.FullyConnected(data, num_hidden=16, name="fc1"), act_type="relu", name="act1")
    with mx.AttrScope(ctx_group="stage2"):
        out = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(act, num_hidden=4, name="fc2"), name="softmax")
    ex = out.simple_bind(mx.gpu(0), group2ctx={"stage1": mx.gpu(0), "stage2": mx.gpu(1)}, data=(8, 10), softmax_label=(8,))
    assert ex.arg_dict["fc1_weight"]._t.device.index == 0 and ex.arg_dict["fc2_weight"]._t.device.index == 1
    rng = np.random.RandomState(4)
    for n, a in ex.arg_dict.items():
        a[:] = (rng.randint(0, 4, a.shape) if n == "softmax_label" else rng.randn(*a.shape) * 0.3).astype(np.float32)
    ref = out.simple_bind(mx.cpu(0), data=(8, 10), softmax_label=(8,))
    for n in ex.arg_dict:
        ref.arg_dict[n][:] = ex.arg_dict[n].asnumpy()
    o = ex.forward(is_train=True); ex.backward()
    r = ref.forward(is_train=True); ref.backward()
    assert o[0]._t.device.index == 1 and np.allclose(o[0].asnumpy(), r[0].asnumpy(), atol=1e-5)
    assert ex.grad_dict["fc1_weight"]._t.device.index == 0
    assert np.allclose(ex.grad_dict["fc1_weight"].asnumpy(), ref.grad_dict["fc1_weight"].asnumpy(), atol=1e-5)


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_sync_batchnorm_across_gpus():
    """SyncBatchNorm with one GPU per rank: per-channel statistics and their gradients all-reduced over NCCL == BatchNorm on the whole batch."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29695",
           os.path.join(ROOT, "tests", "_syncbn_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, SYNCBN_DEVICE="cuda"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0 and "SYNCBN PASS" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_tsengine_flags_are_rejected_on_the_fabric():
    rc, out = _torchrun(2, "fabric_api_check.py", "hostopt", port=29691, env={"ENABLE_INTRA_TS": "1"})
    assert rc != 0 and "TSEngine" in out, out[-2000:]


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("config", ["bsc", "mpq_dgt", "hfa", "mixed_sync"])
def test_bench_configs_train(config):
    """BASELINE.json configs 3-5 (+ MixedSync) through bench.py on 2 GPUs: must run, keep the protocol error flag clear and report a finite loss."""
    import json
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "45", "--warmup", "5", "--config", config]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and line, r.stdout[-3000:]
    out = json.loads(line[-1])
    assert out["baseline_config"] == config and out["value"] > 0 and not out.get("protocol_errors"), out
    assert out["e2e"]["final_loss"] == out["e2e"]["final_loss"] and out["e2e"]["final_loss"] < 3.0, out["e2e"]


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_hfa_rounds():
    rc, out = _torchrun(2, "fabric_hfa_check.py", "1", port=29661)
    assert rc == 0 and "HFA_CHECK PASS" in out, out[-3000:]


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_reference_style_script_under_torchrun():
    """examples/cnn.py (init + pull interleaved key by key, push/pull per key, eval) on the fabric KVStore: must train on the synthetic set."""
    import re
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29671",
           os.path.join(ROOT, "examples", "cnn.py"), "-ep", "8"]
    e = dict(os.environ, GEOMX_SYNTHETIC_SIZE="2048", GEOMX_MAX_ITERS="51", GEOMX_EVAL_EVERY="10", GEOMX_NUM_PARTIES="2")
    r = subprocess.run(cmd, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    accs = [float(x) for x in re.findall(r"Test Acc ([0-9.]+)", r.stdout)]
    assert r.returncode == 0 and accs and max(accs) > 0.55, r.stdout[-3000:]
