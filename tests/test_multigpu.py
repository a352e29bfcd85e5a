"""Data-parallel learner on real GPUs (needs >= 2; skipped on the single-GPU box): tools/dp_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.timeout(180)
def test_data_parallel_learner_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=170)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "-> OK" in out.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.timeout(180)
def test_data_parallel_learner_two_gpus_peer_memory_allreduce():
    """The same check with the gradient all-reduce over NVLink peer memory (rtd3_p2p_allreduce) inside the update's CUDA graph."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tools", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=170, env=dict(os.environ, RTD3_DP_COLLECTIVE="p2p"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "-> OK" in out.stdout


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.timeout(180)
@pytest.mark.parametrize("env", [{"RTD3_P2P_MODE": "1"}, {"RTD3_P2P_FUSE": "1"}, {"RTD3_P2P_FUSE": "0"}],
                         ids=["wgrad-exchange-reduce-scatter", "allreduce-adam-kernel", "allreduce-then-adam"])
def test_data_parallel_learner_two_gpus_other_forms(env):
    """The forms of the peer-memory optimiser step that are not the 2-rank default (which is the exchange inside the weight-gradient
    kernels, all to all): the reduce-scatter + all-gather pattern used on more than 2 ranks, and the all-reduce kernels that larger
    batches keep - each must leave the replicas bit-identical and agree with the single-GPU step on the whole minibatch."""
    port = {"RTD3_P2P_MODE": 29535, "RTD3_P2P_FUSE": 29536 + int(env.get("RTD3_P2P_FUSE", 0))}[next(iter(env))]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=170, env=dict(os.environ, RTD3_DP_COLLECTIVE="p2p", **env))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "-> OK" in out.stdout
