"""Runs tests/run_dist_check.py with 2 ranks: on 2 GPUs over NCCL when the box has them (the collectives then run
inside the library, csrc/comm.cu), otherwise with both ranks on GPU 0 over gloo -- never skipped."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, port, env=None, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))


def test_two_rank_step_equals_single_gpu_step():
    env = {} if torch.cuda.device_count() >= 2 else {"QTX_DIST_CHECK_BACKEND": "gloo"}
    out = _torchrun("run_dist_check.py", 29533, env)
    assert "DIST_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="the collectives of csrc/comm.cu need one GPU per NCCL rank")
def test_two_gpu_step_with_torch_distributed_collectives():
    """The same check with the collectives in torch.distributed (QTX_DIST_C=0), the path the gloo tests share."""
    out = _torchrun("run_dist_check.py", 29537, {"QTX_DIST_C": "0"})
    assert "DIST_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="peer memory needs two GPUs")
def test_fused_gram_exchange_equals_nccl_allreduce():
    """qtx_gram_push + qtx_peer_signal + qtx_gram_reduce over NVLink peer memory against Gram + NCCL all-reduce."""
    out = _torchrun("run_p2p_gram_check.py", 29541, {"QTX_P2P_TIMEOUT_S": "20"}, timeout=300)
    assert "P2P_GRAM_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
