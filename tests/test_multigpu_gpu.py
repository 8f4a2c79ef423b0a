"""Runs tests/run_dist_check.py on 2 GPUs when the box has them (skipped on 1-GPU boxes)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_step_equals_single_gpu_step():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "run_dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "DIST_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_gram_exchange_equals_nccl_allreduce():
    """qtx_gram_push + qtx_peer_signal + qtx_gram_reduce over NVLink peer memory against Gram + NCCL all-reduce."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "run_p2p_gram_check.py")]
    env = dict(os.environ, QTX_P2P_TIMEOUT_S="20")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert "P2P_GRAM_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
