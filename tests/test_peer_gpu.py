"""GPU, >= 2 devices: K7 (q100_allreduce_oneshot, one-shot all-reduce over peer memory) through torchrun --
bit-equality with an fp32 sum of the all-gathered partials, eager and inside a CUDA graph
(tools/peer_allreduce_check.py).  Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_oneshot_allreduce_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "peer_allreduce_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "PEER ALLREDUCE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_scatter_reduce_two_ranks():
    """K7b: GEMM with the reduce-scatter as its epilogue + reduce / broadcast kernel == fp32 sum of the gathered
    partials, bit for bit, eagerly and in a CUDA graph (tools/peer_scatter_check.py)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.join(ROOT, "tools", "peer_scatter_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PEER SCATTER OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
