"""2-GPU parity of the fused (peer-store) all-gather against the NCCL path and the oracle; skipped on a
single-GPU box.  The work is done by tests/mr_gpu_check.py under torchrun."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_gpu_peer_exchange_matches_nccl_and_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(here, "mr_gpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MULTIGPU_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
