"""N-GPU == 1-GPU bit-exactness, launched as one process per GPU under torchrun."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("halo", ["p2p", "nccl"])
def test_two_slabs_equal_one_gpu(halo):
    """Both ghost-refresh paths of spk_cuda_sweeps: peer-memory stores (3-D default) and the NCCL
    send/recv exchange (SPK_CUDA_NO_P2P=1; always used by the 2-D lattices)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    if halo == "nccl":
        env["SPK_CUDA_NO_P2P"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "MGPU_ALL_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
