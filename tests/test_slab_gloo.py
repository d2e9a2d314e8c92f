"""Slab decomposition + ghost-plane protocol on CPU: world_size 2 over gloo (host-side logic of the
multi-GPU path; the CUDA/NCCL version of the same check is tests/test_multigpu.py)."""
import os
import subprocess
import sys

import pytest

from spparks_b200 import slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds():
    assert slab.slab_bounds("sc/26n", (8, 8, 16), 1, 4) == (4, 8)
    assert slab.site_range("sc/26n", (8, 6, 16), 3, 4) == (12 * 48, 16 * 48)
    assert slab.site_range("tri", (6, 8, 1), 1, 2) == (4 * 12, 8 * 12)
    assert slab.ring(0, 4) == (3, 1)
    with pytest.raises(ValueError):
        slab.slab_bounds("sc/26n", (8, 8, 12), 0, 4)  # thickness 3 is odd
    with pytest.raises(ValueError):
        slab.slab_bounds("sq/4n", (8, 10, 1), 0, 4)


def test_two_ranks_gloo_equal_single_domain():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29741", os.path.join(ROOT, "tests", "gloo_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert "GLOO_SLAB_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
