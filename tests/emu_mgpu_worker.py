"""Worker for tests/test_emulated_library.py (one process per rank, gloo): the N-slab == 1-slab check
of spparks_b200/selfcheck.py — the one tests/mgpu_worker.py and bench.py run on the GPUs — against
the emulated build of the library: CUDA IPC over a shared file mapping, NCCL over shared memory
(tests/host_shim/emu)."""
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spparks_b200.selfcheck import mgpu_parity  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    halo = "nccl" if os.environ.get("SPK_CUDA_NO_P2P") == "1" else "p2p"
    verdict = mgpu_parity(dist, rank, world, local, halo=halo, log=lambda m: print(m, flush=True))
    if rank == 0:
        print("MGPU_ALL_OK" if verdict == "ok" else "MGPU_FAIL " + verdict, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
