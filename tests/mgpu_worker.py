"""Worker for tests/test_multigpu.py (run under torchrun, one rank per GPU): N slabs == 1 GPU,
spin for spin, through both ghost-refresh paths (spparks_b200/selfcheck.py; bench.py runs the
same check before it times a multi-GPU run)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spparks_b200.selfcheck import mgpu_parity  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    halo = "nccl" if os.environ.get("SPK_CUDA_NO_P2P") == "1" else "p2p"
    verdict = mgpu_parity(dist, rank, world, local, halo=halo, log=lambda m: print(m, flush=True))
    if rank == 0:
        print("MGPU_ALL_OK" if verdict == "ok" else "MGPU_FAIL " + verdict, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
