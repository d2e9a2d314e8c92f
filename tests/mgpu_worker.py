"""Worker for tests/test_multigpu.py (run under torchrun, one rank per GPU).  SPK_CUDA_NO_P2P=1
in the environment keeps the NCCL exchange for the 3-D cases (default: peer-memory halo).

Checks that an N-slab run (ghost planes exchanged with NCCL before every colour, as
CommLattice::all() does, src/app_lattice.cpp:650-651) reproduces the 1-GPU run spin for spin, for
3-D (z slabs) and 2-D (y slabs) lattices, and that counters and energy sum up."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spparks_b200 import LatticeSweep  # noqa: E402

CASES = [
    ("potts", "sc/26n", (32, 16, 24), 20, 0.6),
    ("potts/neighonly", "sc/26n", (24, 12, 16), 50, 0.5),
    ("potts/pin", "sc/26n", (16, 16, 8), 9, 0.9),
    ("potts/neigh", "sc/26n", (16, 12, 8), 6, 0.6),
    ("ising", "sc/6n", (16, 8, 16), 2, 2.2),
    ("ising", "sq/4n", (64, 32, 1), 2, 1.13),
    ("potts", "tri", (40, 24, 1), 7, 0.5),
    ("potts", "sq/8n", (20, 16, 1), 5, 0.7),
]
NBASIS = {"tri": 2}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for model, lat, n, q, T in CASES:
        dim3 = lat.startswith("sc")
        nb = NBASIS.get(lat, 1)
        nsites = n[0] * n[1] * n[2] * nb
        rng = np.random.default_rng(42)
        spins = rng.integers(1, (2 if model.startswith("ising") else q) + 1, size=nsites, dtype=np.int32)
        if model == "potts/pin":
            spins[rng.random(nsites) < 0.05] = q + 1
        s = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=777, device=local,
                         rank=rank, nranks=world)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(LatticeSweep.comm_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        s.comm_init(idt.cpu().numpy().tobytes())
        fused2d = not dim3 and (n[0] * nb) % 8 == 0  # the colour-fused 2-D kernel takes rows of 8-element multiples
        want_mode = 2 if (dim3 or fused2d) and os.environ.get("SPK_CUDA_NO_P2P") != "1" else 1
        assert s.halo_mode() == want_mode, (model, lat, "halo mode", s.halo_mode())
        lo, hi = rank * s.nlocal, (rank + 1) * s.nlocal
        s.upload(spins[lo:hi])
        s.sweeps(3, 2)
        s.sweeps(5, 2)  # second call: ghosts are still valid, no re-seeding exchange
        mine = torch.from_numpy(s.download()).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        acc = torch.tensor(list(s.counters()) + [s.energy()], dtype=torch.float64, device="cuda")
        dist.all_reduce(acc)
        if rank == 0:
            got = torch.cat(parts).cpu().numpy()
            one = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=777, device=local)
            one.upload(spins)
            one.sweeps(3, 4)
            want = one.download()
            assert np.array_equal(got, want), (model, lat, "spins differ from the 1-GPU run")
            a, b = one.counters()
            assert [int(acc[0]), int(acc[1])] == [a, b], (model, lat, "counters")
            assert float(acc[2]) == one.energy(), (model, lat, "energy")
            print("ok", model, lat, n, "ranks", world, flush=True)
        s.close()
        dist.barrier()
    if rank == 0:
        print("MGPU_ALL_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
