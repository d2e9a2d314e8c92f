"""N-GPU == 1-GPU check of the slab decomposition, shared by tests/mgpu_worker.py and
bench.py (`mgpu_parity`).  One process per GPU under torchrun; `dist` is an initialised
torch.distributed (NCCL) module.  No oracle involved: the product's multi-GPU path (ghost refresh
as CommLattice::all() does before every colour set, src/app_lattice.cpp:648-653,
src/comm_lattice.cpp:912-946) is compared with the product's single-GPU path, which the GPU parity
tests pin to the reference."""
import os

import numpy as np
import torch

from .sweep import LatticeSweep

NBASIS = {"tri": 2}


def cases(world):
    """(model, lattice, n, Q, T): slab dimensions split into `world` even slabs (2-D: >= 4 rows each)."""
    z = max(16, 2 * world)
    y = max(32, 4 * world)
    return [
        ("potts", "sc/26n", (32, 16, z + 16 if (z + 16) % (2 * world) == 0 else 2 * z), 20, 0.6),
        ("potts/neighonly", "sc/26n", (24, 12, z), 50, 0.5),
        ("potts/pin", "sc/26n", (16, 16, z), 9, 0.9),
        ("potts/neigh", "sc/26n", (16, 12, z), 6, 0.6),
        ("ising", "sc/6n", (16, 8, z), 2, 2.2),
        ("ising", "sq/4n", (64, y, 1), 2, 1.13),
        ("potts", "tri", (40, y, 1), 7, 0.5),
        ("potts", "sq/8n", (20, y, 1), 5, 0.7),
    ]


def mgpu_parity(dist, rank, world, local, halo="p2p", log=None):
    """Runs every case on `world` slabs and, on rank 0, on one GPU; returns "ok" or a description
    of the first failing case (the same string on every rank).  halo: "p2p" (peer-memory stores
    where available) or "nccl" (send/recv exchange)."""
    # the tensors of the check's own plumbing live where `dist` moves them: on the GPU for NCCL; a
    # CPU backend (gloo) is what the emulated test build of the library is driven with (tests/host_shim)
    on_gpu = torch.cuda.is_available() and str(dist.get_backend()) != "gloo"
    tdev = "cuda" if on_gpu else "cpu"
    old = os.environ.get("SPK_CUDA_NO_P2P")
    if halo == "nccl":
        os.environ["SPK_CUDA_NO_P2P"] = "1"
    else:
        os.environ.pop("SPK_CUDA_NO_P2P", None)
    verdict = "ok"
    try:
        for model, lat, n, q, T in cases(world):
            dim3 = lat.startswith("sc")
            nb = NBASIS.get(lat, 1)
            nsites = n[0] * n[1] * n[2] * nb
            rng = np.random.default_rng(42)
            spins = rng.integers(1, (2 if model.startswith("ising") else q) + 1, size=nsites, dtype=np.int32)
            if model == "potts/pin":
                spins[rng.random(nsites) < 0.05] = q + 1
            s = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=777, device=local,
                             rank=rank, nranks=world)
            idt = torch.zeros(128, dtype=torch.uint8, device=tdev)
            if rank == 0:
                idt = torch.frombuffer(bytearray(LatticeSweep.comm_id()), dtype=torch.uint8).to(tdev)
            dist.broadcast(idt, 0)
            s.comm_init(idt.cpu().numpy().tobytes())
            fused2d = not dim3 and (n[0] * nb) % 8 == 0  # the colour-fused 2-D kernel takes rows of 8-element multiples
            want_mode = 2 if (dim3 or fused2d) and halo == "p2p" else 1
            fail = None
            if s.halo_mode() != want_mode:
                fail = "halo mode %d, expected %d" % (s.halo_mode(), want_mode)
            lo, hi = rank * s.nlocal, (rank + 1) * s.nlocal
            s.upload(spins[lo:hi])
            s.sweeps(3, 2)
            s.sweeps(5, 2)  # second call: ghosts are still valid, no re-seeding exchange
            mine = torch.from_numpy(s.download()).to(tdev)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            mine_counts = list(s.counters())
            acc = torch.tensor(mine_counts + [s.energy()], dtype=torch.float64, device=tdev)
            per_rank = [torch.zeros_like(acc) for _ in range(world)]
            dist.all_gather(per_rank, acc)
            acc = torch.stack(per_rank).sum(0)
            if on_gpu:
                torch.cuda.synchronize()  # torch's collectives are done before the library's own (another communicator) start
            clus = s.cluster_stats()  # collective: clusters merged across the slab boundaries
            if rank == 0 and fail is None:
                got = torch.cat(parts).cpu().numpy()
                one = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=777, device=local)
                one.upload(spins)
                one.sweeps(3, 4)
                a, b = one.counters()
                if not np.array_equal(got, one.download()):
                    fail = "spins differ from the 1-GPU run"
                elif [int(acc[0]), int(acc[1])] != [a, b]:
                    fail = "counters differ: per rank %s, 1 GPU %s" % ([tuple(int(v) for v in t[:2].tolist()) for t in per_rank], (a, b))
                elif float(acc[2]) != one.energy():
                    fail = "energy differs"
                else:
                    c1 = one.cluster_stats()
                    if clus[0] != c1[0] or clus[1] != c1[1] or abs(clus[2] - c1[2]) > 1e-9 * c1[2]:
                        fail = "cluster statistics %r differ from the 1-GPU run's %r" % (clus, c1)
                one.close()
            s.close()
            flag = torch.tensor([0 if fail is None else 1], dtype=torch.int32, device=tdev)
            dist.broadcast(flag, 0)
            if log and rank == 0:
                log("%s %s %s %s ranks %d halo %s" % ("ok" if fail is None else "FAIL", model, lat, n, world, halo))
            if int(flag.item()):
                verdict = "%s %s %s halo=%s: %s" % (model, lat, "x".join(map(str, n)), halo, fail or "failed on rank 0")
                break
        # clusters that cross every slab boundary and wrap around the periodic box: diagonal bands and
        # one percolating background, no sweeps (pure test of the cross-slab merge)
        for lat, n in (("sc/26n", (16, 12, max(16, 2 * world))), ("sq/4n", (40, max(32, 4 * world), 1)), ("tri", (24, max(32, 4 * world), 1))):
            if verdict != "ok":
                break
            dim3 = lat.startswith("sc")
            nb = NBASIS.get(lat, 1)
            zz, yy, xx, mm = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), np.arange(nb), indexing="ij")
            slab = zz if dim3 else yy
            spins = np.where((xx + 3 * slab) % 8 < 3, 2 + (xx + 3 * slab) // 8 % 3, 1).astype(np.int32).reshape(-1)
            s = LatticeSweep("potts", lat, n[0], n[1], n[2], nspins=5, temperature=0.5, seed=1, device=local, rank=rank, nranks=world)
            idt = torch.zeros(128, dtype=torch.uint8, device=tdev)
            if rank == 0:
                idt = torch.frombuffer(bytearray(LatticeSweep.comm_id()), dtype=torch.uint8).to(tdev)
            dist.broadcast(idt, 0)
            s.comm_init(idt.cpu().numpy().tobytes())
            s.upload(spins[rank * s.nlocal:(rank + 1) * s.nlocal])
            clus = s.cluster_stats()
            s.close()
            fail = None
            if rank == 0:
                one = LatticeSweep("potts", lat, n[0], n[1], n[2], nspins=5, temperature=0.5, seed=1, device=local)
                one.upload(spins)
                c1 = one.cluster_stats()
                one.close()
                if clus[0] != c1[0] or clus[1] != c1[1] or abs(clus[2] - c1[2]) > 1e-9 * c1[2]:
                    fail = "cluster statistics %r differ from the 1-GPU run's %r" % (clus, c1)
            flag = torch.tensor([0 if fail is None else 1], dtype=torch.int32, device=tdev)
            dist.broadcast(flag, 0)
            if log and rank == 0:
                log("%s banded clusters %s %s ranks %d: %r" % ("ok" if fail is None else "FAIL", lat, n, world, clus))
            if int(flag.item()):
                verdict = "banded clusters %s: %s" % (lat, fail or "failed on rank 0")
    finally:
        if old is None:
            os.environ.pop("SPK_CUDA_NO_P2P", None)
        else:
            os.environ["SPK_CUDA_NO_P2P"] = old
    return verdict
