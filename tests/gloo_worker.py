"""CPU worker (torch.distributed, gloo, world_size 2) for tests/test_slab_gloo.py.

Every rank keeps a full-size array but owns only its slab: it updates its slab's sites with the
oracle (box = slab) and, before every colour, refreshes the one ghost plane on each side from its
ring neighbours — the protocol libspparks_cuda runs over NCCL.  Planes outside slab+ghosts are
poisoned, so a wrong exchange shows up as a mismatch with the single-domain run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
from spparks_b200 import slab  # noqa: E402


def exchange(full, lattice, n, rank, world):
    p = slab.plane_sites(lattice, n)
    lo, hi = slab.slab_bounds(lattice, n, rank, world)
    ext = slab.slab_extent(lattice, n)
    lower, upper = slab.ring(rank, world)
    first = torch.from_numpy(full[lo * p:(lo + 1) * p].copy())
    last = torch.from_numpy(full[(hi - 1) * p:hi * p].copy())
    ghost_lo, ghost_hi = torch.empty_like(first), torch.empty_like(first)
    # same pairing as spk_cuda_halo_exchange: sends [first -> lower, last -> upper],
    # receives [upper's first -> my upper ghost, lower's last -> my lower ghost]
    ops = [dist.P2POp(dist.isend, first, lower, tag=0), dist.P2POp(dist.isend, last, upper, tag=1),
           dist.P2POp(dist.irecv, ghost_hi, upper, tag=0), dist.P2POp(dist.irecv, ghost_lo, lower, tag=1)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    glo, ghi = (lo - 1) % ext, hi % ext
    full[glo * p:(glo + 1) * p] = ghost_lo.numpy()
    full[ghi * p:(ghi + 1) * p] = ghost_hi.numpy()


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = [("potts", "sc/26n", (8, 6, 12), 5, 0.7), ("ising", "sq/4n", (10, 16, 1), 2, 1.2),
             ("potts/neighonly", "tri", (6, 8, 1), 4, 0.5), ("potts/neigh", "sc/26n", (6, 6, 8), 4, 0.6),
             ("ising/single", "sc/6n", (6, 4, 8), 2, 2.0)]
    for model, lat, n, q, T in cases:
        o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
        rng = np.random.default_rng(3)
        spins = rng.integers(1, (2 if model.startswith("ising") else q) + 1, size=o.nsites, dtype=np.int32)
        want = spins.copy()
        for sw in range(3):
            o.sweep_philox(want, 99, sw)
        a, b = slab.site_range(lat, n, rank, world)
        full = np.full(o.nsites, -12345, dtype=np.int32)  # poison what this rank must never read
        full[a:b] = spins[a:b]
        box = slab.slab_box(lat, n, rank, world)
        for sw in range(3):
            for c in range(1, o.ncolors + 1):
                exchange(full, lat, n, rank, world)
                o.halfsweep_philox(full, 99, sw, c, box=box)
        mine = torch.from_numpy(full[a:b].copy())
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        got = torch.cat(parts).numpy()
        assert np.array_equal(got, want), (model, lat)
    if rank == 0:
        print("GLOO_SLAB_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
