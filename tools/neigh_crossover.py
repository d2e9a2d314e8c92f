#!/usr/bin/env python
"""Chain vs candidate-elimination kernel for potts/neighonly sc/26n along a coarsening run.

    python tools/neigh_crossover.py [N] [checkpoints...]

Two handles on the same lattice (same seed => identical trajectories), one forced to each
variant (SPK_CUDA_NEIGH_MODE is read at create); at each checkpoint both time 3 sweeps.
Prints sweep, mean distinct neighbour values, ms/sweep and Gsites/s of each variant."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spparks_b200 import LatticeSweep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
checks = [int(v) for v in sys.argv[2:]] or [0, 3, 5, 8, 10, 15, 20, 30, 50, 100, 200]
hs = {}
for mode in ("chain", "adaptive"):
    os.environ["SPK_CUDA_NEIGH_MODE"] = mode
    s = LatticeSweep("potts/neighonly", "sc/26n", n, n, n, nspins=10000, temperature=0.5, seed=12345, device=0)
    s.randomize(12345, 1, 10000)
    s.sync()
    hs[mode] = s
sweep = 0
for c in checks:
    for s in hs.values():
        if c > sweep:
            s.sweeps(sweep, c - sweep)
    sweep = c
    row = []
    for mode, s in hs.items():
        s.sync()
        s.reset_counters()
        s.timer_start()
        s.sweeps(sweep, 3)
        ms = s.timer_stop() / 3
        nacc, natt = s.counters()
        row.append("%s %.3f ms %.1f Gsites/s acc %.3f" % (mode, ms, n ** 3 / ms / 1e6, nacc / natt))
    sweep += 3
    print("sweep %4d  %s" % (c, "   ".join(row)), flush=True)
a, b = hs["chain"].download(), hs["adaptive"].download()
print("identical lattices:", bool((a == b).all()))
