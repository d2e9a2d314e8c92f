#!/usr/bin/env python
"""potts/neighonly sc/26n N^3 in automatic mode: per-sweep time, variant and mean distinct values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spparks_b200 import LatticeSweep  # noqa: E402

n, nsweeps = int(sys.argv[1]), int(sys.argv[2])
s = LatticeSweep("potts/neighonly", "sc/26n", n, n, n, nspins=10000, temperature=0.5, seed=12345, device=0)
s.randomize(12345, 1, 10000)
s.sync()
for sw in range(nsweeps):
    s.timer_start()
    s.sweeps(sw, 1)
    ms = s.timer_stop()
    v, m = s.neigh_variant(True)
    if sw < 40 or sw % 10 == 0:
        print("sweep %3d %.3f ms %.1f Gsites/s %s mean distinct %.2f" % (sw, ms, n ** 3 / ms / 1e6, v, m), flush=True)
