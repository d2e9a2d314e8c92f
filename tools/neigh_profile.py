#!/usr/bin/env python
"""potts/neighonly sc/26n: W sweeps from a random start, then one more (for ncu --launch-skip 8*W+1).
    python tools/neigh_profile.py N W     (variant from SPK_CUDA_NEIGH_MODE)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spparks_b200 import LatticeSweep  # noqa: E402

n, w = int(sys.argv[1]), int(sys.argv[2])
s = LatticeSweep("potts/neighonly", "sc/26n", n, n, n, nspins=10000, temperature=0.5, seed=12345, device=0)
s.randomize(12345, 1, 10000)
s.sweeps(0, w)
s.sync()
s.timer_start()
s.sweeps(w, 1)
print("sweep %d: %.3f ms, variant %s" % (w, s.timer_stop(), s.neigh_variant(True)))
