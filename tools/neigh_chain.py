#!/usr/bin/env python
"""potts/neighonly sc/26n N^3, chain mode: ms/sweep over sweeps 5..25 (the driver's bench window)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SPK_CUDA_NEIGH_MODE", "chain")
from spparks_b200 import LatticeSweep  # noqa: E402

n = int(sys.argv[1])
s = LatticeSweep("potts/neighonly", "sc/26n", n, n, n, nspins=10000, temperature=0.5, seed=12345, device=0)
s.randomize(12345, 1, 10000)
s.sweeps(0, 5)
s.sync()
s.timer_start()
s.sweeps(5, 20)
ms = s.timer_stop() / 20
print("chain8=%s lib=%s: %.3f ms/sweep %.1f Gsites/s" % (os.environ.get("SPK_CUDA_CHAIN8"), os.environ.get("SPK_CUDA_LIB"), ms, n ** 3 / ms / 1e6))
