#!/usr/bin/env python
"""potts/pin sc/26n N^3: W sweeps then one more (ncu --launch-skip).  python tools/pin_profile.py N W [model]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spparks_b200 import LatticeSweep  # noqa: E402

n, w = int(sys.argv[1]), int(sys.argv[2])
model = sys.argv[3] if len(sys.argv) > 3 else "potts/pin"
s = LatticeSweep(model, "sc/26n", n, n, n, nspins=10000, temperature=0.5, seed=12345, device=0)
s.randomize(12345, 1, 10000)
if model == "potts/pin":
    s.pin_random(777, 0.01)
s.sweeps(0, w)
s.sync()
s.timer_start()
s.sweeps(w, 3)
ms = s.timer_stop() / 3
print("%s %d^3 after %d sweeps: %.3f ms/sweep %.1f Gsites/s" % (model, n, w, ms, n ** 3 / ms / 1e6))
