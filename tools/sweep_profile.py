#!/usr/bin/env python
"""Any bench workload: W sweeps then 3 timed ones (for ncu --launch-skip).  tools/sweep_profile.py WORKLOAD W"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spparks_b200 import LatticeSweep  # noqa: E402

w = bench.WORKLOADS[sys.argv[1]]
warm = int(sys.argv[2])
n = w["n"]
s = LatticeSweep(w["model"], w["lattice"], n[0], n[1], n[2], nspins=w["q"], temperature=w["T"], seed=12345, device=0)
s.randomize(12345, 1, 2 if w["model"] == "ising" else w["q"])
if w["pin"] > 0:
    s.pin_random(777, w["pin"])
s.sweeps(0, warm)
s.sync()
s.timer_start()
s.sweeps(warm, 3)
ms = s.timer_stop() / 3
print("%s after %d sweeps: %.3f ms/sweep %.1f Gsites/s" % (sys.argv[1], warm, ms, s.nlocal / ms / 1e6))
