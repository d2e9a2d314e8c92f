#!/usr/bin/env python
"""Grain-growth reference curve from the UNMODIFIED reference (oracle/_ref/spk_ref).

potts/neighonly Q=10000, sc/26n 96^3, T=0.5, `sweep color`, `sector no`, random start, 60 sweeps,
`diag_style energy` + `diag_style cluster`, 8 seeds -> tests/golden/growth_reference.json
(mean over seeds and the seed-to-seed standard error).  Run in the build container:
    python tests/golden/make_growth_reference.py        (about 40 s on 8 cores)
"""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SPK = os.path.join(ROOT, "oracle", "_ref", "spk_ref")
L, SEEDS = 96, [11, 22, 33, 44, 55, 66, 77, 88]

SCRIPT = """seed {seed}
app_style potts/neighonly 10000
dimension 3
lattice sc/26n 1.0
region box block 0 {L} 0 {L} 0 {L}
create_box box
create_sites box
set site range 1 10000
sweep color
sector no
temperature 0.5
diag_style energy
diag_style cluster stats yes
stats 0.3846153846153846
run 2.3076
"""


def main():
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for seed in SEEDS:
            path = os.path.join(tmp, "in.%d" % seed)
            open(path, "w").write(SCRIPT.format(seed=seed, L=L))
            procs.append(subprocess.Popen([SPK, "-in", path, "-log", os.path.join(tmp, "log.%d" % seed),
                                           "-screen", "none"], cwd=tmp))
        for p in procs:
            p.wait()
        tables = []
        for seed in SEEDS:
            rows, on = [], False
            for line in open(os.path.join(tmp, "log.%d" % seed)):
                f = line.split()
                if f[:2] == ["Time", "Naccept"]:
                    on = True
                    continue
                if on:
                    if len(f) == 9:
                        rows.append((int(f[3]), float(f[5]), int(f[6])))
                    else:
                        break
            tables.append(rows)
    sweeps = [r[0] for r in tables[0]]
    assert all([r[0] for r in t] == sweeps for t in tables)
    n = L ** 3
    e = np.array([[r[1] / n for r in t] for t in tables])
    vol = np.array([[n / r[2] for r in t] for t in tables])
    out = {"lattice": "sc/26n", "L": L, "app_style": "potts/neighonly 10000", "temperature": 0.5, "seeds": SEEDS,
           "sweeps": sweeps[1:], "energy_per_site": e.mean(0)[1:].tolist(),
           "energy_stderr": (e.std(0, ddof=1) / np.sqrt(len(SEEDS)))[1:].tolist(),
           "mean_grain_volume": vol.mean(0)[1:].tolist(),
           "grain_volume_stderr": (vol.std(0, ddof=1) / np.sqrt(len(SEEDS)))[1:].tolist(),
           "cluster_checkpoints": [sweeps[3], sweeps[-1]]}
    json.dump(out, open(os.path.join(HERE, "growth_reference.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
