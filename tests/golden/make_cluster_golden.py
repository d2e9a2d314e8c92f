#!/usr/bin/env python
"""Cluster-analysis fixtures from the UNMODIFIED reference (`diag_style cluster`, src/diag_cluster.cpp)
through oracle/_ref/ref_driver: tests/golden/cluster_<lattice>_<tag>.npz = spins, Nclust, <N>, <R>
and the per-cluster (id, ivalue, size) table, all from the diagnostic's own output file (%g format).  Run in the build container:  python tests/golden/make_cluster_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import DIM, header, nsites, run_driver, tag  # noqa: E402

CASES = [
    # lattice, n, Q, how the spins are made
    ("sq/4n", (24, 20, 1), 3, "random"),
    ("sq/8n", (16, 18, 1), 4, "random"),
    ("tri", (14, 10, 1), 3, "random"),
    ("sc/6n", (10, 8, 12), 3, "random"),
    ("sc/26n", (12, 10, 8), 12, "random"),
    ("sc/26n", (16, 16, 16), 40, "blocky"),   # grains of 4^3 cells: clusters wrap the periodic box
    ("sq/4n", (32, 32, 1), 2, "stripes"),     # two stripes that wrap in y
]


def make_spins(rng, kind, lattice, n, q):
    ns = nsites(lattice, n)
    if kind == "random":
        return rng.integers(1, q + 1, size=ns, dtype=np.int32)
    nb = ns // (n[0] * n[1] * (n[2] if DIM[lattice] == 3 else 1))
    i = np.arange(n[0])[None, None, :, None]
    j = np.arange(n[1])[None, :, None, None]
    k = np.arange(n[2] if DIM[lattice] == 3 else 1)[:, None, None, None]
    if kind == "blocky":
        grains = rng.integers(1, q + 1, size=(n[2] // 4 + 1, n[1] // 4 + 1, n[0] // 4 + 1), dtype=np.int32)
        a = grains[(k + 1) // 4 % grains.shape[0], (j + 2) // 4 % grains.shape[1], (i + 3) // 4 % grains.shape[2]]
        return np.broadcast_to(a, (k.shape[0], n[1], n[0], nb)).reshape(-1).astype(np.int32)
    a = 1 + ((i // 8) % 2) + 0 * j + 0 * k
    return np.broadcast_to(a, (k.shape[0], n[1], n[0], nb)).reshape(-1).astype(np.int32)


def main():
    rng = np.random.default_rng(20261017)
    for lattice, n, q, kind in CASES:
        spins = make_spins(rng, kind, lattice, n, q)
        with tempfile.TemporaryDirectory() as tmp:
            spins.tofile(os.path.join(tmp, "spins.bin"))
            lines = header("potts", q, lattice, n) + [
                "sweep color", "sector no", "temperature 0.5", "run 0.0", "@load spins.bin",
                "diag_style cluster stats yes filename clusters.txt", "stats 1.0", "run 0.0"]
            run_driver(lines, tmp)
            row, table, on = {}, [], False
            for line in open(os.path.join(tmp, "clusters.txt")):
                f = line.split()
                if len(f) == 3 and f[1] == "=" and f[0] in ("ncluster", "<N>", "<R>"):
                    row[f[0]] = f[2]
                if f[:2] == ["id", "ivalue"]:
                    on = True
                    continue
                if on and len(f) >= 4:
                    table.append((int(f[0]), int(f[1]), int(float(f[3]))))
                elif on:
                    break
            row = {"Nclust": row["ncluster"], "<N>": row["<N>"], "<R>": row["<R>"]}
        table = np.array(table, dtype=np.int64)
        assert len(table) == int(row["Nclust"]) and table[:, 2].sum() == spins.size
        name = "cluster_%s_%s.npz" % (tag(lattice), kind)
        np.savez_compressed(os.path.join(HERE, name), n=np.array(n, dtype=np.int32), q=q, spins=spins,
                            nclust=int(row["Nclust"]), mean_volume=float(row["<N>"]), mean_radius=float(row["<R>"]),
                            table=table)
        print(name, row["Nclust"], row["<N>"], row["<R>"])


if __name__ == "__main__":
    main()
