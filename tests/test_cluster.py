"""diag_style cluster: the oracle against fixtures written by the UNMODIFIED
reference's own diagnostic (tests/golden/make_cluster_golden.py), and the device implementation
(spk_cuda_cluster_stats / spk_cuda_cluster_ids) against both."""
import glob
import os

import numpy as np
import pytest

import oracle_py as O
from test_oracle_golden import GOLDEN

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "cluster_*.npz")))


def _lattice(path):
    stem = os.path.basename(path)[len("cluster_"):].rsplit("_", 1)[0]
    return stem.replace("sq_", "sq/").replace("sc_", "sc/")


@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_oracle_clusters_match_reference(path):
    """ncluster, the volume and the spin value of every cluster in the reference's numbering, and
    the <N>, <R> columns to the 6 digits the reference prints."""
    g = np.load(path)
    n = [int(v) for v in g["n"]]
    o = O.Oracle("potts", _lattice(path), n[0], n[1], n[2], int(g["q"]), 0.5)
    nc, mv, mr, ids, vol = o.clusters(g["spins"])
    assert nc == int(g["nclust"])
    assert np.array_equal(vol, g["table"][:, 2])
    first = np.array([np.flatnonzero(ids == k + 1)[0] for k in range(nc)])
    assert np.array_equal(g["spins"][first], g["table"][:, 1])
    assert float("%g" % mv) == float(g["mean_volume"]) and float("%g" % mr) == float(g["mean_radius"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_device_clusters_match_reference(path):
    from spparks_b200 import LatticeSweep
    g = np.load(path)
    n = [int(v) for v in g["n"]]
    s = LatticeSweep("potts", _lattice(path), n[0], n[1], n[2], nspins=int(g["q"]), temperature=0.5)
    s.upload(g["spins"])
    nc, mv, mr = s.cluster_stats()
    assert nc == int(g["nclust"])
    assert float("%g" % mv) == float(g["mean_volume"]) and float("%g" % mr) == float(g["mean_radius"])
    ids, vol = s.cluster_ids()
    assert np.array_equal(vol, g["table"][:, 2])
    o = O.Oracle("potts", _lattice(path), n[0], n[1], n[2], int(g["q"]), 0.5)
    assert np.array_equal(ids, o.clusters(g["spins"])[3])  # same numbering as the reference's scan


@pytest.mark.gpu
@pytest.mark.parametrize("lat,n,q", [("sc/26n", (48, 40, 36), 6), ("sc/6n", (40, 32, 48), 3), ("tri", (96, 80, 1), 4),
                                     ("sq/8n", (128, 96, 1), 5), ("sq/4n", (200, 120, 1), 2)])
def test_device_clusters_after_coarsening_vs_oracle(lat, n, q):
    """a lattice coarsened by a few device sweeps (large, wrapped, irregular clusters): ids, volumes
    and the stats columns against the oracle (the sum of radii is accumulated in a different order
    on the device: relative tolerance 1e-12)."""
    from spparks_b200 import LatticeSweep
    s = LatticeSweep("potts", lat, n[0], n[1], n[2], nspins=q, temperature=0.3, seed=5)
    s.randomize(99, 1, q)
    s.sweeps(0, 12)
    spins = s.download()
    o = O.Oracle("potts", lat, n[0], n[1], n[2], q, 0.3)
    nc, mv, mr, ids, vol = o.clusters(spins)
    got = s.cluster_stats()
    assert got[0] == nc and got[1] == mv and abs(got[2] - mr) <= 1e-12 * mr
    dids, dvol = s.cluster_ids()
    assert np.array_equal(dvol, vol) and np.array_equal(dids, ids)


@pytest.mark.gpu
def test_device_clusters_properties_large():
    """256^3 sc/26n: one cluster for a uniform lattice; volumes always sum to the number of sites;
    a two-value stripe pattern gives exactly two wrapped slabs."""
    from spparks_b200 import LatticeSweep
    L = 256
    s = LatticeSweep("potts", "sc/26n", L, L, L, nspins=100, temperature=0.0, seed=1)
    s.upload(np.full(L ** 3, 7, dtype=np.int32))
    assert s.cluster_stats() == (1, float(L ** 3), float(L ** 3) ** (1.0 / 3.0))
    a = np.ones((L, L, L), dtype=np.int32)
    a[L // 4: 3 * L // 4] = 2
    s.upload(a.reshape(-1))
    nc, mv, _ = s.cluster_stats()
    assert nc == 2 and mv == L ** 3 / 2
    s.randomize(3, 1, 100)
    nc, mv, _ = s.cluster_stats()
    ids, vol = s.cluster_ids()
    assert len(vol) == nc and int(vol.sum()) == L ** 3 and ids.min() == 1 and ids.max() == nc
    assert np.array_equal(np.bincount(ids, minlength=nc + 1)[1:], vol)
