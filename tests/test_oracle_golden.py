"""The CPU oracle (oracle/spk_oracle.c) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py -> oracle/_ref/ref_driver) and against the reference's shipped logs."""
import glob
import os

import numpy as np
import pytest

import oracle_py as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAT = {"sq_4n": "sq/4n", "sq_8n": "sq/8n", "tri": "tri", "sc_6n": "sc/6n", "sc_26n": "sc/26n"}
MODEL = {"ising": "ising", "potts": "potts", "potts_neighonly": "potts/neighonly", "potts_pin": "potts/pin",
         "ising_single": "ising/single", "potts_neigh": "potts/neigh"}


def split_name(path, prefix):
    stem = os.path.basename(path)[len(prefix) + 1:-4]
    for mk in sorted(MODEL, key=len, reverse=True):
        if stem.startswith(mk + "_"):
            rest = stem[len(mk) + 1:]
            for lk in sorted(LAT, key=len, reverse=True):
                if rest.startswith(lk):
                    return MODEL[mk], LAT[lk]
    raise ValueError(path)


@pytest.mark.parametrize("lat", sorted(LAT))
def test_geometry_matches_reference(lat):
    """neighbour order (src/create_sites.cpp:1303-1371), colours (src/lattice.cpp:234-283) and
    sector/colour sets (src/app_lattice.cpp:954-1026) of the reference."""
    g = np.load(os.path.join(GOLDEN, "geom_%s.npz" % lat))
    nx, ny, nz = [int(v) for v in g["n"]]
    o = O.Oracle("potts", LAT[lat], nx, ny, nz, 7, 0.5)
    assert o.nsites == g["neigh"].shape[0] and o.maxneigh == g["neigh"].shape[1]
    assert np.array_equal(o.neighbor_table(), g["neigh"].astype(np.int64))
    assert np.array_equal(o.colors(), g["colors"])
    sets = g["sets"]
    nset, nsector, ncolors = [int(v) for v in sets[:3]]
    assert ncolors == o.ncolors and nset == nsector + nsector * ncolors
    colors, sectors = o.colors(), o.sectors(nsector)
    p = 3
    expect = [np.nonzero(sectors == s + 1)[0] for s in range(nsector)]
    expect += [np.nonzero((sectors == s + 1) & (colors == c + 1))[0] for s in range(nsector) for c in range(ncolors)]
    for e in expect:
        n = int(sets[p]); p += 1
        assert np.array_equal(sets[p:p + n], e.astype(np.int32))
        p += n
    # the sector as a cell box (what the CUDA ABI takes) selects the same sites
    nb = o.nbasis
    cell = np.arange(o.nsites) // nb
    i, j, k = cell % nx, (cell // nx) % ny, cell // (nx * ny)
    for s in range(nsector):
        b = o.sector_box(nsector, s + 1)
        inbox = (i >= b[0]) & (i < b[1]) & (j >= b[2]) & (j < b[3]) & (k >= b[4]) & (k < b[5])
        assert np.array_equal(inbox, sectors == s + 1)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "half_*.npz"))), ids=os.path.basename)
def test_injected_halfsweeps_match_reference(path):
    """site_event_rejection of the reference under injected uniforms, colour after colour."""
    g = np.load(path)
    model, lat = split_name(path, "half")
    nx, ny, nz = [int(v) for v in g["n"]]
    o = O.Oracle(model, lat, nx, ny, nz, int(g["q"]), float(g["T"]))
    spin = g["spins0"].copy()
    for c in range(o.ncolors):
        u1, u2 = O.injected_uniforms(g["k"][c])
        before = spin.copy()
        nacc = o.halfsweep_injected(spin, c + 1, u1, u2)
        assert np.array_equal(spin, g["after"][c]), "colour %d" % (c + 1)
        assert nacc == int(np.count_nonzero(spin != before))
    assert o.total_energy(spin) == float(g["energy"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "sets_*.npz"))), ids=os.path.basename)
def test_sector_color_pass_matches_reference(path):
    """iterate_rejection's sector -> colour order (src/app_lattice.cpp:684-691), sectors as boxes."""
    g = np.load(path)
    model, lat = split_name(path, "sets")
    nx, ny, nz = [int(v) for v in g["n"]]
    o = O.Oracle(model, lat, nx, ny, nz, int(g["q"]), float(g["T"]))
    nsector = 4 if nz == 1 else 8
    spin = g["spins0"].copy()
    u1, u2 = O.injected_uniforms(g["k"])
    for s in range(nsector):
        box = o.sector_box(nsector, s + 1)
        for c in range(o.ncolors):
            o.halfsweep_injected(spin, c + 1, u1, u2, box=box)
    assert np.array_equal(spin, g["after"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "masksets_*.npz"))), ids=os.path.basename)
def test_masked_sector_passes_match_reference(path):
    """`sweep color mask yes` with sectors (src/app_lattice.cpp:655, 784-792, 1158-1164 and the mask
    updates inside the events): spins AND mask bytes after every pass over the sets."""
    g = np.load(path)
    model, lat = split_name(path, "masksets")
    nx, ny, nz = [int(v) for v in g["n"]]
    o = O.Oracle(model, lat, nx, ny, nz, int(g["q"]), float(g["T"]))
    nsector = int(g["nsector"])
    spins = g["spins0"].copy()
    mask = np.zeros(o.nsites, dtype=np.uint8)
    for p in range(g["k"].shape[0]):
        u1, u2 = O.injected_uniforms(g["k"][p])
        for sec in range(nsector):
            box = o.sector_box(nsector, sec + 1)
            o.boundary_clear_mask(mask, box)
            for c in range(o.ncolors):
                o.halfsweep_injected_mask(spins, mask, c + 1, u1, u2, box=box)
        assert np.array_equal(spins, g["after"][p]), "spins after pass %d" % p
        assert np.array_equal(mask, g["mask"][p]), "mask after pass %d" % p
    assert g["mask"].any() and not g["mask"].all()  # the fixture exercises both skipped and live sites


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "strict_*.npz"))), ids=os.path.basename)
def test_strict_runs_match_reference(path):
    """`sweep color/strict` full runs: spins, per-site seeds, Naccept and energy."""
    g = np.load(path)
    model, lat = split_name(path, "strict")
    nx, ny, nz = [int(v) for v in g["n"]]
    o = O.Oracle(model, lat, nx, ny, nz, int(g["q"]), float(g["T"]))
    # per-site seeds from the master RNG: `set site range` consumed one master uniform,
    # AppLattice::init two for ranapp (src/app_lattice.cpp:274-278), then two for ranstrict
    m = O.mars(int(g["master_seed"]))
    O.lib().orc_mars_uniform(m)
    O.lib().orc_mars_uniform(m)
    O.lib().orc_mars_uniform(m)
    assert np.array_equal(o.strict_siteseeds(m), g["seeds0"])
    spin, seeds = g["spins0"].copy(), g["seeds0"].copy()
    nacc = o.sweeps_strict(spin, seeds, int(g["nsweeps"]))
    assert np.array_equal(spin, g["after"])
    assert np.array_equal(seeds, g["seeds1"])
    assert nacc == int(g["naccept"])
    assert o.total_energy(spin) == float(g["energy"])


def _log_rows(text):
    rows, on = [], False
    for line in text.splitlines():
        f = line.split()
        if f[:2] == ["Time", "Naccept"]:
            on = True
            continue
        if on:
            if len(f) == 6 and f[0][0].isdigit():
                rows.append((float(f[0]), int(f[1]), int(f[2]), int(f[3]), float(f[5])))
            else:
                break
    return rows


# stats tables of the reference's shipped logs (values only; examples/README:10-16 names them as
# the answers to compare to).  examples/ising/log.ising.11Nov09.linux.1:33-44
ISING_LOG = """
      Time    Naccept    Nreject    Nsweeps        CPU     Energy
         0          0          0          0          0       5112
        10       4332      45668         20    0.00687        816
        20       5683      94317         40     0.0136        576
        30       6762     143238         60     0.0204        480
        40       7710     192290         80     0.0271        420
        50       8508     241492        100     0.0338        400
        60       9209     290791        120     0.0406        384
        70       9996     340004        140     0.0474        380
        80      10706     389294        160     0.0541        356
        90      11416     438584        180     0.0608        340
       100      12214     487786        200     0.0674        320
"""
# examples/potts/log.potts.11Nov09.linux.1:34-45
POTTS_LOG = """
      Time    Naccept    Nreject    Nsweeps        CPU     Energy
         0          0          0          0          0     205912
     10.01      88984    7919016       1001      0.994      71310
        20      95018   15904982       2000       1.96      56802
        30      98551   23901449       3000       2.92      49940
        40     101484   31898516       4000       3.89      43924
     50.01     104024   39903976       5001       4.85      38658
     60.01     105943   47902057       6001       5.81      34696
     70.01     107446   55900554       7001       6.78      31620
        80     108635   63891365       8000       7.74      29556
        90     109733   71890267       9000        8.7      26986
       100     110585   79889415      10000       9.67      25382
"""


def _check_rows(rows, log):
    want = _log_rows(log)
    assert len(rows) == len(want)
    for got, w in zip(rows, want):
        assert float("%g" % got[0]) == w[0]
        assert got[1:4] == w[1:4]
        assert got[4] == w[4]


def test_golden_log_in_ising():
    """examples/ising/in.ising (seed 56789, sq/4n 50x50, sweep random, sector yes, T=0, run 100)."""
    m = O.mars(56789)
    o = O.Oracle("ising", "sq/4n", 50, 50, 1, 2, 0.0)
    spin = o.set_site_range(1, 2, m)
    _check_rows(o.run_sweep_random(spin, m, True, 10.0, 100.0), ISING_LOG)


def test_golden_log_in_potts_first_rows():
    """examples/potts/in.potts (potts 100, sc/26n 20^3, sweep random, sector yes, T=0): the first
    two stats intervals (2000 sweeps, 16 M events); the full table is checked in the slow test."""
    m = O.mars(56789)
    o = O.Oracle("potts", "sc/26n", 20, 20, 20, 100, 0.0)
    spin = o.set_site_range(1, 100, m)
    rows = o.run_sweep_random(spin, m, True, 10.0, 20.0)
    want = _log_rows(POTTS_LOG)[:3]
    assert [r[1:] for r in rows] == [w[1:] for w in want]


@pytest.mark.slow
def test_golden_log_in_potts_full():
    m = O.mars(56789)
    o = O.Oracle("potts", "sc/26n", 20, 20, 20, 100, 0.0)
    spin = o.set_site_range(1, 100, m)
    _check_rows(o.run_sweep_random(spin, m, True, 10.0, 100.0), POTTS_LOG)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "pins_*.npz"))), ids=os.path.basename)
def test_oracle_pin_create_equals_the_reference(path):
    """AppPottsPin::pin_create run by the reference (`pin fraction multi nthresh` after `set site range`,
    spins loaded in between): single sites anywhere, at grain boundaries only, site + neighbours"""
    d = np.load(path)
    n, q = [int(x) for x in d["n"]], int(d["q"])
    o = O.Oracle("potts/pin", str(d["lattice"]), n[0], n[1], n[2], q, 0.5)
    m = O.mars(int(d["seed"]))
    o.set_site_range(1, q, m)  # the master generator's first draw went to `set site range`
    s = d["s0"].copy()
    npin = o.pin_create(s, float(d["fraction"]), int(d["multi"]), int(d["nthresh"]), m)
    assert np.array_equal(s, d["s1"])
    assert npin == int((d["s1"] > q).sum())
