"""The DEVICE source of the per-site events, compiled for the host, against the oracle (CPU only).

spparks_b200/csrc/event.cuh holds everything the kernels compute per site: Philox4x32-10, the
all-pairs event with 32-bit compares and with packed 16-bit compares (VIADDMNMX.U16x2 encoding),
candidate elimination for sc/26n (site_event_nb26 / site_event_nb26x2), the injected-uniform and
`sweep color/strict` double paths.  tests/host_shim/ compiles that header with g++ (CUDA
qualifiers and the few integer intrinsics restated in cuda_shim.h; the two inline-PTX spots have a
plain C++ twin under SPK_HOST_SHIM whose SASS-neutrality was checked when it was added) and this
file compares it with the oracle event by event: same neighbourhood, same draws, same result.  It
does not replace the GPU parity tests (the kernels' addressing, halos and word assignment are not in
event.cuh); it pins the arithmetic the kernels share on a machine without a GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_DIR = os.path.join(ROOT, "tests", "host_shim")
SHIM_SO = os.path.join(SHIM_DIR, "libevent_host.so")
CSRC = os.path.join(ROOT, "spparks_b200", "csrc")
M32 = 1 << 32

LATS = {"sq/4n": (8, 8, 1), "sq/8n": (8, 8, 1), "tri": (8, 8, 1), "sc/6n": (4, 4, 4), "sc/26n": (4, 4, 4)}
MODELS = ["ising", "ising/single", "potts", "potts/pin", "potts/neighonly", "potts/neigh"]


@pytest.fixture(scope="module")
def shim():
    srcs = [os.path.join(SHIM_DIR, f) for f in ("event_host.cpp", "cuda_shim.h")]
    srcs += [os.path.join(CSRC, f) for f in ("event.cuh", "dedupe_chain.cuh")]
    if not os.path.exists(SHIM_SO) or any(os.path.getmtime(s) > os.path.getmtime(SHIM_SO) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"),
                               "-I" + CSRC, srcs[0], "-o", SHIM_SO])
    L = C.CDLL(SHIM_SO)
    ip = C.POINTER(C.c_int)
    u32p = C.POINTER(C.c_uint32)
    L.shim_site_event.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, ip, C.c_uint32, C.c_uint32]
    L.shim_site_event_injected.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, ip, C.c_double, C.c_double]
    L.shim_site_event_strict.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, ip, ip]
    L.shim_site_event_nb26x2.argtypes = [C.c_int, C.c_int, C.c_double, ip, ip, ip, u32p]
    L.shim_philox.argtypes = [u32p, u32p, u32p, C.c_int]
    return L


def _lattice(model, lattice, Q, temperature, seed, pinned=0.0):
    nx, ny, nz = LATS[lattice]
    o = orc.Oracle(model, lattice, nx, ny, nz, nspins=Q, temperature=temperature)
    rng = np.random.default_rng(seed)
    spins = rng.integers(1, o.cfg.nspins + 1, size=o.nsites).astype(np.int32)
    if pinned:
        spins[rng.random(o.nsites) < pinned] = o.cfg.nspins + 1
    return o, spins, rng


def _words(rng, o, temperature):
    """words that decide something: range ends, the trial boundaries, the thresholds +-1, random ones"""
    ws = {0, 1, M32 - 1, M32 - 2, 1 << 31, (1 << 31) - 1}
    if temperature > 0:
        for dE in range(1, o.maxneigh + 1):
            t = orc.accept_threshold(temperature, dE)
            ws |= {t, (t + 1) % M32, (t - 1) % M32}
    ws |= {int(x) for x in rng.integers(0, M32, size=12)}
    return sorted(ws)


def _iv(vals):
    return (C.c_int * len(vals))(*[int(x) for x in vals])


def test_philox_of_the_device_source(shim):
    rng = np.random.default_rng(5)
    for _ in range(200):
        ctr = [int(x) for x in rng.integers(0, M32, size=4)]
        key = [int(x) for x in rng.integers(0, M32, size=2)]
        want = orc.philox(ctr, key)
        for pre in (0, 1):
            out = (C.c_uint32 * 4)()
            shim.shim_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out, pre)
            assert list(out) == want, (ctr, key, pre)


@pytest.mark.parametrize("lattice", list(LATS))
@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("Q,temperature", [(3, 0.0), (5, 0.7), (60, 1.9)])
def test_word_events(shim, model, lattice, Q, temperature):
    """Philox-mode event: 32-bit chain, packed 16-bit chain and (sc/26n candidate-list apps) candidate
    elimination == the oracle's integer event for the same two words."""
    o, spins, rng = _lattice(model, lattice, Q, temperature, 11, pinned=0.1 if model == "potts/pin" else 0.0)
    nn = o.maxneigh
    variants = [0, 1] + ([2] if nn == 26 and model in ("potts/neighonly", "potts/neigh") else [])
    w0s = _words(rng, o, temperature)
    w1s = w0s if len(w0s) < 20 else w0s[:: max(1, len(w0s) // 20)] + [w0s[-1]]
    sites = rng.choice(o.nsites, size=6, replace=False)
    if nn == 26:  # neighbourhoods with few distinct values, the case candidate elimination is for
        for s in sites[:3]:
            nb = o.neighbors(int(s))
            spins[nb] = rng.integers(1, 3, size=nn)
    checked = 0
    for site in map(int, sites):
        v = _iv(spins[o.neighbors(site)])
        old = int(spins[site])
        for w0 in w0s:
            for w1 in w1s:
                if model.startswith("ising"):
                    w1 = (w0 << 1) % M32  # event.cuh: philox_words
                a = spins.copy()
                o.site_event_words(a, site, w0, w1)
                for variant in variants:
                    got = shim.shim_site_event(o.model, nn, variant, o.cfg.nspins, temperature, old, v, w0, w1)
                    assert got == a[site], (site, w0, w1, variant, old, list(v))
                checked += 1
                if model.startswith("ising"):
                    break
    assert checked >= 60


@pytest.mark.parametrize("model", ["potts/neighonly", "potts/neigh"])
@pytest.mark.parametrize("Q,temperature", [(2, 0.0), (4, 0.7), (1000, 1.9)])
def test_candidate_elimination_pair_form(shim, model, Q, temperature):
    """site_event_nb26x2 (two sites share one Philox block) == two oracle events"""
    o, spins, rng = _lattice(model, "sc/26n", Q, temperature, 23)
    for trial in range(300):
        s0, s1 = (int(x) for x in rng.choice(o.nsites, size=2, replace=False))
        blk = [int(x) for x in rng.integers(0, M32, size=4)]
        if trial % 3 == 0 and temperature > 0:
            blk[1] = orc.accept_threshold(temperature, 1 + trial % 5) + (trial % 2)
        v0, v1 = spins[o.neighbors(s0)], spins[o.neighbors(s1)]
        a, b = spins.copy(), spins.copy()
        o.site_event_words(a, s0, blk[0], blk[1])
        o.site_event_words(b, s1, blk[2], blk[3])
        cur = _iv([spins[s0], spins[s1]])
        assert shim.shim_site_event_nb26x2(o.model, o.cfg.nspins, temperature, cur, _iv(v0), _iv(v1), (C.c_uint32 * 4)(*blk)) == 0
        assert (cur[0], cur[1]) == (a[s0], b[s1]), (s0, s1, blk)


@pytest.mark.parametrize("lattice", list(LATS))
@pytest.mark.parametrize("model", MODELS)
def test_injected_events(shim, model, lattice):
    """injected-uniform mode (the reference-pinned double formulas): one colour set of the device
    event == orc_halfsweep_injected"""
    Q, temperature = 7, 1.1
    o, spins, rng = _lattice(model, lattice, Q, temperature, 31, pinned=0.1 if model == "potts/pin" else 0.0)
    colors = o.colors()
    u1, u2 = rng.random(o.nsites), rng.random(o.nsites)
    u1[::7] = 0.0
    u2[::5] = 0.0
    u1[3::11] = np.nextafter(1.0, 0.0)
    for color in range(1, o.ncolors + 1):
        want = spins.copy()
        o.halfsweep_injected(want, color, u1, u2)
        got = spins.copy()
        for site in np.flatnonzero(colors == color):
            site = int(site)
            got[site] = shim.shim_site_event_injected(o.model, o.maxneigh, o.cfg.nspins, temperature, int(spins[site]),
                                                      _iv(spins[o.neighbors(site)]), u1[site], u2[site])
        assert np.array_equal(got, want), (model, lattice, color)
        spins = want


@pytest.mark.parametrize("lattice", ["sq/4n", "sc/26n"])
@pytest.mark.parametrize("model", MODELS)
def test_strict_events(shim, model, lattice):
    """`sweep color/strict`: per-site Park-Miller streams of the device event == orc_sweeps_strict"""
    Q, temperature = 6, 0.9
    o, spins, rng = _lattice(model, lattice, Q, temperature, 41, pinned=0.1 if model == "potts/pin" else 0.0)
    seeds = o.strict_siteseeds(orc.mars(12345))
    want, want_seeds = spins.copy(), seeds.copy()
    o.sweeps_strict(want, want_seeds, 2)
    colors = o.colors()
    got, got_seeds = spins.copy(), seeds.copy()
    for _ in range(2):
        for color in range(1, o.ncolors + 1):
            for site in map(int, np.flatnonzero(colors == color)):
                sd = C.c_int(int(got_seeds[site]))
                got[site] = shim.shim_site_event_strict(o.model, o.maxneigh, o.cfg.nspins, temperature, int(got[site]),
                                                        _iv(got[o.neighbors(site)]), C.byref(sd))
                got_seeds[site] = sd.value
    assert np.array_equal(got, want)
    assert np.array_equal(got_seeds, want_seeds)
