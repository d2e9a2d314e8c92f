"""Known-answer tests for the oracle's generators."""
import numpy as np
import pytest

import oracle_py as O


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors (philox4x32 10 rounds)
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_park_miller_ten_thousandth():
    # Park & Miller 1988: seed 1 -> after 10000 steps the state is 1043618065
    import ctypes as C

    class P(C.Structure):
        _fields_ = [("seed", C.c_int)]
    L = O.lib()
    L.orc_park_uniform.restype = C.c_double
    p = P(1)
    for _ in range(10000):
        L.orc_park_uniform(C.byref(p))
    assert p.seed == 1043618065


def test_injection_identity():
    # 16807^-1 mod (2^31-1) = 1407677000 (used by oracle/ref/ref_driver.cpp to inject uniforms)
    assert (16807 * 1407677000) % O.IM == 1


def test_accept_threshold_monotone():
    t = [O.accept_threshold(0.7, d) for d in range(1, 27)]
    assert all(a >= b for a, b in zip(t, t[1:])) and t[0] > t[5] > 0 and t[0] < 2 ** 32


# ---- the integer reduction of the Philox mode (oracle/spk_oracle.c above orc_accept_threshold) ----
# With u = w/2^32 the reference's double formulas ((int)(Q*u)+1, u < 0.5, (int)(nevent*u), u > exp(-dE/T))
# must give the same event as the integer forms the CUDA kernels evaluate.  orc_site_event (the double
# path) is pinned to the reference by the injected-uniform fixtures; here the same words go through both.
REDUCTION_CASES = [
    ("ising", "sq/4n", (12, 10, 1), 2), ("ising", "sc/6n", (8, 6, 6), 2), ("ising/single", "sq/8n", (10, 8, 1), 2),
    ("ising/single", "sc/26n", (6, 6, 4), 2), ("potts", "sq/4n", (12, 10, 1), 3), ("potts", "tri", (10, 8, 1), 1000),
    ("potts", "sc/26n", (6, 6, 6), 10000), ("potts", "sc/6n", (8, 6, 6), 65534), ("potts/pin", "sc/26n", (6, 6, 6), 50),
    ("potts/pin", "sq/8n", (10, 8, 1), 7), ("potts/neighonly", "sc/26n", (6, 6, 6), 10000),
    ("potts/neighonly", "sc/26n", (6, 6, 6), 4), ("potts/neighonly", "tri", (10, 8, 1), 5),
    ("potts/neigh", "sc/26n", (6, 6, 6), 6), ("potts/neigh", "sq/8n", (10, 8, 1), 3),
]


@pytest.mark.parametrize("model,lat,n,q", REDUCTION_CASES, ids=lambda v: str(v).replace(" ", ""))
@pytest.mark.parametrize("T", [0.0, 0.3, 1.1346, 7.5])
def test_philox_integer_reduction_equals_double_formulas(model, lat, n, q, T):
    """Whole Philox half-sweeps through the integer forms and through the double formulas."""
    rng = np.random.default_rng(17)
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
    hi = 2 if model.startswith("ising") else q
    a = rng.integers(1, min(hi, 6) + 1, size=o.nsites, dtype=np.int32)  # few values: many dE > 0 cases
    if model == "potts/pin":
        a[rng.random(o.nsites) < 0.1] = q + 1
    b = a.copy()
    for sweep in range(3):
        for c in range(1, o.ncolors + 1):
            na = o.halfsweep_philox(a, 99, sweep, c)
            nb = o.halfsweep_philox_double(b, 99, sweep, c)
            assert na == nb and np.array_equal(a, b), (sweep, c)


@pytest.mark.parametrize("model,lat,n,q", REDUCTION_CASES, ids=lambda v: str(v).replace(" ", ""))
@pytest.mark.parametrize("T", [0.0, 0.5, 2.0])
def test_philox_integer_reduction_at_the_thresholds(model, lat, n, q, T):
    """Words chosen on the decision boundaries: the acceptance word at thr(dE)-1, thr(dE), thr(dE)+1 for
    every dE, the trial word at the bin edges of (Q*w)>>32 / (nevent*w)>>32 and at 2^31."""
    rng = np.random.default_rng(23)
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
    hi = 2 if model.startswith("ising") else q
    spins = rng.integers(1, min(hi, 4) + 1, size=o.nsites, dtype=np.int32)
    if model == "potts/pin":
        spins[rng.random(o.nsites) < 0.1] = q + 1
    M = 1 << 32
    w1s = {0, 1, M - 1}
    if T > 0:
        for dE in range(1, o.maxneigh + 1):
            t = O.accept_threshold(T, dE)
            w1s.update(v for v in (t - 1, t, t + 1) if 0 <= v < M)
    bins = 2 if model.startswith("ising") else (q if model in ("potts", "potts/pin") else o.maxneigh)
    w0s = {0, 1, (M >> 1) - 1, M >> 1, M - 1}
    for k in sorted({1, 2, 3, bins // 2, bins - 1, bins}):
        for nb in ({bins} if model in ("potts", "potts/pin") else set(range(1, o.maxneigh + 1))):
            edge = -(-k * M // nb)  # smallest w with (nb*w)>>32 == k
            w0s.update(v for v in (edge - 1, edge, edge + 1) if 0 <= v < M)
    w0s, w1s = sorted(w0s), sorted(w1s)
    sites = rng.choice(o.nsites, size=min(o.nsites, 12), replace=False)
    for site in sites:
        for w0 in w0s:
            for w1 in (w1s if len(w0s) * len(w1s) < 4000 else w1s[::3]):
                if model.startswith("ising"):
                    w1 = (w0 << 1) & (M - 1) if model == "ising" else w1  # ising: the kernels derive w1 from w0
                a, b = spins.copy(), spins.copy()
                ra = o.site_event_words(a, int(site), w0, w1)
                rb = o.site_event_words(b, int(site), w0, w1, double=True)
                assert ra == rb and a[site] == b[site], (int(site), w0, w1)
