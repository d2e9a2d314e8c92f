"""Device-side `pin fraction multi nthresh`: spk_cuda_pin_create against the pinned
sites the reference's own AppPottsPin::pin_create produced (tests/golden/pins_*.npz, written through
oracle/ref/ref_driver by tests/golden/make_golden.py pins) and against the oracle's restatement, which
the same fixtures pin (tests/test_oracle_golden.py).  (The file sorts last on purpose: it is the
newest code.)"""
import glob
import os

import numpy as np
import pytest

import oracle_py as O
from test_oracle_golden import GOLDEN

pytestmark = pytest.mark.gpu


def _park_seed(master_seed, draws_before):
    """state of `new RandomPark(ranmaster->uniform())` (src/app_potts_pin.cpp:248) after `draws_before`
    earlier commands took their draw from the master generator"""
    m = O.mars(master_seed)
    for _ in range(draws_before):
        O.lib().orc_mars_uniform(O.C.byref(m))
    return int(O.lib().orc_mars_uniform(O.C.byref(m)) * O.IM) or 1


def _sweep(lat, n, q, **kw):
    from spparks_b200 import LatticeSweep
    return LatticeSweep("potts/pin", lat, int(n[0]), int(n[1]), int(n[2]), nspins=int(q), temperature=0.5, **kw)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "pins_*.npz"))), ids=os.path.basename)
def test_device_pin_equals_the_reference(path):
    d = np.load(path)
    n, q, lat = [int(x) for x in d["n"]], int(d["q"]), str(d["lattice"])
    s = _sweep(lat, n, q)
    s.upload(d["s0"])
    npin = s.pin_create(_park_seed(int(d["seed"]), 1), float(d["fraction"]), int(d["multi"]), int(d["nthresh"]))
    assert np.array_equal(s.download(), d["s1"])
    assert npin == int((d["s1"] > q).sum())
    s.close()


@pytest.mark.parametrize("lat,n,q,fraction,multi,nthresh,prepinned", [
    ("sc/26n", (16, 12, 10), 30, 0.02, 0, 0, 0.0),
    ("sc/26n", (16, 12, 10), 30, 0.5, 0, 0, 0.0),      # many rounds
    ("sc/26n", (16, 12, 10), 30, 0.1, 0, 0, 0.3),      # pins already there: one round of ndesired attempts, then the recount
    ("sq/4n", (64, 48, 1), 4, 0.2, 0, 2, 0.0),
    ("tri", (20, 14, 1), 3, 0.15, 0, 3, 0.05),
    ("sq/8n", (30, 20, 1), 6, 0.25, 1, 0, 0.0),
    ("sc/6n", (12, 10, 8), 3, 0.05, 1, 3, 0.0),
    ("sc/26n", (8, 8, 8), 4, 0.0, 0, 0, 0.0),          # nothing to do
    ("sq/4n", (2, 2, 1), 3, 0.5, 1, 0, 0.0),           # smallest lattice: every neighbour is a periodic image
], ids=lambda v: str(v).replace(" ", ""))
def test_device_pin_equals_the_oracle(lat, n, q, fraction, multi, nthresh, prepinned):
    o = O.Oracle("potts/pin", lat, n[0], n[1], n[2], q, 0.5)
    rng = np.random.default_rng(17)
    for master in (4711, 20261017):
        spins = np.repeat(rng.integers(1, q + 1, size=(o.nsites + 5) // 6, dtype=np.int32), 6)[:o.nsites].copy()
        spins[rng.random(o.nsites) < prepinned] = q + 1
        want = spins.copy()
        npin = o.pin_create(want, fraction, multi, nthresh, O.mars(master))
        s = _sweep(lat, n, q)
        s.upload(spins)
        got_n = s.pin_create(_park_seed(master, 0), fraction, multi, nthresh)
        assert np.array_equal(s.download(), want), (lat, master)
        assert got_n == int((want > q).sum())
        if multi or fraction > 0:
            assert got_n == npin or prepinned > 0
        # the pinned lattice is a valid potts/pin lattice: a sweep runs and the pins stay (src/app_potts_pin.cpp:97)
        s.sweeps(0, 1)
        assert np.array_equal(s.download() > q, want > q)
        s.close()


def test_device_pin_errors():
    from spparks_b200 import LatticeSweep, SpkCudaError
    p = LatticeSweep("potts", "sc/26n", 8, 8, 8, nspins=5)
    with pytest.raises(SpkCudaError, match="only defined for app_style potts/pin"):
        p.pin_create(1, 0.1)
    s = _sweep("sq/4n", (16, 16, 1), 3)
    s.upload(np.ones(256, dtype=np.int32))
    for bad in [(-0.1, 0, 0), (1.5, 0, 0), (0.1, 2, 0), (0.1, 0, -1)]:
        with pytest.raises(SpkCudaError, match="Illegal pin command"):
            s.pin_create(1, *bad)
    with pytest.raises(SpkCudaError, match="seed"):
        s.pin_create(0, 0.1)
    # a uniform lattice has no grain boundary: with nthresh > 0 the reference would loop forever
    with pytest.raises(SpkCudaError, match="no attempt of the last 64 rounds succeeded"):
        s.pin_create(7, 0.1, 0, 1)
    assert (s.download() == 1).all()
    # multi: once every site has a pinned neighbour nothing can be inserted any more
    with pytest.raises(SpkCudaError, match="no site left"):
        s.pin_create(7, 0.9, 1, 0)
    s.close()


def test_device_pin_large_draw_numbers():
    """draw numbers far along the stream and several rounds on a 2M-site lattice, against the sequential
    stream walked on the host (tagrandom, src/random_park.cpp:103-108)"""
    n, q, fraction, seed = (128, 128, 128), 100, 0.3, 987654321
    N = n[0] * n[1] * n[2]
    s = _sweep("sc/26n", n, q)
    s.upload(np.ones(N, dtype=np.int32))
    npin = s.pin_create(seed, fraction)
    got = s.download() > q
    want = np.zeros(N, dtype=bool)
    state, ndesired, count = seed, int(fraction * N), 0
    am = 1.0 / O.IM
    while count < ndesired:
        for _ in range(ndesired - count):
            state = state * 16807 % O.IM
            want[min(int(state * am * N), N - 1)] = True
        count = int(want.sum())
    assert np.array_equal(got, want) and npin == count
    s.close()


@pytest.mark.parametrize("pin", ["0.05 0 0", "0.04 0 8", "0.1 1 0"])
def test_cuda_pin_script_identical_to_reference(pin):
    """`cuda_pin f multi nthresh` in a potts/pin/cuda script pins the sites the reference's `pin f multi nthresh`
    pins: the whole color/strict table (time-0 energy included) of the reference run is reproduced, and the draws
    that follow from the master generator (the strict site seeds) are unshifted."""
    import test_plugin as P
    if not (os.path.exists(P.SPK_CUDA) and os.path.exists(P.SPK_REF)):
        pytest.skip("plugin/_build/spk_cuda or oracle/_ref/spk_ref not built")
    case = dict(P.STRICT_CASES[2], extra="pin " + pin)
    assert case["app"] == "potts/pin"
    ref_script = P.STRICT.format(**dict(case, app="%s %d" % (case["app"], case["q"])))
    gpu_script = P.STRICT.format(**dict(case, app="%s/cuda %d" % (case["app"], case["q"]))).replace("\npin ", "\ncuda_pin ")
    assert "cuda_pin " + pin in gpu_script
    ref = P.stats_rows(P.run(P.SPK_REF, ref_script))
    assert P.stats_rows(P.run(P.SPK_CUDA, gpu_script)) == ref and len(ref) >= 4
