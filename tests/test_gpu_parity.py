"""GPU parity: the CUDA path (through the C ABI) against the reference fixtures and the oracle."""
import glob
import os

import numpy as np
import pytest

import oracle_py as O
from test_oracle_golden import GOLDEN, split_name

pytestmark = pytest.mark.gpu


def _sweep(model, lat, n, q, T, **kw):
    from spparks_b200 import LatticeSweep
    return LatticeSweep(model, lat, int(n[0]), int(n[1]), int(n[2]), nspins=int(q), temperature=float(T), **kw)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "half_*.npz"))), ids=os.path.basename)
def test_injected_halfsweeps_bit_exact_vs_reference(path):
    """Single colour half-sweeps with the reference's uniforms injected: integer spins bit-exact
    against the UNMODIFIED reference (fixtures from oracle/_ref/ref_driver)."""
    g = np.load(path)
    model, lat = split_name(path, "half")
    s = _sweep(model, lat, g["n"], g["q"], g["T"])
    s.upload(g["spins0"])
    prev = g["spins0"]
    nacc = 0
    for c in range(s.ncolors):
        u1, u2 = O.injected_uniforms(g["k"][c])
        s.sweep_color_injected(c + 1, u1, u2)
        got = s.download()
        assert np.array_equal(got, g["after"][c]), "colour %d" % (c + 1)
        nacc += int(np.count_nonzero(got != prev))
        prev = got
    assert s.counters()[0] == nacc
    assert s.energy() == float(g["energy"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "sets_*.npz"))), ids=os.path.basename)
def test_sector_color_sets_bit_exact_vs_reference(path):
    g = np.load(path)
    model, lat = split_name(path, "sets")
    nx, ny, nz = [int(v) for v in g["n"]]
    s = _sweep(model, lat, g["n"], g["q"], g["T"])
    o = O.Oracle(model, lat, nx, ny, nz, int(g["q"]), float(g["T"]))
    nsector = 4 if nz == 1 else 8
    s.upload(g["spins0"])
    u1, u2 = O.injected_uniforms(g["k"])
    for sec in range(nsector):
        box = o.sector_box(nsector, sec + 1)
        for c in range(s.ncolors):
            s.sweep_color_injected(c + 1, u1, u2, box=box)
    assert np.array_equal(s.download(), g["after"])
    assert s.counters()[1] == s.nlocal


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "masksets_*.npz"))), ids=os.path.basename)
def test_masked_sector_passes_bit_exact_vs_reference(path):
    """`sweep color mask yes` with sectors: spins and mask bytes after every pass over the sets
    against the UNMODIFIED reference (its own boundary_clear_mask, mask rule and events)."""
    g = np.load(path)
    model, lat = split_name(path, "masksets")
    nx, ny, nz = [int(v) for v in g["n"]]
    s = _sweep(model, lat, g["n"], g["q"], g["T"])
    o = O.Oracle(model, lat, nx, ny, nz, int(g["q"]), float(g["T"]))
    nsector = int(g["nsector"])
    s.upload(g["spins0"])
    s.mask_enable()
    nattempt = 0
    for p in range(g["k"].shape[0]):
        u1, u2 = O.injected_uniforms(g["k"][p])
        for sec in range(nsector):
            box = o.sector_box(nsector, sec + 1)
            s.mask_clear_border(box)
            for c in range(s.ncolors):
                s.sweep_color_injected(c + 1, u1, u2, box=box)
        assert np.array_equal(s.download(), g["after"][p]), "spins after pass %d" % p
        assert np.array_equal(s.download_mask(), g["mask"][p]), "mask after pass %d" % p


@pytest.mark.parametrize("model,lat,n,q,nsector", [
    ("potts", "sc/26n", (24, 16, 12), 5, 8), ("potts/neighonly", "sc/26n", (16, 12, 10), 4, 2),
    ("ising", "sq/4n", (64, 40, 1), 2, 4), ("potts", "tri", (24, 18, 1), 4, 4), ("potts/pin", "sc/6n", (12, 10, 8), 3, 1),
], ids=lambda v: str(v).replace(" ", ""))
def test_masked_philox_sweeps_vs_oracle(model, lat, n, q, nsector):
    """Production (Philox) mode with masks at T = 0, sectors x colours and (nsector = 1) plain colour
    sets through spk_cuda_sweeps: spins, masks and counters against the oracle."""
    rng = np.random.default_rng(13)
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, 0.0)
    spins = np.repeat(rng.integers(1, (2 if model == "ising" else q) + 1, size=(o.nsites + 7) // 8, dtype=np.int32), 8)[:o.nsites].copy()
    if model == "potts/pin":
        spins[rng.random(o.nsites) < 0.05] = q + 1
    s = _sweep(model, lat, n, q, 0.0, seed=4711)
    s.upload(spins)
    s.mask_enable()
    want, mask = spins.copy(), np.zeros(o.nsites, dtype=np.uint8)
    nacc = 0
    colors = o.colors()
    for sw in range(4):
        if nsector == 1:
            for c in range(1, o.ncolors + 1):
                mask[colors == c] = 0  # a colour set as loop set: all of it is border (boundary_clear_mask)
                nacc += o.halfsweep_philox_mask(want, mask, 4711, sw, c)
            s.sweeps(sw, 1)
        else:
            for sec in range(nsector):
                box = o.sector_box(nsector, sec + 1)
                o.boundary_clear_mask(mask, box)
                s.mask_clear_border(box)
                for c in range(1, o.ncolors + 1):
                    nacc += o.halfsweep_philox_mask(want, mask, 4711, sw, c, box=box)
                    s.sweep_color(sw, c, box=box)
        assert np.array_equal(s.download(), want), sw
        assert np.array_equal(s.download_mask(), mask), sw
    assert s.counters() == (nacc, 4 * o.nsites)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "strict_*.npz"))), ids=os.path.basename)
def test_strict_runs_bit_exact_vs_reference(path):
    """`sweep color/strict`: whole multi-sweep runs of the reference, spin for spin."""
    g = np.load(path)
    model, lat = split_name(path, "strict")
    s = _sweep(model, lat, g["n"], g["q"], g["T"])
    s.upload(g["spins0"])
    s.set_siteseeds(g["seeds0"])
    for _ in range(int(g["nsweeps"])):
        for c in range(s.ncolors):
            s.sweep_color_strict(c + 1)
    assert np.array_equal(s.download(), g["after"])
    assert np.array_equal(s.get_siteseeds(), g["seeds1"])
    assert s.counters() == (int(g["naccept"]), int(g["nattempt"]))
    assert s.energy() == float(g["energy"])


CASES = [
    ("ising", "sq/4n", (64, 48, 1), 2, 1.1346),
    ("ising", "sq/4n", (50, 50, 1), 2, 0.0),
    ("ising", "sq/4n", (128, 66, 1), 2, 0.0),       # T = 0 on the bit shadow (thresholds unused)
    ("ising", "sq/4n", (192, 40, 1), 2, 7.0),       # hot: most uphill moves accepted
    ("ising/single", "sq/4n", (256, 36, 1), 2, 0.8),
    ("ising", "sc/6n", (16, 12, 10), 2, 2.25),
    ("potts", "sq/8n", (34, 22, 1), 7, 0.6),
    ("potts", "tri", (30, 17, 1), 9, 0.5),
    ("potts", "sc/6n", (12, 10, 14), 4, 0.7),
    ("potts", "sc/26n", (20, 14, 12), 100, 0.0),
    ("potts", "sc/26n", (18, 12, 16), 6, 1.0),
    ("potts/neighonly", "sc/26n", (20, 12, 14), 10000, 0.5),
    ("potts/neighonly", "sc/26n", (16, 12, 10), 3, 0.0),
    ("potts/neighonly", "tri", (20, 16, 1), 5, 0.4),
    ("potts/neighonly", "sq/4n", (24, 24, 1), 3, 0.9),
    ("potts/pin", "sc/26n", (16, 16, 12), 50, 0.5),
    ("potts/pin", "sq/8n", (20, 20, 1), 8, 0.8),
    # row lengths that are multiples of 8 elements: the tuned strip kernel takes these
    ("ising", "sq/4n", (256, 34, 1), 2, 1.1346),
    ("ising", "sq/8n", (40, 22, 1), 2, 2.0),
    ("potts", "sq/8n", (72, 18, 1), 7, 0.6),
    ("potts", "tri", (36, 11, 1), 9, 0.5),
    ("potts", "tri", (128, 6, 1), 3, 0.0),
    ("potts", "sc/6n", (24, 10, 14), 4, 0.7),
    ("potts", "sc/26n", (24, 14, 12), 100, 0.0),
    ("potts", "sc/26n", (264, 6, 4), 6, 1.0),
    ("potts", "sc/26n", (8, 8, 8), 3, 0.4),
    ("potts/neighonly", "sc/26n", (32, 12, 14), 10000, 0.5),
    ("potts/neighonly", "sc/26n", (16, 12, 10), 3, 0.0),
    ("potts/neighonly", "tri", (20, 16, 1), 5, 0.4),
    ("potts/neighonly", "sq/4n", (24, 24, 1), 3, 0.9),
    ("potts/pin", "sc/26n", (48, 10, 6), 50, 0.5),
    # several x/y tiles and plane chunks of the streaming kernel, ragged tile edges
    ("potts", "sc/26n", (272, 70, 40), 30, 0.7),
    ("potts", "sc/6n", (260, 40, 18), 5, 0.9),
    ("potts/neighonly", "sc/26n", (132, 34, 12), 40, 0.5),
    # several tiles of the colour-fused 2-D kernel (interior 240 x 64), ragged edges, short rows
    ("ising", "sq/4n", (496, 130, 1), 2, 1.1346),
    ("ising", "sq/4n", (8, 6, 1), 2, 1.5),
    ("potts", "sq/8n", (264, 70, 1), 6, 0.7),
    ("potts", "tri", (252, 66, 1), 9, 0.5),
    ("potts/neighonly", "tri", (124, 10, 1), 4, 0.6),
    ("potts/pin", "sq/8n", (128, 64, 1), 8, 0.8),
    # the variants on the same template, through every kernel family
    ("ising/single", "sq/4n", (64, 48, 1), 2, 1.2),
    ("ising/single", "sq/4n", (496, 130, 1), 2, 1.1346),
    ("ising/single", "sc/26n", (16, 12, 10), 2, 6.0),
    ("ising/single", "sc/6n", (24, 10, 14), 2, 2.2),
    ("potts/neigh", "sc/26n", (20, 12, 14), 50, 0.5),
    ("potts/neigh", "sc/26n", (132, 34, 12), 5, 0.5),
    ("potts/neigh", "sc/26n", (16, 12, 10), 70000, 0.0),
    ("potts/neigh", "tri", (252, 66, 1), 4, 0.6),
    ("potts/neigh", "sq/8n", (40, 22, 1), 3, 0.7),
    # largest Q of the packed 16-bit compare path (spins <= 65534) and the first Q beyond it
    ("potts/neighonly", "sc/26n", (16, 12, 10), 65534, 0.5),
    ("potts/pin", "sc/26n", (16, 12, 10), 65533, 0.5),
    ("potts/neighonly", "sc/26n", (16, 12, 10), 65535, 0.5),
    ("potts/pin", "sc/26n", (16, 12, 10), 65534, 0.5),
]


def _kernel_cases():
    out = []
    for c in CASES:
        for kernel in ["tuned", "tuned32", "pair", "strip", "generic"]:
            out.append(c + (kernel,))
        if c[0] in ("potts/neighonly", "potts/neigh") and c[1] == "sc/26n":
            out.append(c + ("adaptive",))  # candidate elimination instead of the all-pairs chain
        if c[0] in ("ising", "ising/single") and c[1] == "sq/4n" and c[2][0] % 64 == 0:
            out.append(c + ("nobits",))  # "tuned" runs on the bit shadow there: keep the colour-fused int32 kernel covered
        if c[0] in ("potts", "potts/pin") and c[1] == "sc/26n":
            out.append(c + ("int32stream",))  # "tuned" sweeps on the 16-bit shadow where it applies: keep the int32 kernel covered
    return out


@pytest.mark.parametrize("model,lat,n,q,T,kernel", _kernel_cases(), ids=lambda v: str(v).replace(" ", ""))
def test_philox_sweeps_bit_exact_vs_oracle(model, lat, n, q, T, kernel, monkeypatch):
    """Production (Philox) mode, three full sweeps, through every kernel family (TMA streaming
    kernel where it applies — colour-pair launches, per-colour launches, packed 16-bit and 32-bit
    compares, candidate elimination —, strip kernel, generic kernel): bit-exact against the oracle
    evaluating the same counter-based stream."""
    if kernel == "adaptive":
        monkeypatch.setenv("SPK_CUDA_NEIGH_MODE", "adaptive")
    if kernel == "nobits":
        monkeypatch.setenv("SPK_CUDA_NO_BITS", "1")
    if kernel == "int32stream":
        monkeypatch.setenv("SPK_CUDA_NO_SHADOW16", "1")
    if kernel == "tuned":
        monkeypatch.setenv("SPK_CUDA_NEIGH_MODE", "chain")
    if kernel == "generic":
        monkeypatch.setenv("SPK_CUDA_FORCE_GENERIC", "1")
    if kernel == "tuned32":
        monkeypatch.setenv("SPK_CUDA_NO_PK16", "1")
    if kernel == "pair":  # sc/26n: one launch per colour PAIR (opt-in) instead of one per colour
        monkeypatch.setenv("SPK_CUDA_PAIR", "1")
    if kernel == "strip":
        monkeypatch.setenv("SPK_CUDA_NO_STREAM", "1")
        monkeypatch.setenv("SPK_CUDA_NO_FUSED", "1")
    rng = np.random.default_rng(7)
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
    spins = rng.integers(1, (2 if model.startswith("ising") else q) + 1, size=o.nsites, dtype=np.int32)
    if model == "potts/pin":
        spins[rng.random(o.nsites) < 0.05] = q + 1
    s = _sweep(model, lat, n, q, T, seed=0xC0FFEE12345)
    s.upload(spins)
    want = spins.copy()
    nacc = 0
    for sw in range(3):
        nacc += o.sweep_philox(want, 0xC0FFEE12345, 10 + sw)
    s.sweeps(10, 3)
    assert np.array_equal(s.download(), want)
    assert s.counters() == (nacc, 3 * o.nsites)
    assert s.energy() == o.total_energy(want)


@pytest.mark.parametrize("model,q", [("potts/neighonly", 65534), ("potts", 65534), ("potts/pin", 65533)])
def test_packed16_extreme_spin_values(model, q):
    """Packed 16-bit compares: few distinct spin values taken from both ends of the 16-bit range
    (many equal neighbours, halfword borrows and the +1 of the replicated operand all exercised)."""
    rng = np.random.default_rng(31)
    n = (24, 16, 12)
    o = O.Oracle(model, "sc/26n", n[0], n[1], n[2], q, 0.6)
    palette = np.array([1, 2, 255, 256, 32767, 32768, 65279, 65280, q - 1, q], dtype=np.int32)
    spins = palette[rng.integers(0, len(palette), size=o.nsites)]
    if model == "potts/pin":
        spins[rng.random(o.nsites) < 0.05] = q + 1
    s = _sweep(model, "sc/26n", n, q, 0.6, seed=2718)
    s.upload(spins)
    want = spins.copy()
    nacc = sum(o.sweep_philox(want, 2718, sw) for sw in range(3))
    s.sweeps(0, 3)
    assert np.array_equal(s.download(), want)
    assert s.counters() == (nacc, 3 * o.nsites)


@pytest.mark.parametrize("model", ["potts/neighonly", "potts/neigh"])
@pytest.mark.parametrize("mode", ["adaptive", "auto"])
@pytest.mark.parametrize("start", ["random", "grains", "two-values"])
def test_candidate_elimination_vs_oracle(model, mode, start, monkeypatch):
    """sc/26n candidate-list apps: the candidate-elimination kernel (work ~ distinct neighbour values)
    and the automatic switch between it and the all-pairs chain, over a run that coarsens from 26
    distinct neighbours to a few; also from lattices that start with few distinct values (grains,
    a two-valued lattice where most sites have no event).  Several x/y tiles and plane chunks."""
    monkeypatch.setenv("SPK_CUDA_NEIGH_MODE", mode)
    monkeypatch.setenv("SPK_CUDA_NEIGH_SWITCH", "9")  # auto: switch early enough to happen inside the run
    rng = np.random.default_rng(77)
    n, q, T, seed = (136, 38, 12), 5000, 0.4, 0xABCDEF
    o = O.Oracle(model, "sc/26n", n[0], n[1], n[2], q, T)
    if start == "random":
        spins = rng.integers(1, q + 1, size=o.nsites, dtype=np.int32)
    elif start == "grains":  # 4 x 4 x 3 blocks of one value
        g = rng.integers(1, q + 1, size=(n[2] // 3, (n[1] + 3) // 4, n[0] // 4), dtype=np.int32)
        spins = np.repeat(np.repeat(np.repeat(g, 3, 0), 4, 1), 4, 2)[:, :n[1], :].reshape(-1).copy()
    else:
        spins = np.where(rng.random(o.nsites) < 0.03, 7, 4999).astype(np.int32)
    s = _sweep(model, "sc/26n", n, q, T, seed=seed)
    s.upload(spins)
    want = spins.copy()
    nsweeps = 10
    nacc = sum(o.sweep_philox(want, seed, sw) for sw in range(nsweeps))
    for sw in range(nsweeps):  # one call per sweep: the automatic mode re-decides at each first colour
        s.sweeps(sw, 1)
    assert np.array_equal(s.download(), want)
    assert s.counters() == (nacc, nsweeps * o.nsites)
    if mode == "auto":
        assert s.neigh_variant() in ("chain", "adaptive")


@pytest.mark.parametrize("model", ["ising", "ising/single"])
def test_bit_shadow_interleaved_with_int32_calls(model):
    """ising / ising/single on sq/4n: spk_cuda_sweeps runs multi-spin coded on a bit copy of the
    lattice; interleaved with calls that work on the int32 lattice (per-colour sweeps, energy, download,
    upload, randomize) nothing may be lost in the lazy conversions."""
    rng = np.random.default_rng(8)
    n, T, seed = (320, 46, 1), 1.2, 97531
    o = O.Oracle(model, "sq/4n", n[0], n[1], 1, 2, T)
    spins = rng.integers(1, 3, size=o.nsites, dtype=np.int32)
    s = _sweep(model, "sq/4n", n, 2, T, seed=seed)
    s.upload(spins)
    want = spins.copy()
    sw = 0
    s.sweeps(sw, 3)
    nacc = 0
    for _ in range(3):
        nacc += o.sweep_philox(want, seed, sw); sw += 1
    assert s.counters() == (nacc, 3 * o.nsites)
    assert s.energy() == o.total_energy(want)
    for c in (1, 2):
        s.sweep_color(sw, c)
        o.halfsweep_philox(want, seed, sw, c)
    sw += 1
    s.sweeps(sw, 2)
    for _ in range(2):
        o.sweep_philox(want, seed, sw); sw += 1
    assert np.array_equal(s.download(), want)
    s.randomize(5, 1, 2)
    got0 = s.download()
    s.sweeps(sw, 1)
    o.sweep_philox(got0, seed, sw)
    assert np.array_equal(s.download(), got0)


@pytest.mark.parametrize("model", ["potts", "potts/pin"])
def test_shadow16_interleaved_with_int32_calls(model):
    """sc/26n potts / potts/pin: spk_cuda_sweeps runs on the 16-bit shadow of the lattice; every other
    entry point works on the int32 lattice.  Interleave them (sweeps, per-colour calls, energy, pin,
    download, upload, randomize) and compare with the oracle after every step: the lazy conversions
    must never lose an update.  Several x / y tiles, ragged edges, Q near the 16-bit limit."""
    rng = np.random.default_rng(3)
    n, q, T, seed = (520, 70, 12), 65000, 0.6, 2468
    o = O.Oracle(model, "sc/26n", n[0], n[1], n[2], q, T)
    spins = rng.integers(1, q + 1, size=o.nsites, dtype=np.int32)
    if model == "potts/pin":
        spins[rng.random(o.nsites) < 0.03] = q + 1
    s = _sweep(model, "sc/26n", n, q, T, seed=seed)
    s.upload(spins)
    want = spins.copy()
    sw = 0
    s.sweeps(sw, 2)                                   # shadow
    for _ in range(2):
        o.sweep_philox(want, seed, sw); sw += 1
    assert s.energy() == o.total_energy(want)        # int32 (converted back)
    s.sweeps(sw, 1)                                   # shadow again (still current: no re-pack needed)
    o.sweep_philox(want, seed, sw); sw += 1
    for c in range(1, 9):                             # per-colour calls on the int32 lattice
        s.sweep_color(sw, c)
        o.halfsweep_philox(want, seed, sw, c)
    sw += 1
    s.sweeps(sw, 1)                                   # shadow must be re-packed
    o.sweep_philox(want, seed, sw); sw += 1
    assert np.array_equal(s.download(), want)
    want2 = np.roll(want, 7)
    s.upload(want2)                                   # replaces both copies
    s.sweeps(sw, 2)
    for _ in range(2):
        o.sweep_philox(want2, seed, sw); sw += 1
    assert s.digest() is not None and np.array_equal(s.download(), want2)
    assert s.energy() == o.total_energy(want2)


def test_per_colour_calls_equal_fused_sweeps():
    """spk_cuda_sweeps == halo_exchange + sweep_color per colour (the contract of the fast path)."""
    rng = np.random.default_rng(11)
    n, q = (32, 24, 20), 12
    spins = rng.integers(1, q + 1, size=n[0] * n[1] * n[2], dtype=np.int32)
    a = _sweep("potts", "sc/26n", n, q, 0.7, seed=99)
    b = _sweep("potts", "sc/26n", n, q, 0.7, seed=99)
    a.upload(spins); b.upload(spins)
    a.sweeps(0, 4)
    for sw in range(4):
        for c in range(1, 9):
            b.halo_exchange()
            b.sweep_color(sw, c)
    assert np.array_equal(a.download(), b.download())
    assert a.counters() == b.counters()


@pytest.mark.parametrize("n", [(12, 10, 8), (32, 10, 8)], ids=["generic", "strip"])
def test_nloop_and_box_vs_oracle(n):
    """nloop > 1 passes over one sector's colour set (src/app_lattice.cpp:674-676)."""
    rng = np.random.default_rng(5)
    q = 5
    o = O.Oracle("potts", "sc/26n", n[0], n[1], n[2], q, 0.9)
    spins = rng.integers(1, q + 1, size=o.nsites, dtype=np.int32)
    s = _sweep("potts", "sc/26n", n, q, 0.9, seed=4242)
    s.upload(spins)
    want = spins.copy()
    box = o.sector_box(8, 6)
    for c in range(1, 9):
        for l in range(3):
            o.halfsweep_philox(want, 4242, 7, c, iloop=l, box=box)
        s.sweep_color(7, c, box=box, nloop=3)
    assert np.array_equal(s.download(), want)


def test_edge_sizes_and_errors():
    from spparks_b200 import LatticeSweep, SpkCudaError
    # smallest legal lattices: every neighbour is a periodic image
    for model, lat, n, q in [("potts", "sc/26n", (2, 2, 2), 3), ("ising", "sq/4n", (2, 2, 1), 2),
                             ("potts", "tri", (2, 2, 1), 4), ("potts/neighonly", "sq/8n", (2, 2, 1), 3)]:
        o = O.Oracle(model, lat, n[0], n[1], n[2], q, 0.8)
        rng = np.random.default_rng(3)
        spins = rng.integers(1, (2 if model.startswith("ising") else q) + 1, size=o.nsites, dtype=np.int32)
        s = _sweep(model, lat, n, q, 0.8, seed=5)
        s.upload(spins)
        want = spins.copy()
        for sw in range(4):
            o.sweep_philox(want, 5, sw)
        s.sweeps(0, 4)
        assert np.array_equal(s.download(), want), (model, lat)
    with pytest.raises(SpkCudaError, match="incommensurate"):
        LatticeSweep("potts", "sc/26n", 7, 8, 8, nspins=5)
    with pytest.raises(SpkCudaError, match="nspins"):
        LatticeSweep("potts", "sc/26n", 8, 8, 8, nspins=0)
    s = LatticeSweep("potts", "sc/26n", 8, 8, 8, nspins=5)
    with pytest.raises(SpkCudaError, match="colour"):
        s.sweep_color(0, 9)
    with pytest.raises(SpkCudaError, match="slab owns"):
        s.upload(np.zeros(10, dtype=np.int32))
    # init_app's spin check (src/app_potts.cpp:83-90): values outside 1..Q are refused on upload
    bad = np.ones(512, dtype=np.int32)
    bad[100] = 6
    with pytest.raises(SpkCudaError, match="invalid values"):
        s.upload(bad)
    bad[100] = 0
    with pytest.raises(SpkCudaError, match="invalid values"):
        s.upload(bad)
    p = LatticeSweep("potts/pin", "sc/26n", 8, 8, 8, nspins=5)
    bad[100] = 6  # Q+1 marks a pinned site (src/app_potts_pin.cpp:81)
    p.upload(bad)


def test_device_randomize_distribution():
    s = _sweep("potts", "sc/26n", (64, 64, 64), 100, 0.0, seed=1)
    s.randomize(12345, 1, 100)
    a = s.download()
    assert a.min() == 1 and a.max() == 100
    counts = np.bincount(a, minlength=101)[1:]
    assert abs(counts - a.size / 100).max() < 6 * np.sqrt(a.size / 100)


@pytest.mark.parametrize("model,lat,n,q,T", [
    ("ising", "sq/4n", (4096, 2048, 1), 2, 1.1346),     # 576 tiles of the fused 2-D kernel: more than are resident
    ("potts", "tri", (2048, 1280, 1), 50, 0.5),
    ("potts", "sc/26n", (256, 160, 96), 200, 0.6),      # 40 plane chunks x 10 tiles of the streaming kernel
], ids=["ising-sq4n", "potts-tri", "potts-sc26n"])
def test_large_lattices_bit_exact_vs_oracle(model, lat, n, q, T):
    """Lattices with more tiles than resident CTAs, so that tiles are loaded after other tiles have
    already been stored (catches in-place halo hazards that small lattices cannot show)."""
    rng = np.random.default_rng(21)
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
    spins = rng.integers(1, (2 if model.startswith("ising") else q) + 1, size=o.nsites, dtype=np.int32)
    s = _sweep(model, lat, n, q, T, seed=777)
    s.upload(spins)
    want = spins.copy()
    nacc = 0
    for sw in range(2):
        nacc += o.sweep_philox(want, 777, sw)
    s.sweeps(0, 2)
    assert np.array_equal(s.download(), want)
    assert s.counters() == (nacc, 2 * o.nsites)


@pytest.mark.parametrize("model,lat,n,q", [
    ("ising", "sq/4n", (50, 50, 1), 2),          # examples/ising/in.ising: `set site range 1 2`
    ("potts", "tri", (30, 22, 1), 1000),
    ("potts/neighonly", "sc/26n", (20, 14, 12), 10000),
    ("potts/pin", "sc/6n", (12, 10, 8), 50),
], ids=lambda v: str(v).replace(" ", ""))
def test_device_set_site_range_is_the_reference_stream(model, lat, n, q):
    """spk_cuda_set_site_range reproduces `set site range 1 Q` value for value: the one Park-Miller
    stream of src/set.cpp:617-627 (the oracle's restatement of it replays the shipped logs'
    time-0 energies, tests/test_oracle_golden.py) evaluated per site with a modular power."""
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, 0.5)
    for master_seed in (56789, 4711):
        m = O.mars(master_seed)
        want = o.set_site_range(1, q, m)
        m = O.mars(master_seed)
        park = int(O.lib().orc_mars_uniform(O.C.byref(m)) * O.IM) or 1   # new RandomPark(ranmaster->uniform())
        s = _sweep(model, lat, n, q, 0.5)
        s.set_site_range(park, 1, q)
        assert np.array_equal(s.download(), want)
        assert s.energy() == o.total_energy(want)
        # per-site energies (dump ... energy)
        e = s.site_energy()
        assert e.sum() == o.total_energy(want)
        for site in (0, 1, o.nsites // 2, o.nsites - 1):
            assert e[site] == O.lib().orc_site_energy(o.ctx, want.ctypes.data_as(O.C.POINTER(O.C.c_int32)), site)
        s.close()


def test_device_set_site_range_large_offsets():
    """draw numbers far along the stream (modular power with a 31-bit exponent and beyond): the last
    sites of a 2^27-site lattice against the sequential stream advanced on the host."""
    n = (512, 512, 512)
    s = _sweep("potts", "sc/26n", n, 50000, 0.0)
    s.set_site_range(123456789, 1, 50000)
    got = s.download()
    state, IM = 123456789, O.IM
    for i in range(1000):
        state = state * 16807 % IM
        assert got[i] == min(int(state * (1.0 / IM) * 50000), 49999) + 1, i
    N = got.size
    state = 123456789 * pow(16807, N - 1000, IM) % IM
    for i in range(N - 1000, N):
        state = state * 16807 % IM
        assert got[i] == min(int(state * (1.0 / IM) * 50000), 49999) + 1, i
    s.close()
