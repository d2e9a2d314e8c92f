"""More GPU parity cases, found missing by the mutation check of the emulated tier (tools/emu_mutants.py):
small lattices never took the multi-step work items of the streaming kernels (several class planes through
the shared-memory ring), never let one CTA process several items, and never had a y extent that is a
whole number of tiles — the paths every BASELINE-size run lives on (there they are covered by
tests/test_full_size.py, tuned vs generic digests).  SPK_CUDA_MAX_PLANE_CHUNKS / SPK_CUDA_MAX_CTAS
(model_entry.cuh) force those paths at sizes the oracle finishes in seconds."""
import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu

TIERS = {
    "tuned": {},
    "chain": {"SPK_CUDA_NEIGH_MODE": "chain"},
    "adaptive": {"SPK_CUDA_NEIGH_MODE": "adaptive"},
    "tuned32": {"SPK_CUDA_NO_PK16": "1"},
    "pair": {"SPK_CUDA_PAIR": "1", "SPK_CUDA_NO_SHADOW16": "1"},
    "int32stream": {"SPK_CUDA_NO_SHADOW16": "1"},
    "nobits": {"SPK_CUDA_NO_BITS": "1"},
}
CASES = [
    # 3-D: x tiles of 128 (256 on the 16-bit shadow), y tiles of 32 rows (16: sc/6n, candidate elimination)
    ("potts", "sc/26n", (136, 64, 12), 30, 0.7, ["tuned", "int32stream", "pair", "tuned32"]),
    ("potts/pin", "sc/26n", (264, 32, 10), 9, 0.9, ["tuned", "int32stream"]),
    ("potts/pin", "sc/26n", (24, 70, 16), 50, 0.5, ["tuned", "int32stream"]),
    ("potts/neighonly", "sc/26n", (132, 32, 12), 40, 0.5, ["chain", "adaptive", "tuned32"]),
    ("potts/neighonly", "sc/26n", (16, 16, 20), 3, 0.0, ["chain", "adaptive"]),
    ("potts/neigh", "sc/26n", (40, 34, 14), 5, 0.6, ["chain", "adaptive"]),
    ("ising", "sc/6n", (260, 32, 10), 2, 2.2, ["tuned"]),
    ("potts", "sc/6n", (16, 16, 14), 4, 0.7, ["tuned"]),
    ("ising/single", "sc/26n", (16, 32, 10), 2, 6.0, ["tuned"]),
    # 2-D: several tiles of the colour-fused kernel (interior 240 x 64) for one CTA
    ("ising", "sq/4n", (496, 130, 1), 2, 1.1346, ["nobits"]),
    ("ising", "sq/4n", (512, 128, 1), 2, 1.0, ["tuned", "nobits"]),
    ("potts", "sq/8n", (264, 70, 1), 6, 0.7, ["tuned"]),
    ("potts", "tri", (252, 66, 1), 9, 0.5, ["tuned"]),
    ("potts/neighonly", "tri", (128, 64, 1), 4, 0.6, ["tuned"]),
]


def _cases():
    return [(m, lat, n, q, T, tier, grid) for m, lat, n, q, T, tiers in CASES for tier in tiers
            for grid in ("default", "one-chunk-two-ctas", "two-chunks-one-cta")]


@pytest.mark.parametrize("model,lat,n,q,T,tier,grid", _cases(), ids=lambda v: str(v).replace(" ", ""))
def test_multi_step_items_and_several_items_per_cta(model, lat, n, q, T, tier, grid, monkeypatch):
    for k, v in TIERS[tier].items():
        monkeypatch.setenv(k, v)
    if grid == "one-chunk-two-ctas":
        monkeypatch.setenv("SPK_CUDA_MAX_PLANE_CHUNKS", "1")
        monkeypatch.setenv("SPK_CUDA_MAX_CTAS", "2")
    elif grid == "two-chunks-one-cta":
        monkeypatch.setenv("SPK_CUDA_MAX_PLANE_CHUNKS", "2")
        monkeypatch.setenv("SPK_CUDA_MAX_CTAS", "1")
    from spparks_b200 import LatticeSweep
    rng = np.random.default_rng(11)
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
    spins = rng.integers(1, (2 if model.startswith("ising") else q) + 1, size=o.nsites, dtype=np.int32)
    if model == "potts/pin":
        spins[rng.random(o.nsites) < 0.05] = q + 1
    s = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=0xBEEF1234)
    s.upload(spins)
    want = spins.copy()
    nacc = sum(o.sweep_philox(want, 0xBEEF1234, 40 + sw) for sw in range(2))
    s.sweeps(40, 2)
    assert np.array_equal(s.download(), want)
    assert s.counters() == (nacc, 2 * o.nsites)
    assert s.energy() == o.total_energy(want)
    s.close()


@pytest.mark.parametrize("model,lat,n,nsector", [
    ("potts", "sc/26n", (8, 8, 8), 8), ("potts", "sc/6n", (8, 6, 8), 2), ("potts", "sq/8n", (12, 12, 1), 4),
    ("potts", "tri", (8, 8, 1), 2), ("ising", "sq/4n", (16, 12, 1), 4),
], ids=lambda v: str(v).replace(" ", ""))
def test_mask_border_of_every_sector_face(model, lat, n, nsector):
    """`sweep color mask yes` on a uniform lattice at T = 0: nothing changes and after one pass every site is
    masked (einitial = 0), so the second pass shows exactly which sites boundary_clear_mask unmasks for each
    sector — every face of every sector box (src/app_lattice.cpp:1092-1164)."""
    from spparks_b200 import LatticeSweep
    o = O.Oracle(model, lat, n[0], n[1], n[2], 3, 0.0)
    spins = np.ones(o.nsites, dtype=np.int32)
    s = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=3, temperature=0.0, seed=5)
    s.upload(spins)
    s.mask_enable(True)
    mask = np.zeros(o.nsites, dtype=np.uint8)
    want = spins.copy()
    for p in range(2):
        for isec in range(nsector):
            box = o.sector_box(nsector, isec)
            o.boundary_clear_mask(mask, box)
            s.mask_clear_border(box, 0)
            if p == 1:  # what the clearing left of an all-masked lattice
                assert np.array_equal(s.download_mask(), mask), (p, isec)
            for color in range(1, o.ncolors + 1):
                o.halfsweep_philox_mask(want, mask, 5, p, color, 0, box)
                s.sweep_color(p, color, box, 1)
        assert np.array_equal(s.download_mask(), mask)
        if p == 0:
            assert mask.all()
    assert np.array_equal(s.download(), want) and (want == 1).all()
    s.close()


@pytest.mark.parametrize("T", [0.0, 0.05, 0.3, 0.8, 1.1346, 2.5, 40.0])
def test_acceptance_thresholds_are_the_oracles(T):
    """the table the Philox-mode kernels compare the acceptance word with == floor(exp(-dE/T) 2^32) as the oracle
    computes it (whose integer reduction tests/test_oracle_rng.py pins to the reference's double formulas); a
    table that is off by one would bias one draw in 2^32 and no trajectory test could see it"""
    from spparks_b200 import LatticeSweep
    s = LatticeSweep("potts", "sc/26n", 8, 8, 8, nspins=5, temperature=0.7)
    s.set_temperature(T)
    thr = s.accept_thresholds()
    if T == 0.0:
        assert all(t == 0 for t in thr[1:])
    else:
        assert thr[0] == 2**32 - 1
        for dE in range(1, 27):
            assert thr[dE] == O.accept_threshold(T, dE), dE
    s.close()


@pytest.mark.parametrize("model,lat,n,q", [("potts", "sc/26n", (16, 8, 8), 20), ("potts/pin", "sc/26n", (8, 8, 8), 9),
                                            ("ising", "sq/4n", (64, 16, 1), 2)], ids=lambda v: str(v).replace(" ", ""))
def test_masks_switched_off_again(model, lat, n, q):
    """spk_cuda_mask_enable(h, 0) after masked sweeps: the following sweeps run on the 16-bit / 1-bit shadow again
    (switching the masks off once freed the shadows: found by the device-lattice mode of the plugin, which
    switches them off at every init)"""
    from spparks_b200 import LatticeSweep
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, 0.6)
    rng = np.random.default_rng(3)
    spins = np.repeat(rng.integers(1, (2 if model == "ising" else q) + 1, size=(o.nsites + 3) // 4, dtype=np.int32), 4)[:o.nsites].copy()
    s = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=0.6, seed=99)
    s.upload(spins)
    want = spins.copy()
    s.sweeps(0, 1)
    o.sweep_philox(want, 99, 0)
    s.mask_enable(True)
    mask = np.zeros(o.nsites, dtype=np.uint8)
    ns = 8 if lat.startswith("sc") else 4
    for isec in range(ns):
        box = o.sector_box(ns, isec)
        o.boundary_clear_mask(mask, box)
        s.mask_clear_border(box, 0)
        for color in range(1, o.ncolors + 1):
            o.halfsweep_philox_mask(want, mask, 99, 1, color, 0, box)
            s.sweep_color(1, color, box, 1)
    assert np.array_equal(s.download(), want)
    s.mask_enable(False)
    for call in range(2):
        s.sweeps(2 + 2 * call, 2)
        for sw in range(2):
            o.sweep_philox(want, 99, 2 + 2 * call + sw)
        assert np.array_equal(s.download(), want), call
    assert s.energy() == o.total_energy(want)
    s.close()
