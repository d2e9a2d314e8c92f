"""The drop-in: unchanged SPPARKS input scripts with `app_style <name>/cuda`, run by the reference
host (unmodified objects) + plugin/ + libspparks_cuda  (plugin/_build/spk_cuda), next to the
unmodified reference binary (oracle/_ref/spk_ref).  Both binaries are prebuilt in the build
container and travel to the GPU box; the tests skip if they are absent."""
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPK_CUDA = os.path.join(ROOT, "plugin", "_build", "spk_cuda")
SPK_REF = os.path.join(ROOT, "oracle", "_ref", "spk_ref")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(SPK_CUDA) and os.path.exists(SPK_REF)),
                                 reason="plugin/_build/spk_cuda or oracle/_ref/spk_ref not built")]


def run(binary, script, expect_error=False):
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "in.script")
        open(path, "w").write(script)
        out = subprocess.run([binary, "-in", path, "-log", "none"], cwd=tmp, capture_output=True, text=True,
                             timeout=600)
    if expect_error:
        return out.stdout + out.stderr
    assert "ERROR" not in out.stdout, out.stdout[-2000:]
    return out.stdout


def stats_rows(text):
    """rows of every stats table: (time, naccept, nreject, nsweeps, extra columns without CPU)"""
    rows, on = [], False
    for line in text.splitlines():
        f = line.split()
        if f[:2] == ["Time", "Naccept"]:
            on = True
            continue
        if on:
            if f and f[0] == "WARNING:":
                continue
            if len(f) >= 5 and re.match(r"^[0-9.eE+-]+$", f[0]):
                rows.append(tuple(f[:4] + f[5:]))
            else:
                on = False
    return rows


STRICT = """seed 4711
app_style {app}
dimension {dim}
lattice {lat} 1.0
region box block 0 {nx} 0 {ny} {zr}
create_box box
create_sites box
set site range 1 {hi}
{extra}
sweep color/strict
sector no
temperature {T}
diag_style energy
stats {stats}
run {run}
run {run}
"""

STRICT_CASES = [
    dict(app="potts", q=20, dim=3, lat="sc/26n", nx=16, ny=12, zr="0 10", hi=20, T=0.8, stats=0.1, run=0.5, extra=""),
    dict(app="potts/neighonly", q=30, dim=3, lat="sc/26n", nx=16, ny=8, zr="0 12", hi=30, T=0.5, stats=0.5, run=1.0,
         extra=""),
    dict(app="potts/pin", q=12, dim=3, lat="sc/26n", nx=8, ny=8, zr="0 8", hi=12, T=1.0, stats=0.25, run=0.5,
         extra="pin 0.05 0 0"),
    dict(app="ising", q=0, dim=2, lat="sq/4n", nx=64, ny=32, zr="-0.5 0.5", hi=2, T=1.1346, stats=2.0, run=10.0,
         extra=""),
    dict(app="potts", q=9, dim=2, lat="tri", nx=16, ny=10, zr="-0.5 0.5", hi=9, T=0.5, stats=0.5, run=1.0, extra=""),
    dict(app="ising/single", q=0, dim=2, lat="sq/4n", nx=48, ny=32, zr="-0.5 0.5", hi=2, T=1.3, stats=4.0, run=20.0,
         extra=""),
    dict(app="potts/neigh", q=15, dim=3, lat="sc/26n", nx=12, ny=8, zr="0 10", hi=15, T=0.6, stats=0.5, run=1.0,
         extra=""),
]


@pytest.mark.parametrize("case", STRICT_CASES, ids=lambda c: c["app"] + "-" + c["lat"])
def test_color_strict_script_identical_to_reference(case):
    """`sweep color/strict` makes the reference independent of traversal order, so the GPU run of
    the same script must print the same Naccept / Nreject / Nsweeps / Energy table, over two
    consecutive `run` commands."""
    def script(style):
        app = style if case["app"].startswith("ising") else "%s %d" % (style, case["q"])
        return STRICT.format(**dict(case, app=app))
    ref = stats_rows(run(SPK_REF, script(case["app"])))
    gpu = stats_rows(run(SPK_CUDA, script(case["app"] + "/cuda")))
    assert len(ref) >= 4 and gpu == ref


def test_in_ising_with_cuda_style():
    """examples/ising/in.ising with only the style name changed (sweep random is mapped to colour
    sweeps): same attempts, sweeps and time axis; energy relaxes as in the shipped log
    (examples/ising/log.ising.11Nov09.linux.1: 5112 -> 320)."""
    script = """seed 56789
app_style ising/cuda
dimension 2
lattice sq/4n 1.0
region box block 0 50 0 50 -0.5 0.5
create_box box
create_sites box
set site range 1 2
sweep random
sector yes
diag_style energy
temperature 0.0
stats 10.0
run 100.0
"""
    rows = stats_rows(run(SPK_CUDA, script))
    assert len(rows) == 11
    assert rows[0][4] == "5112"  # same `set site range` stream as the reference
    assert [r[3] for r in rows] == [str(20 * i) for i in range(11)]
    assert int(rows[-1][1]) + int(rows[-1][2]) == 500000
    assert 150 <= float(rows[-1][4]) <= 700


@pytest.mark.parametrize("app,q,dim,lat,n,T", [
    ("ising", 2, 2, "sq/4n", (64, 32, 1), 1.2),
    ("potts", 30, 2, "tri", (32, 20, 1), 0.5),
    ("potts/neighonly", 40, 3, "sc/26n", (16, 16, 12), 0.5),
], ids=["ising-sq4n", "potts-tri", "neighonly-sc26n"])
def test_color_sweep_script_matches_library_and_oracle(app, q, dim, lat, n, T):
    """`sweep color` through the plugin (Philox stream, `cuda_seed` fixed) prints the Naccept and
    Energy that the C ABI driven from Python and the CPU oracle produce for the same seed and the same
    `set site range` lattice (whose reference stream the oracle restates)."""
    import numpy as np
    import oracle_py as O
    from spparks_b200 import LatticeSweep
    style = app + "/cuda" if app == "ising" else "%s/cuda %d" % (app, q)
    dt = 0.5 if app == "ising" else (1.0 / 26 if app == "potts/neighonly" else 1.0 / q)
    zr = "0 %d" % n[2] if dim == 3 else "-0.5 0.5"
    script = """seed 4242
app_style %s
dimension %d
lattice %s 1.0
region box block 0 %d 0 %d %s
create_box box
create_sites box
set site range 1 %d
cuda_seed 987654321
sweep color
sector no
temperature %r
diag_style energy
stats %.17g
run %.17g
""" % (style, dim, lat, n[0], n[1], zr, q, T, dt * 0.999, dt * 5.5)
    rows = stats_rows(run(SPK_CUDA, script))
    assert [r[3] for r in rows] == [str(i) for i in range(7)]
    o = O.Oracle(app, lat, n[0], n[1], n[2], q, T)
    m = O.mars(4242)
    spins = o.set_site_range(1, q, m)
    s = LatticeSweep(app, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=987654321)
    s.upload(spins)
    want = spins.copy()
    nacc = 0
    assert float(rows[0][4]) == o.total_energy(want)
    for sw in range(6):
        s.sweeps(sw, 1)
        nacc += o.sweep_philox(want, 987654321, sw)
        assert int(rows[sw + 1][1]) == s.counters()[0] == nacc
        assert float(rows[sw + 1][4]) == s.energy() == o.total_energy(want)


@pytest.mark.parametrize("app,q,dim,lat,n,T", [
    ("potts/neighonly", 60, 3, "sc/26n", (24, 16, 20), 0.5),
    ("potts", 12, 2, "tri", (48, 30, 1), 0.4),
    ("ising", 2, 2, "sq/4n", (64, 48, 1), 0.9),
], ids=["neighonly-sc26n", "potts-tri", "ising-sq4n"])
def test_device_diagnostics_equal_host_diagnostics(app, q, dim, lat, n, T):
    """`diag_style energy/cuda` and `diag_style cluster/cuda` (plugin/diag_*_cuda.h: the sums and
    the connected-component analysis on the device) next to the reference's own host-side
    `diag_style energy` and `diag_style cluster` in the same run: identical stats columns at every
    output, including the one at time 0 that precedes the first sweep."""
    style = app + "/cuda" if app == "ising" else "%s/cuda %d" % (app, q)
    dt = 0.5 if app == "ising" else (1.0 / 26 if app == "potts/neighonly" else 1.0 / q)
    zr = "0 %d" % n[2] if dim == 3 else "-0.5 0.5"
    script = """seed 777
app_style %s
dimension %d
lattice %s 1.0
region box block 0 %d 0 %d %s
create_box box
create_sites box
set site range 1 %d
sweep color
sector no
temperature %r
diag_style energy
diag_style cluster
diag_style energy/cuda
diag_style cluster/cuda
stats %.17g
run %.17g
""" % (style, dim, lat, n[0], n[1], zr, q, T, dt * 3.999, dt * 20.5)
    rows = stats_rows(run(SPK_CUDA, script))
    assert len(rows) >= 6
    for r in rows:
        host, dev = r[4:8], r[8:12]   # Energy Nclust <N> <R> | the same four from the device
        assert host == dev, r
    assert int(rows[-1][5]) < int(rows[0][5])  # the lattice coarsens


@pytest.mark.parametrize("case", [STRICT_CASES[0], STRICT_CASES[1], STRICT_CASES[3]], ids=lambda c: c["app"] + "-" + c["lat"])
def test_mask_yes_script_identical_to_reference(case):
    """`sweep color/strict mask yes`: the plugin keeps the mask bytes on the device
    (spk_cuda_mask_enable / _clear_border); the reference forbids sectors with color/strict, so the
    loop sets are the colour sets, whose masks boundary_clear_mask clears before every pass — the
    table must equal the reference's own `mask yes` run (and its `mask no` run)."""
    def script(style, mask):
        app = style if case["app"].startswith("ising") else "%s %d" % (style, case["q"])
        return STRICT.format(**dict(case, app=app)).replace("sweep color/strict", "sweep color/strict mask " + mask)
    ref = stats_rows(run(SPK_REF, script(case["app"], "yes")))
    assert ref == stats_rows(run(SPK_REF, script(case["app"], "no")))
    gpu = stats_rows(run(SPK_CUDA, script(case["app"] + "/cuda", "yes")))
    assert len(ref) >= 4 and gpu == ref


def test_mask_yes_with_sectors_relaxes_like_mask_no():
    """`sweep color mask yes` + `sector yes` at T = 0 through the plugin (the masked path that the
    reference fixtures pin at the ABI level): same attempts and time axis as `mask no`, energies
    relax to the same level (masking must not change the answer in a statistical sense, doc/sweep.txt)."""
    def script(mask):
        return """seed 99
app_style potts/cuda 20
dimension 3
lattice sc/26n 1.0
region box block 0 24 0 24 0 24
create_box box
create_sites box
set site range 1 20
cuda_seed 5
sweep color mask %s
sector yes
temperature 0.0
diag_style energy
stats 0.25
run 1.0
""" % mask
    yes, no = stats_rows(run(SPK_CUDA, script("yes"))), stats_rows(run(SPK_CUDA, script("no")))
    assert [r[3] for r in yes] == [r[3] for r in no] and len(yes) == 5
    assert int(yes[-1][1]) + int(yes[-1][2]) == int(no[-1][1]) + int(no[-1][2]) == 20 * 24 ** 3
    e_yes, e_no = float(yes[-1][4]), float(no[-1][4])
    assert e_yes < 0.8 * float(yes[0][4]) and abs(e_yes - e_no) < 0.03 * e_no, (e_yes, e_no)


def test_finish_line_and_host_sync_option():
    """The plugin prints its sweep-loop throughput next to the reference's timing breakdown, and
    `cuda_host_sync no` (only /cuda diagnostics read the lattice during the run) leaves the table
    unchanged while the host copy is refreshed only when the run ends."""
    base = """seed 3
app_style ising/cuda
dimension 2
lattice sq/4n 1.0
region box block 0 128 0 96 -0.5 0.5
create_box box
create_sites box
set site range 1 2
cuda_seed 17
%s
sweep color
sector no
temperature 1.0
diag_style energy/cuda
stats 2.0
run 10.0
diag_style energy
run 0.0
"""
    a = run(SPK_CUDA, base % "cuda_host_sync yes")
    b = run(SPK_CUDA, base % "cuda_host_sync no")
    ra, rb = stats_rows(a), stats_rows(b)
    assert ra == rb and len(ra) >= 7
    # the final `run 0.0` adds the host-side energy of the downloaded lattice: equal to the device's
    assert ra[-1][4] == ra[-1][5]
    assert re.search(r"CUDA sweep loop: .* Gsites/s", a)


def test_cuda_set_site_range_equals_set_site_range():
    """`cuda_set site range 1 Q` initialises the lattice on the device with the reference's own
    stream: the whole color/strict table (time-0 energy included) equals the reference's run of
    the script with `set site range 1 Q`, and a following `set`-dependent draw (the strict site
    seeds come from the same master generator) is unshifted."""
    case = dict(STRICT_CASES[1])
    ref_script = STRICT.format(**dict(case, app="%s %d" % (case["app"], case["q"])))
    gpu_script = STRICT.format(**dict(case, app="%s/cuda %d" % (case["app"], case["q"]))).replace(
        "set site range 1 %d" % case["hi"], "cuda_set site range 1 %d" % case["hi"])
    assert "cuda_set" in gpu_script
    ref = stats_rows(run(SPK_REF, ref_script))
    out = run(SPK_CUDA, gpu_script)
    assert "%d settings made for site" % (16 * 8 * 12) in out
    assert stats_rows(out) == ref and len(ref) >= 4
