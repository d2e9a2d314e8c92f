"""`cuda_create_sites`: the lattice lives on the device only (no per-site host arrays, so the
BASELINE lattices can be driven from an input script).  The same script with `create_sites box` (host copy,
uploaded before the run) and with `cuda_create_sites` must print the same table: same `set site range` stream,
same pins, same Philox seed, same sweeps, same device diagnostics."""
import os
import re

import pytest

import test_plugin as P

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(P.SPK_CUDA), reason="plugin/_build/spk_cuda not built")]

SCRIPT = """seed 56789
app_style {app}
dimension {dim}
lattice {lat} 1.0
region box block 0 {nx} 0 {ny} {zr}
create_box box
{create}
cuda_set site range 1 {hi}
{extra}
sweep color {sweepopts}
sector {sector}
temperature {T}
diag_style energy/cuda
{diag}
stats {stats}
run {run}
run {run}
"""

CASES = [
    dict(app="potts/cuda 20", dim=3, lat="sc/26n", nx=16, ny=12, zr="0 10", hi=20, T=0.8, stats=0.1, run=0.5, extra="",
         sector="no", sweepopts="", diag=""),
    dict(app="potts/neighonly/cuda 30", dim=3, lat="sc/26n", nx=16, ny=8, zr="0 12", hi=30, T=0.5, stats=0.5, run=1.0, extra="",
         sector="no", sweepopts="", diag="diag_style cluster/cuda"),
    dict(app="potts/pin/cuda 12", dim=3, lat="sc/26n", nx=8, ny=8, zr="0 8", hi=12, T=1.0, stats=0.25, run=0.5,
         extra="cuda_pin 0.05 0 0", sector="no", sweepopts="", diag=""),
    dict(app="ising/cuda", dim=2, lat="sq/4n", nx=64, ny=32, zr="-0.5 0.5", hi=2, T=1.1346, stats=2.0, run=10.0, extra="",
         sector="no", sweepopts="", diag=""),
    dict(app="potts/cuda 9", dim=2, lat="tri", nx=16, ny=10, zr="-0.5 0.5", hi=9, T=0.5, stats=0.5, run=1.0, extra="",
         sector="yes", sweepopts="", diag=""),
    dict(app="potts/cuda 6", dim=3, lat="sc/6n", nx=12, ny=8, zr="0 10", hi=6, T=0.7, stats=0.5, run=1.0, extra="",
         sector="yes", sweepopts="mask yes", diag=""),
    dict(app="potts/neigh/cuda 15", dim=3, lat="sc/26n", nx=12, ny=8, zr="0 10", hi=15, T=0.6, stats=0.5, run=1.0, extra="",
         sector="8 nstop 2.0", sweepopts="", diag=""),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: (c["app"] + "-" + c["lat"] + "-" + c["sector"]).replace(" ", "_"))
def test_device_lattice_script_equals_host_lattice_script(case):
    host = P.run(P.SPK_CUDA, SCRIPT.format(**dict(case, create="create_sites box")))
    dev = P.run(P.SPK_CUDA, SCRIPT.format(**dict(case, create="cuda_create_sites")))
    assert "sites on the device (no host copy)" in dev and "sites on the device" not in host
    rh, rd = P.stats_rows(host), P.stats_rows(dev)
    assert len(rh) >= 4 and rd == rh
    # nothing was uploaded or downloaded in the device-lattice run: the loop line is there, the rates differ
    assert re.search(r"CUDA sweep loop: .* Gsites/s", dev)


def test_device_lattice_refuses_what_needs_the_host_copy():
    base = SCRIPT.format(**dict(CASES[0], create="cuda_create_sites"))
    # the reference's own commands on per-site data find no sites (a dump writes empty snapshots)
    out = P.run(P.SPK_CUDA, base.replace("sweep color", "set site range 1 5\nsweep color", 1), expect_error=True)
    assert "ERROR: Set command before sites exist" in out, out[-500:]
    out = P.run(P.SPK_CUDA, base.replace("sweep color ", "sweep color/strict ", 1), expect_error=True)
    assert "needs sweep color" in out
    # a lattice that was never set holds zeros: init_app's check of the site values refuses it
    out = P.run(P.SPK_CUDA, base.replace("cuda_set site range 1 20\n", ""), expect_error=True)
    assert "invalid values" in out


MASKRUN = """seed 4711
app_style potts/cuda 2
dimension 3
lattice sc/6n 1.0
region box block 0 12 0 12 0 12
create_box box
{create}
cuda_set site range 1 2
sweep color mask {mask}
sector yes
temperature 1.0
diag_style energy/cuda
stats 1.0
{runs}
"""


@pytest.mark.parametrize("create", ["create_sites box", "cuda_create_sites"])
def test_two_runs_equal_one_run_with_masks(create):
    """The reference keeps its mask array from run to run (created and zeroed once, src/app_lattice.cpp:258-267) and
    the counter-based stream continues, so `run 3; run 3` ends where `run 6` ends - also with `mask yes` + sectors,
    where the masks decide which sites are attempted (the plugin carries the bytes over when it rebuilds the device
    lattice for a new run; the device-only lattice keeps them).  Q = 2 at T = 1: masked sites would flip readily,
    so masks that were reset between the runs show (they did, before the bytes were carried over)."""
    def last(script):
        rows = P.stats_rows(P.run(P.SPK_CUDA, script))
        return rows[-1][:4] + rows[-1][-1:]
    one = last(MASKRUN.format(create=create, mask="yes", runs="run 6.0"))
    two = last(MASKRUN.format(create=create, mask="yes", runs="run 3.0\nrun 3.0"))
    assert two == one
    # the masks do decide something here: without them the trajectory is another one
    assert last(MASKRUN.format(create=create, mask="no", runs="run 6.0")) != one
