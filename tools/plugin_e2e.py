#!/usr/bin/env python
"""e2e through the REAL plugin binary: plugin/_build/spk_cuda runs an in.* script (ising/cuda on
a 4096^2 sq/4n lattice, `sweep color`, one stats line per 20 sweeps with `diag_style energy/cuda`,
200 sweeps) and the site-updates/s of its sweep loop is read back from the reference's own
"Loop time" line: uploads, sweeps, counter reads, downloads and output all inside.
bench.py puts the result under "e2e_plugin"."""
import os
import re
import subprocess
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPK_CUDA = os.path.join(ROOT, "plugin", "_build", "spk_cuda")

SCRIPT = """seed 56789
app_style ising/cuda
dimension 2
lattice sq/4n 1.0
region box block 0 {n} 0 {n} -0.5 0.5
create_box box
create_sites box
cuda_set site range 1 2
cuda_seed 12345
cuda_host_sync {sync}
sweep color
sector no
temperature 1.1346
diag_style energy/cuda
stats 10.0
run {t}
"""


HEADLINE_SCRIPT = """seed 56789
app_style potts/neighonly/cuda 10000
dimension 3
lattice sc/26n 1.0
region box block 0 {n} 0 {n} 0 {n}
create_box box
cuda_create_sites
cuda_set site range 1 10000
cuda_seed 12345
sweep color
sector no
temperature 0.5
diag_style energy/cuda
stats {stats}
run {t}
"""


def plugin_e2e(n=4096, sweeps=200):
    """Default plugin behaviour (host copy refreshed at every stats line) and, next to it,
    `cuda_host_sync no` (only /cuda diagnostics read the lattice during the run); then the headline
    configuration itself from an input script (device-only lattice, `cuda_create_sites`)."""
    rec = _run(n, sweeps, "yes")
    rec["cuda_host_sync_no"] = _run(n, sweeps, "no")
    try:
        rec["headline_from_script"] = headline_from_script()
    except Exception as e:  # noqa: BLE001 - the other two records stand on their own
        rec["headline_from_script"] = {"unavailable": str(e)[:300]}
    return rec


def headline_from_script(n=1024, sweeps=25):
    """BASELINE configs[2] (potts/neighonly, Q = 10000, sc/26n n^3, T = 0.5) driven by the plugin binary from an in.*
    script: the lattice is created (cuda_create_sites), initialised with the reference's `set site range` stream
    (cuda_set) and swept on the device; one stats line with the device energy per 5 sweeps.  Nothing crosses PCIe
    but counters and energies, so this is the device rate seen through the reference's own run loop and timers."""
    if not os.path.exists(SPK_CUDA):
        raise RuntimeError("plugin/_build/spk_cuda not built (needs /root/reference at build time)")
    dt = 1.0 / 26.0  # dt_sweep of potts/neighonly on sc/26n (src/app_potts_neighonly.cpp:62)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "in.headline")
        open(path, "w").write(HEADLINE_SCRIPT.format(n=n, stats=5 * dt, t=(sweeps - 0.5) * dt))
        t0 = time.perf_counter()
        out = subprocess.run([SPK_CUDA, "-in", path, "-log", "none"], cwd=tmp, capture_output=True, text=True, timeout=600)
        wall = time.perf_counter() - t0
    if out.returncode != 0 or "ERROR" in out.stdout:
        raise RuntimeError("plugin run failed: " + out.stdout[-400:] + out.stderr[-400:])
    loop = float(re.findall(r"Loop time of ([0-9.eE+-]+)", out.stdout)[-1])
    rows = [l.split() for l in out.stdout.splitlines() if re.match(r"^\s+[0-9.]+\s+\d+\s+\d+\s+\d+", l)]
    nsweeps = int(rows[-1][3])
    return {"value": float(n) ** 3 * nsweeps / loop / 1e9, "unit": "Gsites/s",
            "script": "potts/neighonly/cuda 10000, sc/26n %d^3, cuda_create_sites + cuda_set site range, sweep color, %d sweeps, "
                      "stats + diag_style energy/cuda every 5 sweeps" % (n, nsweeps),
            "loop_seconds": loop, "process_wall_seconds": wall, "energy_first_last": [float(rows[0][-1]), float(rows[-1][-1])],
            "definition": "site-updates / the reference's own 'Loop time' of the run; the lattice never leaves the device"}


def _run(n, sweeps, sync):
    if not os.path.exists(SPK_CUDA):
        raise RuntimeError("plugin/_build/spk_cuda not built (needs /root/reference at build time)")
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "in.e2e")
        open(path, "w").write(SCRIPT.format(n=n, sync=sync, t=0.5 * sweeps - 0.25))
        t0 = time.perf_counter()
        out = subprocess.run([SPK_CUDA, "-in", path, "-log", "none"], cwd=tmp, capture_output=True, text=True, timeout=1200)
        wall = time.perf_counter() - t0
    if out.returncode != 0 or "ERROR" in out.stdout:
        raise RuntimeError("plugin run failed: " + out.stdout[-400:] + out.stderr[-400:])
    loop = float(re.findall(r"Loop time of ([0-9.eE+-]+)", out.stdout)[-1])
    nsweeps = int([l.split() for l in out.stdout.splitlines() if re.match(r"^\s+[0-9.]+\s+\d+\s+\d+\s+\d+", l)][-1][3])
    sites = n * n * nsweeps
    return {"value": sites / loop / 1e9, "unit": "Gsites/s", "script": "ising/cuda sq/4n %dx%d, sweep color, %d sweeps, "
            "stats + diag_style energy/cuda every 20 sweeps, cuda_host_sync %s" % (n, n, nsweeps, sync),
            "loop_seconds": loop, "process_wall_seconds": wall,
            "h2d_bytes_per_step": 4.0 * n * n / nsweeps,
            "d2h_bytes_per_step": (4.0 * n * n * (nsweeps // 20 if sync == "yes" else 1) + 24 * (nsweeps // 20)) / nsweeps,
            "definition": "site-updates / the reference's own 'Loop time' of the run: lattice upload at the start of the run, "
                          "the sweeps, counters + device energy at every stats line, lattice download at every stats "
                          "line (cuda_host_sync yes) or only when the run ends (no)"}


if __name__ == "__main__":
    import json
    print(json.dumps(plugin_e2e()))
