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


def plugin_e2e(n=4096, sweeps=200):
    """Default plugin behaviour (host copy refreshed at every stats line) and, next to it,
    `cuda_host_sync no` (only /cuda diagnostics read the lattice during the run)."""
    rec = _run(n, sweeps, "yes")
    rec["cuda_host_sync_no"] = _run(n, sweeps, "no")
    return rec


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
