"""CPU tier: the sources of libspparks_cuda compiled for the host against a CUDA emulator.

tests/host_shim/build_emu.py compiles spparks_b200/csrc — the ABI, the host logic and every kernel:
TMA streaming kernels, 16-bit and 1-bit shadows, colour-fused 2-D kernel, colour-pair launches,
candidate elimination, strip and generic kernels, masks, clusters — with g++ against
tests/host_shim/emu (CTAs on worker threads, the threads of a CTA as fibers that switch at
__syncthreads / warp collectives / mbarrier waits, bulk tensor copies with the driver's argument
checks and zero fill).  The GPU parity tests then run against that build in this container: the same
test functions, the same oracle and reference fixtures, through the same C ABI; the plugin tests run
the real SPPARKS binary on its `in.*` scripts with the emulated library in place of the CUDA one.

What this proves: the source of the kernels and of the host logic computes the reference's results.
What it cannot prove: anything about the hardware (timing, races inside a CTA, asynchrony of TMA and
of peer-memory stores, ptxas) — that stays with `-m gpu` on the B200.  The emulated build is test
infrastructure: spparks_b200/_lib.py refuses to load it unless a test opts in.
"""
import ctypes as C
import importlib.util
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPK_CUDA = os.path.join(ROOT, "plugin", "_build", "spk_cuda")


def _build_emu():
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "host_shim", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def emu_lib():
    mod = _build_emu()
    lib = mod.build()
    libdir = os.path.join(mod.BUILD, "libdir")  # the plugin binary asks the loader for "libspparks_cuda.so"
    os.makedirs(libdir, exist_ok=True)
    link = os.path.join(libdir, "libspparks_cuda.so")
    if not os.path.islink(link):
        os.symlink(os.path.join("..", os.path.basename(lib)), link)
    return lib


def _env(emu_lib):
    env = dict(os.environ, SPK_CUDA_LIB=emu_lib, SPK_CUDA_ALLOW_EMULATOR="1")
    libdir = os.path.join(os.path.dirname(emu_lib), "libdir")
    env["LD_LIBRARY_PATH"] = libdir + (":" + env["LD_LIBRARY_PATH"] if env.get("LD_LIBRARY_PATH") else "")
    return env


def _run_gpu_tests(emu_lib, args):
    """`pytest -m gpu` on the given files in a fresh interpreter that loads the emulated library"""
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"] + args
    r = subprocess.run(cmd, cwd=ROOT, env=_env(emu_lib), capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and " failed" not in r.stdout and " error" not in r.stdout, r.stdout[-2000:]
    return int(m.group(1))


def test_emulator_selftest(emu_lib):
    """tensor-map argument checks (box extent <= 256, 16-byte rules), zero fill outside the tensor,
    mbarrier phases, warp collectives, the restated integer intrinsics"""
    L = C.CDLL(emu_lib)
    assert L.spk_emu_selftest() == 0


@pytest.mark.parametrize("guard,inside,outside", [("after", 99, 112), ("before", 0, -1)])
def test_guard_pages_fault_on_accesses_outside_an_allocation(emu_lib, guard, inside, outside):
    """SPK_EMU_GUARD: the memcheck of the CPU tier really faults (112 = the first byte past the 16-byte rounded end)"""
    code = "import ctypes as C, sys; L = C.CDLL(%r); L.spk_emu_guard_probe.argtypes = [C.c_long]; print(L.spk_emu_guard_probe(int(sys.argv[1])))"
    env = dict(os.environ, SPK_EMU_GUARD=guard)
    ok = subprocess.run([sys.executable, "-c", code % emu_lib, str(inside)], env=env, capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0 and ok.stdout.strip() == str(0xCD), ok.stdout + ok.stderr
    bad = subprocess.run([sys.executable, "-c", code % emu_lib, str(outside)], env=env, capture_output=True, text=True, timeout=120)
    assert bad.returncode == -11, (bad.returncode, bad.stdout, bad.stderr)


@pytest.mark.parametrize("env,want", [({}, 0 * 2 + 1), ({"SPK_EMU_ORDER": "reverse"}, 63 * 2 + 1), ({"SPK_EMU_TMA": "lazy"}, 0 * 2 + 0)])
def test_order_and_lazy_copy_modes_do_what_they_say(emu_lib, env, want):
    """SPK_EMU_ORDER decides which CUDA thread of a CTA runs first; with SPK_EMU_TMA=lazy a bulk tensor copy is not
    visible until its barrier has been waited for (and the destination is poisoned meanwhile)"""
    code = "import ctypes as C; print(C.CDLL(%r).spk_emu_probe_modes())" % emu_lib
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and int(r.stdout.strip()) == want, r.stdout + r.stderr


def test_gpu_parity_tests_under_shuffled_threads_lazy_copies_and_guard_pages(emu_lib):
    """a slice of the parity tests (shadows interleaved with int32 calls, masks, several-tile lattices, ...) with
    the emulator's three checks on at once: the threads of every CTA started in a shuffled order, bulk tensor
    copies landing only when their barrier is waited for (a kernel that misses a barrier, reads a tile early or
    lets a copy overwrite planes still in use computes something else), and an inaccessible page behind every
    device allocation (no kernel reads or writes past the end of the lattice, its ghost layers, the shadows)"""
    saved = dict(os.environ)
    os.environ.update(SPK_EMU_ORDER="shuffle:7", SPK_EMU_TMA="lazy", SPK_EMU_GUARD="after")
    try:
        n = _run_gpu_tests(emu_lib, ["tests/test_gpu_parity.py", "-k", "interleaved or masked or large_lattices or extreme or elimination or nloop or edge_sizes or per_colour"])
    finally:
        os.environ.clear()
        os.environ.update(saved)
    assert n >= 30


def test_product_loader_refuses_the_emulated_build(emu_lib):
    code = ("import spparks_b200._lib as l\n"
            "try:\n    l.load()\nexcept l.SpkCudaError as e:\n    print('REFUSED', e)\nelse:\n    print('LOADED')\n")
    env = _env(emu_lib)
    del env["SPK_CUDA_ALLOW_EMULATOR"]
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert "REFUSED" in r.stdout and "LOADED" not in r.stdout, r.stdout + r.stderr


def test_gpu_parity_tests_pass_on_the_emulated_library(emu_lib):
    """every test of tests/test_gpu_parity.py: reference fixtures (injected half-sweeps, sector sets,
    masked passes, strict runs), Philox sweeps through every kernel tier, shadows interleaved with
    int32 entry points, candidate elimination, device `set site range`, error paths"""
    assert _run_gpu_tests(emu_lib, ["tests/test_gpu_parity.py"]) >= 380


def test_cluster_tests_pass_on_the_emulated_library(emu_lib):
    """union-find clustering (atomics, __match_any_sync, scans) against the reference's clusters"""
    assert _run_gpu_tests(emu_lib, ["tests/test_cluster.py", "-k", "not large"]) >= 12


@pytest.mark.skipif(not os.path.exists(SPK_CUDA), reason="plugin/_build/spk_cuda not built (needs the reference tree)")
def test_device_lattice_scripts_pass_on_the_emulated_library(emu_lib):
    """`cuda_create_sites`: scripts whose lattice lives on the device only print the table of the same script
    with a host lattice (tests/test_zz_device_lattice.py)"""
    assert _run_gpu_tests(emu_lib, ["tests/test_zz_device_lattice.py"]) >= 10


def test_more_parity_tests_pass_on_the_emulated_library(emu_lib):
    """tests/test_zz_more_parity.py: multi-step work items, several items per CTA, y extents of whole tiles,
    every face of every sector's mask border, the acceptance threshold table"""
    assert _run_gpu_tests(emu_lib, ["tests/test_zz_more_parity.py"]) >= 87


def test_device_pin_tests_pass_on_the_emulated_library(emu_lib):
    """device-side `pin f multi nthresh` against the reference's pinned sites (fixtures), the oracle, and through
    the plugin's `cuda_pin` against the reference binary's `pin`"""
    assert _run_gpu_tests(emu_lib, ["tests/test_zz_device_pin.py"]) >= (20 if os.path.exists(SPK_CUDA) else 17)


@pytest.mark.skipif(not os.path.exists(SPK_CUDA), reason="plugin/_build/spk_cuda not built (needs the reference tree)")
def test_plugin_scripts_pass_on_the_emulated_library(emu_lib):
    """the SPPARKS binary with the app-style plugin runs its in.* scripts against the emulated library"""
    r = subprocess.run(["ldd", SPK_CUDA], env=_env(emu_lib), capture_output=True, text=True)
    assert "host_shim/_build/libdir/libspparks_cuda.so" in r.stdout, r.stdout
    assert _run_gpu_tests(emu_lib, ["tests/test_plugin.py"]) >= 20


def test_smoke_passes_on_the_emulated_library(emu_lib):
    """__graft_entry__.smoke(): the driver's one small invocation of the hot path, checked against the oracle"""
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke(); print('SMOKE_DONE')"], cwd=ROOT,
                       env=_env(emu_lib), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SMOKE_DONE" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


MULTIRANK = [(2, "p2p"), (2, "nccl"), (4, "p2p")]


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


@pytest.fixture(scope="module")
def multirank_runs(emu_lib):
    """the three multi-rank runs, started together (they mostly wait for each other's ranks)"""
    import shutil
    import tempfile
    mod = _build_emu()
    procs = []
    for world, halo in MULTIRANK:
        arena = tempfile.mkdtemp(prefix="spk_emu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        env = _env(emu_lib)
        env.update(SPK_EMU_ARENA_DIR=arena, SPK_EMU_WORKERS="2", SPK_CUDA_NCCL_LIB=mod.NCCL_LIB, OMP_NUM_THREADS="1")
        env.pop("SPK_CUDA_NO_P2P", None)
        if halo == "nccl":
            env["SPK_CUDA_NO_P2P"] = "1"
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
               "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "emu_mgpu_worker.py")]
        procs.append((arena, subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    # next to them: a slice of the randomised multi-rank test at 3 ranks (tools/emu_fuzz_ranks.py)
    fuzz = subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "emu_fuzz_ranks.py"), "--ranks", "3", "--cases", "12", "--seed", "21",
                             "--port", str(_free_port())], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    out = {}
    try:
        out["fuzz3"] = fuzz.communicate(timeout=1500) + (fuzz.returncode,)
    except subprocess.TimeoutExpired:
        fuzz.kill()
        out["fuzz3"] = fuzz.communicate() + (-9,)
    for (world, halo), (arena, p) in zip(MULTIRANK, procs):
        try:
            out[(world, halo)] = p.communicate(timeout=1500)
        except subprocess.TimeoutExpired:
            p.kill()
            out[(world, halo)] = p.communicate()
        shutil.rmtree(arena, ignore_errors=True)
    return out


@pytest.mark.parametrize("world,halo", MULTIRANK)
def test_slabs_on_several_ranks_equal_one_rank_on_the_emulated_library(multirank_runs, world, halo):
    """The multi-GPU path, one process per rank: spparks_b200/selfcheck.py (the check tests/test_multigpu.py and
    bench.py run on the GPUs: spins, counters, energy, cluster statistics of 8 app x lattice cases plus banded
    lattices) with CUDA IPC emulated by shared file mappings — the neighbours' boundary stores and flag words go
    through them — and NCCL by tests/host_shim/emu/nccl_emu.cpp.  4 ranks: lower and upper neighbour differ."""
    stdout, stderr = multirank_runs[(world, halo)]
    assert "MGPU_ALL_OK" in stdout, stdout[-3000:] + stderr[-3000:]
    assert len(re.findall(r"^ok ", stdout, flags=re.M)) == 11, stdout


def test_random_cases_on_three_ranks(multirank_runs):
    """tools/emu_fuzz_ranks.py: random app / lattice / extents / tier / halo mode / call pattern, slabs gathered and
    compared with the oracle (spins, counters, energy, cluster statistics); a fixed seed here, thousands of cases by hand"""
    stdout, stderr, rc = multirank_runs["fuzz3"]
    assert rc == 0 and "3 ranks: 12 / 12 cases ok" in stdout, stdout[-3000:] + stderr[-2000:]


def test_random_cases_on_one_rank(emu_lib):
    """tools/emu_fuzz.py: random app / lattice / tile-edge extents / Q / T / kernel tier / sweeps, colour calls with
    nloop, sector sets, masked sector sets, against the oracle"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_fuzz.py"), "--cases", "80", "--seed", "11"], cwd=ROOT,
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "80 / 80 cases, 0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_random_call_sequences_on_one_handle(emu_lib):
    """tools/emu_fuzz_api.py: random ORDER of ABI calls on one handle (sweeps, colour sets, injected / strict sweeps,
    uploads, energies, clusters, temperature / seed changes, masks on / off, set site range, pins, randomize, ...),
    device == oracle after every call: the library's copies of the lattice stay consistent across entry points"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_fuzz_api.py"), "--cases", "20", "--ops", "16", "--seed", "5"],
                       cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "20 / 20 sequences, 0 failures" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


# ---- several MPI ranks: the reference's own parallel code and the plugin's multi-rank path -------------------------
SPK_CUDA_MPI = os.path.join(ROOT, "plugin", "_build", "spk_cuda_mpi")
SPK_REF_MPI = os.path.join(ROOT, "plugin", "_build", "spk_ref_mpi")
needs_mpi_emu = pytest.mark.skipif(not (os.path.exists(SPK_CUDA_MPI) and os.path.exists(SPK_REF_MPI)),
                                   reason="plugin/_build/spk_*_mpi not built (make -C plugin mpi; needs the reference tree)")


def _mpirun(nranks, binary, script, env, keep=None):
    """ranks = processes of this host (tests/host_shim/mpi/mpi_emu.cpp); returns rank 0's output"""
    import shutil
    import tempfile
    work = tempfile.mkdtemp(prefix="spk_mpi_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        path = os.path.join(work, "in.script")
        open(path, "w").write(script)
        procs = []
        for r in range(nranks):
            e = dict(env, SPK_MPI_RANK=str(r), SPK_MPI_SIZE=str(nranks), SPK_MPI_DIR=work, SPK_EMU_ARENA_DIR=work, SPK_EMU_WORKERS="2")
            cwd = work
            if keep is not None:  # every rank in its own directory, kept for the caller
                cwd = os.path.join(keep, "rank%d" % r)
                os.makedirs(cwd)
                e["SPK_TRACE_SETS"] = os.path.join(cwd, "sets.trace")
            args = [binary, "-in", path, "-log", "none"] if not binary.endswith("ref_driver_mpi") else [binary, path, "-quiet"]
            procs.append(subprocess.Popen(args, cwd=cwd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = []
        for p in procs:
            try:
                outs.append(p.communicate(timeout=600)[0])
            except subprocess.TimeoutExpired:
                p.kill()
                outs.append(p.communicate()[0] + "\nTIMEOUT")
        assert all(p.returncode == 0 for p in procs) and "ERROR" not in outs[0], "\n".join(o[-1500:] for o in outs)
        return outs[0]
    finally:
        shutil.rmtree(work, ignore_errors=True)


def _plugin_helpers():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_plugin as P
    return P


@needs_mpi_emu
def test_reference_on_several_ranks_through_the_mpi_emulation():
    """the MPI emulation itself: the UNMODIFIED reference on 2 and 4 ranks prints the serial table of a
    `sweep color/strict` script (create_sites, ghost communication, reductions: all through mpi_emu.cpp)"""
    P = _plugin_helpers()
    case = dict(P.STRICT_CASES[0])
    script = P.STRICT.format(**dict(case, app="%s %d" % (case["app"], case["q"])))
    serial = P.stats_rows(P.run(P.SPK_REF, script))
    for n in (2, 4):
        par = script.replace("create_box box", "processors 1 1 %d\ncreate_box box" % n)
        assert P.stats_rows(_mpirun(n, SPK_REF_MPI, par, dict(os.environ))) == serial


@needs_mpi_emu
@pytest.mark.parametrize("icase,nranks", [(0, 2), (0, 4), (2, 2), (3, 2), (4, 2), (6, 2)])
def test_plugin_on_several_mpi_ranks_prints_the_reference_table(emu_lib, icase, nranks):
    """The /cuda app styles on more than one MPI rank (never executed before: the image has no MPI): slab check,
    NCCL id broadcast, per-colour ghost refresh, per-slab site seeds, device energies reduced over the ranks.
    `sweep color/strict` scripts print the serial reference's table row for row."""
    P = _plugin_helpers()
    case = dict(P.STRICT_CASES[icase])
    if case["dim"] == 3:
        case["zr"] = "0 16"     # slabs of an even number of planes on 2 and 4 ranks
    else:
        case["ny"] = 32         # 2-D slabs: an even number of rows, at least the 4 ghost rows
    def script(style):
        app = style if case["app"].startswith("ising") else "%s %d" % (style, case["q"])
        return P.STRICT.format(**dict(case, app=app))
    ref = P.stats_rows(P.run(P.SPK_REF, script(case["app"])))
    grid = "1 1 %d" % nranks if case["dim"] == 3 else "1 %d 1" % nranks
    par = script(case["app"] + "/cuda").replace("create_box box", "processors %s\ncreate_box box" % grid)
    par = par.replace("diag_style energy", "diag_style energy/cuda")
    mod = _build_emu()
    env = dict(_env(emu_lib), SPK_CUDA_NCCL_LIB=mod.NCCL_LIB)
    assert P.stats_rows(_mpirun(nranks, SPK_CUDA_MPI, par, env)) == ref and len(ref) >= 4


@needs_mpi_emu
@pytest.mark.parametrize("create", ["create_sites box", "cuda_create_sites"])
def test_plugin_philox_run_on_two_ranks_equals_one_rank(emu_lib, create):
    """production mode (`sweep color`, `sector no`): the random stream is keyed on the global site, so two ranks
    print the table of one rank; with the host lattice and with the device-only lattice (cuda_set / cuda_pin
    collective on the slabs, cluster statistics merged across the slab boundary)"""
    P = _plugin_helpers()
    script = """seed 56789
app_style potts/pin/cuda 12
dimension 3
lattice sc/26n 1.0
region box block 0 16 0 8 0 8
{procs}create_box box
{create}
cuda_set site range 1 12
cuda_pin 0.05 0 0
sweep color
sector no
temperature 0.9
diag_style energy/cuda
{cluster}
stats 0.25
run 1.0
run 0.5
"""
    mod = _build_emu()
    env = dict(_env(emu_lib), SPK_CUDA_NCCL_LIB=mod.NCCL_LIB)
    cluster = "diag_style cluster/cuda" if create == "cuda_create_sites" else ""
    one = P.stats_rows(_mpirun(1, SPK_CUDA_MPI, script.format(procs="", create=create, cluster=cluster), env))
    two = P.stats_rows(_mpirun(2, SPK_CUDA_MPI, script.format(procs="processors 1 1 2\n", create=create, cluster=cluster), env))
    assert len(one) >= 6 and two == one


@needs_mpi_emu
@pytest.mark.parametrize("lat,dim,n,procs,sector", [
    # (with slabs of the last dimension the reference only admits 8 sectors in 3-D and 4 in 2-D, src/app_lattice.cpp:168-185)
    ("sc/26n", 3, (8, 8, 16), "1 1 2", "yes"), ("sc/26n", 3, (8, 12, 16), "1 1 4", "8"), ("sc/6n", 3, (6, 6, 8), "1 1 2", "yes"),
    ("sq/4n", 2, (12, 16, 1), "1 2 1", "yes"), ("sq/8n", 2, (8, 24, 1), "1 3 1", "4"), ("tri", 2, (8, 16, 1), "1 2 1", "yes"),
], ids=lambda v: str(v).replace(" ", ""))
def test_plugin_sector_boxes_equal_the_reference_sets_on_every_rank(emu_lib, tmp_path, lat, dim, n, procs, sector):
    """`sweep color` + `sector yes|N` on several ranks: the reference cuts every rank's OWN sub-domain at its midpoints
    (create_set, src/app_lattice.cpp:954-1026).  The loop sets the plugin hands to the library (logged by an
    LD_PRELOAD interposer on spk_cuda_sweep_color) hold, rank by rank and in the reference's order, exactly the sites
    of the reference's own sets (ref_driver @sets / @ids on the same ranks)."""
    if not os.path.exists(os.path.join(ROOT, "plugin", "_build", "ref_driver_mpi")):
        pytest.skip("plugin/_build/ref_driver_mpi not built")
    mod = _build_emu()
    trace_lib = os.path.join(mod.BUILD, "libtrace_sets.so")
    subprocess.check_call(["gcc", "-O1", "-fPIC", "-shared", os.path.join(ROOT, "tests", "host_shim", "trace", "trace_sets.c"), "-o", trace_lib, "-ldl"])
    nranks = int(procs.split()[dim - 1])
    zr = "0 %d" % n[2] if dim == 3 else "-0.5 0.5"
    head = ("seed 4711\napp_style %s\ndimension %d\nlattice %s 1.0\nregion box block 0 %d 0 %d %s\nprocessors %s\ncreate_box box\n"
            "create_sites box\nset site range 1 5\nsweep color\nsector %s\ntemperature 0.5\n")
    ref_dir, gpu_dir = str(tmp_path / "ref"), str(tmp_path / "gpu")
    os.makedirs(ref_dir)
    os.makedirs(gpu_dir)
    _mpirun(nranks, os.path.join(ROOT, "plugin", "_build", "ref_driver_mpi"),
            head % ("potts 5", dim, lat, n[0], n[1], zr, procs, sector) + "run 0.0\n@ids ids.bin\n@sets sets.bin\n", dict(os.environ), keep=ref_dir)
    P = _plugin_helpers()
    dt = 1.0 / 5
    env = dict(_env(emu_lib), SPK_CUDA_NCCL_LIB=mod.NCCL_LIB, LD_PRELOAD=trace_lib)
    _mpirun(nranks, SPK_CUDA_MPI, head % ("potts/cuda 5", dim, lat, n[0], n[1], zr, procs, sector) + "run %g\n" % (0.5 * dt), env, keep=gpu_dir)
    nb = 2 if lat == "tri" else 1
    o = __import__("oracle_py").Oracle("potts", lat, n[0], n[1], n[2], 5, 0.5)
    colors = o.colors()
    cell = np.arange(o.nsites) // nb
    ci, cj, ck = cell % n[0], (cell // n[0]) % n[1], cell // (n[0] * n[1])
    for r in range(nranks):
        ids = np.fromfile(os.path.join(ref_dir, "rank%d" % r, "ids.bin"), dtype=np.int32)
        raw = np.fromfile(os.path.join(ref_dir, "rank%d" % r, "sets.bin"), dtype=np.int32)
        nset, nsector, ncolors = int(raw[0]), int(raw[1]), int(raw[2])
        sets, pos = [], 3
        for _ in range(nset):
            cnt = int(raw[pos])
            sets.append(np.sort(ids[raw[pos + 1:pos + 1 + cnt]] - 1))
            pos += 1 + cnt
        assert nset == nsector + nsector * ncolors and nsector > 1
        calls = [tuple(int(x) for x in line.split()) for line in open(os.path.join(gpu_dir, "rank%d" % r, "sets.trace"))]
        assert len(calls) == nsector * ncolors   # one sweep: every (sector, colour) set once, sector-major
        for k, (sweep, color, nloop, ilo, ihi, jlo, jhi, klo, khi) in enumerate(calls):
            isec, ic = divmod(k, ncolors)
            assert color == ic + 1 and nloop == 1
            inbox = (ci >= ilo) & (ci < ihi) & (cj >= jlo) & (cj < jhi) & (ck >= klo) & (ck < khi)
            mine = np.flatnonzero(inbox & (colors == color))
            assert np.array_equal(mine, sets[nsector + isec * ncolors + ic]), (r, isec, ic)


TIER_SCRIPT = r"""
import ctypes as C, json, os, sys
import numpy as np
from spparks_b200 import LatticeSweep
L = C.CDLL(os.environ["SPK_CUDA_LIB"])
L.spk_emu_launch_log.restype = C.c_size_t
def log():
    n = L.spk_emu_launch_log(None, 0)
    b = C.create_string_buffer(n)
    L.spk_emu_launch_log(b, n)
    return dict((k, int(v)) for k, v in (line.split("\t") for line in b.value.decode().splitlines()))
out = {}
for key, model, lat, n, q, env in json.loads(sys.argv[1]):
    for k in [k for k in os.environ if k.startswith("SPK_CUDA_") and k not in ("SPK_CUDA_LIB", "SPK_CUDA_ALLOW_EMULATOR")]:
        del os.environ[k]
    os.environ.update(env)
    L.spk_emu_launch_log_reset()
    s = LatticeSweep(model, lat, *n, nspins=q, temperature=0.7, seed=5)
    nb = 2 if lat == "tri" else 1
    s.upload(np.random.default_rng(1).integers(1, (2 if model.startswith("ising") else q) + 1, size=n[0] * n[1] * n[2] * nb, dtype=np.int32))
    s.sweeps(0, 2)
    s.download()
    out[key] = log()
print("TIERS " + json.dumps(out))
"""


def test_every_kernel_tier_is_the_one_that_runs(emu_lib):
    """the instantiation each configuration launches (demangled kernel symbols recorded by the emulator):
    a parity test that passes on a silent fallback to the generic kernel would say nothing about the tuned ones.
    Template arguments: lattice 2 sq/4n, 3 sq/8n, 4 tri, 5 sc/6n, 6 sc/26n; model 0 ising, 1 potts, 2 potts/neighonly, 3 potts/pin."""
    cases = [
        ["shadow16", "potts/pin", "sc/26n", (32, 16, 12), 50, {}],
        ["int32stream", "potts", "sc/26n", (32, 16, 12), 50, {"SPK_CUDA_NO_SHADOW16": "1"}],
        ["stream32", "potts", "sc/26n", (32, 16, 12), 70000, {}],
        ["pair", "potts", "sc/26n", (32, 16, 12), 50, {"SPK_CUDA_PAIR": "1", "SPK_CUDA_NO_SHADOW16": "1"}],
        ["chain", "potts/neighonly", "sc/26n", (32, 12, 14), 100, {"SPK_CUDA_NEIGH_MODE": "chain"}],
        ["elimination", "potts/neighonly", "sc/26n", (32, 12, 14), 100, {"SPK_CUDA_NEIGH_MODE": "adaptive"}],
        ["sc6n", "ising", "sc/6n", (16, 12, 10), 2, {}],
        ["bits", "ising", "sq/4n", (256, 34, 1), 2, {}],
        ["fused2d", "potts", "tri", (252, 66, 1), 9, {}],
        ["strip", "potts", "sq/8n", (72, 18, 1), 7, {"SPK_CUDA_NO_FUSED": "1"}],
        ["generic", "potts", "sq/8n", (34, 22, 1), 7, {}],
    ]
    import json
    r = subprocess.run([sys.executable, "-c", TIER_SCRIPT, json.dumps(cases)], cwd=ROOT, env=_env(emu_lib), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    tiers = json.loads(re.search(r"TIERS (.*)", r.stdout).group(1))
    want = {
        "shadow16": "spk::sweep_stream16_kernel<3>",
        "int32stream": "spk::sweep_stream_kernel<6, 1, true, true, false, 0>",
        "stream32": "spk::sweep_stream_kernel<6, 1, true, false, false, 0>",   # Q > 65533: 32-bit compares
        "pair": "spk::sweep_stream_kernel<6, 1, true, true, true, 0>",
        "chain": "spk::sweep_stream_kernel<6, 2, true, true, false, 0>",
        "elimination": "spk::sweep_stream_kernel<6, 2, true, true, false, 1>",
        "sc6n": "spk::sweep_stream_kernel<5, 0, true, false, false, 0>",
        "bits": "spk::sweep_bits2d_kernel<0>",
        "fused2d": "spk::sweep_tile2d_kernel<4, 1>",
        "strip": "spk::sweep_strip_kernel<3, 1>",
        "generic": "spk::sweep_generic_kernel<3, 1, 0>",
    }
    for key, kernel in want.items():
        assert kernel in tiers[key], (key, tiers[key])
        sweepers = [k for k in tiers[key] if "sweep_" in k]
        assert sweepers == [kernel], (key, tiers[key])  # nothing else swept
    assert "spk::edge_pair_kernel<1, true>" in tiers["pair"]
    assert tiers["shadow16"]["shadow_pack_kernel"] == 1 and tiers["shadow16"]["shadow_unpack_kernel"] == 1
    assert tiers["bits"]["bits2d_pack_kernel"] == 1 and tiers["bits"]["bits2d_unpack_kernel"] == 1
