"""Randomised differential test of the library against the oracle, on the CPU (emulated build).

    python tools/emu_fuzz.py [--cases N] [--seed S] [--max-sites M] [--keep-going]

Every case draws an app, a lattice, extents (ragged and tile-edge sizes on purpose), Q, T, a kernel
tier (environment switches of the library), the number of sweeps and optionally sectors / masks /
nloop > 1, runs it through the C ABI of tests/host_shim/_build/libspparks_cuda_emu.so and compares
spins, counters and energy with oracle/ evaluating the same Philox stream.  The emulated build is
test infrastructure (tests/host_shim/README.md); this script only drives it harder than the fixed
cases of tests/test_gpu_parity.py do.  It re-executes itself per case batch because the library reads
its switches when a handle is created and the emulator lives in the process.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

LATS = ["sq/4n", "sq/8n", "tri", "sc/6n", "sc/26n"]
MODELS = ["ising", "ising/single", "potts", "potts/pin", "potts/neighonly", "potts/neigh"]
TIERS = {
    "tuned": {},
    "chain": {"SPK_CUDA_NEIGH_MODE": "chain"},
    "adaptive": {"SPK_CUDA_NEIGH_MODE": "adaptive"},
    "tuned32": {"SPK_CUDA_NO_PK16": "1"},
    "pair": {"SPK_CUDA_PAIR": "1"},
    "strip": {"SPK_CUDA_NO_STREAM": "1", "SPK_CUDA_NO_FUSED": "1"},
    "generic": {"SPK_CUDA_FORCE_GENERIC": "1"},
    "nobits": {"SPK_CUDA_NO_BITS": "1"},
    "int32stream": {"SPK_CUDA_NO_SHADOW16": "1"},
}
# extents that sit on the edges of the kernels' tiles (x tiles of 128 / 240 / 256 sites, 8- and 64-cell groups)
X_EDGES = [2, 4, 6, 8, 10, 12, 16, 24, 30, 32, 40, 62, 64, 66, 72, 120, 126, 128, 130, 132, 136, 192, 240, 248, 254, 256, 258, 264,
           272, 384, 480, 496, 512, 520]
Y_EDGES = [2, 4, 6, 8, 10, 14, 16, 18, 30, 32, 34, 36, 62, 64, 66, 68, 70, 96, 126, 128, 130]
Z_EDGES = [2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 34]


def draw_case(rng, max_sites):
    lat = LATS[int(rng.integers(len(LATS)))]
    model = MODELS[int(rng.integers(len(MODELS)))]
    dim3 = lat.startswith("sc")
    while True:
        nx = int(rng.choice(X_EDGES))
        ny = int(rng.choice(Y_EDGES))
        if lat == "tri" and rng.random() < 0.3:
            ny += 1  # tri only needs nx even
        nz = int(rng.choice(Z_EDGES)) if dim3 else 1
        nb = 2 if lat == "tri" else 1
        if nx * ny * nz * nb <= max_sites:
            break
    q = 2 if model.startswith("ising") else int(rng.choice([2, 3, 5, 20, 100, 10000, 65533, 65534, 65535, 70000]))
    T = float(rng.choice([0.0, 0.3, 0.8, 1.5, 6.0]))
    tier = "tuned" if rng.random() < 0.4 else list(TIERS)[int(rng.integers(len(TIERS)))]
    mode = str(rng.choice(["sweeps", "sweeps", "sweeps", "colors", "sectors", "mask"]))
    return {"lat": lat, "model": model, "n": [nx, ny, nz], "q": q, "T": T, "tier": tier, "mode": mode,
            "nsweeps": int(rng.integers(1, 4)), "nloop": int(rng.choice([1, 1, 1, 2, 3])),
            "nsector": int(rng.choice([2, 4, 8]) if dim3 else rng.choice([2, 4])), "seed": int(rng.integers(1, 2**40)),
            "start": str(rng.choice(["random", "random", "grains", "few"])), "sweep0": int(rng.integers(0, 1000)),
            # shape of the persistent launches: multi-step work items / several items per CTA (model_entry.cuh)
            "chunks": int(rng.choice([0, 0, 1, 2])), "ctas": int(rng.choice([0, 0, 1, 3]))}


def run_case(c):
    """returns None or a description of the mismatch"""
    import numpy as np
    import oracle_py as O
    from spparks_b200 import LatticeSweep
    for k in [k for k in os.environ if k.startswith("SPK_CUDA_") and k not in ("SPK_CUDA_LIB", "SPK_CUDA_ALLOW_EMULATOR")]:
        del os.environ[k]
    os.environ.update(TIERS[c["tier"]])
    if c.get("chunks"):
        os.environ["SPK_CUDA_MAX_PLANE_CHUNKS"] = str(c["chunks"])
    if c.get("ctas"):
        os.environ["SPK_CUDA_MAX_CTAS"] = str(c["ctas"])
    model, lat, n, q, T = c["model"], c["lat"], c["n"], c["q"], c["T"]
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
    rng = np.random.default_rng(c["seed"])
    top = 2 if model.startswith("ising") else q
    if c["start"] == "random":
        spins = rng.integers(1, top + 1, size=o.nsites, dtype=np.int32)
    elif c["start"] == "few":
        spins = rng.integers(1, min(top, 3) + 1, size=o.nsites, dtype=np.int32)
    else:  # blocks of equal spins
        nb = o.nbasis
        idx = np.arange(o.nsites) // nb
        i, j, k = idx % n[0], (idx // n[0]) % n[1], idx // (n[0] * n[1])
        spins = (1 + ((i // 5) * 7 + (j // 3) * 13 + (k // 2) * 29) % top).astype(np.int32)
    if model == "potts/pin":
        spins[rng.random(o.nsites) < 0.07] = q + 1
    s = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=c["seed"])
    s.upload(spins)
    want = spins.copy()
    nacc = natt = 0
    mode = c["mode"]
    if mode == "sweeps":
        for sw in range(c["nsweeps"]):
            nacc += o.sweep_philox(want, c["seed"], c["sweep0"] + sw)
        natt = c["nsweeps"] * o.nsites
        s.sweeps(c["sweep0"], c["nsweeps"])
    elif mode == "colors":  # per-colour calls with nloop passes
        colors = o.colors()
        for sw in range(c["nsweeps"]):
            for color in range(1, o.ncolors + 1):
                for l in range(c["nloop"]):
                    nacc += o.halfsweep_philox(want, c["seed"], c["sweep0"] + sw, color, iloop=l)
                natt += c["nloop"] * int((colors == color).sum())
                s.sweep_color(c["sweep0"] + sw, color, None, c["nloop"])
    else:  # sector x colour sets, optionally masked
        ns = c["nsector"]
        colors = o.colors()
        cell = np.arange(o.nsites) // o.nbasis
        ci, cj, ck = cell % n[0], (cell // n[0]) % n[1], cell // (n[0] * n[1])
        masked = mode == "mask"
        mask = np.zeros(o.nsites, dtype=np.uint8)
        if masked:
            s.mask_enable(True)
        for sw in range(c["nsweeps"]):
            for isec in range(ns):
                box = o.sector_box(ns, isec)
                inbox = (ci >= box[0]) & (ci < box[1]) & (cj >= box[2]) & (cj < box[3]) & (ck >= box[4]) & (ck < box[5])
                if masked:
                    o.boundary_clear_mask(mask, box)
                    s.mask_clear_border(box, 0)
                for color in range(1, o.ncolors + 1):
                    if masked:
                        nacc += o.halfsweep_philox_mask(want, mask, c["seed"], c["sweep0"] + sw, color, 0, box)
                    else:
                        nacc += o.halfsweep_philox(want, c["seed"], c["sweep0"] + sw, color, 0, box)
                    natt += int((inbox & (colors == color)).sum())
                    s.sweep_color(c["sweep0"] + sw, color, box, 1)
        if masked and not np.array_equal(s.download_mask(), mask):
            return "mask bytes differ"
    got = s.download()
    if not np.array_equal(got, want):
        bad = np.flatnonzero(got != want)
        return "spins differ at %d sites, first %d" % (bad.size, int(bad[0]))
    cnt = s.counters()
    if mode != "mask" and cnt != (nacc, natt):
        return "counters %r, oracle %r" % (cnt, (nacc, natt))
    if mode == "mask" and cnt[0] != nacc:
        return "accepted %r, oracle %r" % (cnt[0], nacc)
    if s.energy() != o.total_energy(want):
        return "energy differs"
    path = s.sweep_path()[0]
    s.close()
    return None, path


def child(cases):
    for c in cases:
        try:
            r = run_case(c)
        except Exception as e:  # noqa: BLE001 - the report matters, not the type
            r = "exception: %r" % (e,)
        if isinstance(r, tuple):
            print("FUZZ_OK " + json.dumps({"path": r[1]}), flush=True)
        else:
            print("FUZZ_FAIL " + json.dumps({"case": c, "what": r}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-sites", type=int, default=300000)
    ap.add_argument("--batch", type=int, default=25)
    ap.add_argument("--child", default=None)
    args = ap.parse_args()
    if args.child:
        child(json.loads(args.child))
        return 0
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "host_shim", "build_emu.py"))
    be = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(be)
    lib = be.build()
    env = dict(os.environ, SPK_CUDA_LIB=lib, SPK_CUDA_ALLOW_EMULATOR="1")
    rng = np.random.default_rng(args.seed)
    cases = [draw_case(rng, args.max_sites) for _ in range(args.cases)]
    fails, paths, done = [], {}, 0
    for b in range(0, len(cases), args.batch):
        batch = cases[b:b + args.batch]
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", json.dumps(batch)], env=env, cwd=ROOT,
                           capture_output=True, text=True)
        lines = [l for l in r.stdout.splitlines() if l.startswith("FUZZ_")]
        for l in lines:
            if l.startswith("FUZZ_FAIL"):
                fails.append(json.loads(l[10:]))
            else:
                p = json.loads(l[8:])["path"]
                paths[p] = paths.get(p, 0) + 1
        done += len(lines)
        if len(lines) != len(batch):  # the child died (emulator abort = deadlock or invalid launch)
            fails.append({"case": batch[len(lines)], "what": "process died: " + r.stderr[-500:]})
            done += 1
            rest = batch[len(lines) + 1:]
            if rest:
                cases[b + args.batch:b + args.batch] = rest
        print("%d / %d cases, %d failures" % (done, len(cases), len(fails)), flush=True)
    print("kernels swept with:", json.dumps(paths, sort_keys=True))
    for f in fails:
        print("FAIL", json.dumps(f))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
