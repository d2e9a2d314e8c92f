"""Randomised CALL SEQUENCES through the C ABI against the oracle, on the CPU (emulated build).

    python tools/emu_fuzz_api.py [--cases N] [--ops K] [--seed S]

tools/emu_fuzz.py varies the lattice and runs one kind of call; this one keeps the lattice small and
varies the ORDER of calls on one handle: sweeps, single colour sets (boxes, nloop), injected and strict
sweeps, uploads, downloads, energy, per-site energies, cluster statistics and ids, temperature and seed
changes, masks on / off, `set site range`, pins, randomize, counters, digests.  The library keeps several
copies of the lattice (int32, 16-bit / 1-bit shadows, mask bytes, ghost flags, a variant statistic): every
entry point has to leave them consistent for every other one.  (spk_cuda_mask_enable(h, 0) freeing the
shadows was a defect of exactly this kind.)  After every step the device is compared with the oracle.
Test infrastructure only (tests/host_shim/README.md).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# lattices on which the shadows / tuned kernels engage, and two that take the generic path
SHAPES = [("sc/26n", (16, 8, 8)), ("sc/26n", (8, 8, 4)), ("sc/26n", (24, 34, 6)), ("sc/6n", (16, 8, 6)), ("sq/4n", (64, 16, 1)),
          ("sq/4n", (10, 8, 1)), ("sq/8n", (16, 12, 1)), ("tri", (16, 9, 1)), ("tri", (6, 4, 1))]
MODELS = ["ising", "ising/single", "potts", "potts/pin", "potts/neighonly", "potts/neigh"]
OPS = ["sweeps", "sweeps", "sweeps", "color", "color", "sectors", "upload", "download", "energy", "site_energy", "clusters",
       "cluster_ids", "temperature", "seed", "mask_on", "mask_off", "set_range", "pin", "randomize", "counters", "digest",
       "injected", "strict", "reset_counters"]


def draw_case(rng, nops):
    lat, n = SHAPES[int(rng.integers(len(SHAPES)))]
    model = MODELS[int(rng.integers(len(MODELS)))]
    q = 2 if model.startswith("ising") else int(rng.choice([3, 9, 200, 65533, 70000]))
    ops = []
    for _ in range(nops):
        op = OPS[int(rng.integers(len(OPS)))]
        ops.append({"op": op, "a": int(rng.integers(0, 1 << 30)), "b": int(rng.integers(0, 1 << 30)), "u": float(rng.random())})
    return {"lat": lat, "n": list(n), "model": model, "q": q, "T": float(rng.choice([0.0, 0.5, 1.2])), "seed": int(rng.integers(1, 2**40)),
            "ops": ops, "tier": str(rng.choice(["tuned", "tuned", "tuned", "generic", "strip", "chunked"]))}


TIER_ENV = {"tuned": {}, "generic": {"SPK_CUDA_FORCE_GENERIC": "1"}, "strip": {"SPK_CUDA_NO_STREAM": "1", "SPK_CUDA_NO_FUSED": "1"},
            "chunked": {"SPK_CUDA_MAX_PLANE_CHUNKS": "1", "SPK_CUDA_MAX_CTAS": "2"}}


def run_case(c):
    import numpy as np
    import oracle_py as O
    from spparks_b200 import LatticeSweep
    for k in [k for k in os.environ if k.startswith("SPK_CUDA_") and k not in ("SPK_CUDA_LIB", "SPK_CUDA_ALLOW_EMULATOR")]:
        del os.environ[k]
    os.environ.update(TIER_ENV[c["tier"]])
    model, lat, n, q = c["model"], c["lat"], c["n"], c["q"]
    T, seed = c["T"], c["seed"]
    top = 2 if model.startswith("ising") else q
    rng = np.random.default_rng(seed)
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
    nsec = 8 if lat.startswith("sc") else 4
    if lat == "tri" and n[1] % 2:
        nsec = 2  # (the plugin refuses y sectors on odd tri rows; the boxes of the oracle are still boxes for 2 sectors)
    colors = o.colors()
    want = rng.integers(1, top + 1, size=o.nsites, dtype=np.int32)
    s = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=seed)
    s.upload(want)
    mask = None          # oracle's mask bytes while masks are on
    nacc = natt = 0
    sweep = 100
    siteseeds = None

    def fresh_oracle(temp):
        return O.Oracle(model, lat, n[0], n[1], n[2], q, temp)

    def check(what):
        got = s.download()
        if not np.array_equal(got, want):
            bad = np.flatnonzero(got != want)
            return "%s: spins differ at %d sites, first %d" % (what, bad.size, int(bad[0]))
        if mask is not None and not np.array_equal(s.download_mask(), mask):
            return "%s: mask bytes differ" % what
        if s.counters() != (nacc, natt):
            return "%s: counters %r, expected %r" % (what, s.counters(), (nacc, natt))
        return None

    for step, opd in enumerate(c["ops"]):
        op, a, b, u = opd["op"], opd["a"], opd["b"], opd["u"]
        tag = "step %d %s" % (step, op)
        if op == "sweeps":
            k = 1 + a % 3
            if mask is None:
                for i in range(k):
                    nacc += o.sweep_philox(want, seed, sweep + i)
            else:  # colour sets as loop sets: all border, so every set is unmasked before it is swept
                for i in range(k):
                    for col in range(1, o.ncolors + 1):
                        mask[colors == col] = 0
                        nacc += o.halfsweep_philox_mask(want, mask, seed, sweep + i, col, 0, None)
            natt += k * o.nsites
            s.sweeps(sweep, k)
            sweep += k
        elif op == "color":
            col, nloop = 1 + a % o.ncolors, 1 + b % 3
            if mask is None:
                for l in range(nloop):
                    nacc += o.halfsweep_philox(want, seed, sweep, col, iloop=l)
                natt += nloop * int((colors == col).sum())
                s.sweep_color(sweep, col, None, nloop)
            else:
                mask[colors == col] = 0
                s.mask_clear_border(None, col)
                nacc += o.halfsweep_philox_mask(want, mask, seed, sweep, col, 0, None)
                natt += int((colors == col).sum())
                s.sweep_color(sweep, col, None, 1)
            sweep += 1
        elif op == "sectors":
            cell = np.arange(o.nsites) // o.nbasis
            ci, cj, ck = cell % n[0], (cell // n[0]) % n[1], cell // (n[0] * n[1])
            for isec in range(nsec):
                box = o.sector_box(nsec, isec)
                inbox = (ci >= box[0]) & (ci < box[1]) & (cj >= box[2]) & (cj < box[3]) & (ck >= box[4]) & (ck < box[5])
                if mask is not None:
                    o.boundary_clear_mask(mask, box)
                    s.mask_clear_border(box, 0)
                for col in range(1, o.ncolors + 1):
                    if mask is not None:
                        nacc += o.halfsweep_philox_mask(want, mask, seed, sweep, col, 0, box)
                    else:
                        nacc += o.halfsweep_philox(want, seed, sweep, col, 0, box)
                    natt += int((inbox & (colors == col)).sum())
                    s.sweep_color(sweep, col, box, 1)
            sweep += 1
        elif op == "injected" and mask is None:
            col = 1 + a % o.ncolors
            kk = np.random.default_rng(a).integers(1, O.IM - 1, size=o.nsites, dtype=np.int32)
            u1, u2 = O.injected_uniforms(kk)
            nacc += o.halfsweep_injected(want, col, u1, u2)
            natt += int((colors == col).sum())
            s.sweep_color_injected(col, u1, u2)
        elif op == "strict" and mask is None:
            if siteseeds is None:
                siteseeds = o.strict_siteseeds(O.mars(1 + a % 900000000))
                s.set_siteseeds(siteseeds)
            nacc += o.sweeps_strict(want, siteseeds, 1)
            natt += o.nsites
            for col in range(1, o.ncolors + 1):
                s.sweep_color_strict(col, None, 1)
            if not np.array_equal(s.get_siteseeds(), siteseeds):
                return tag + ": site seeds differ"
        elif op == "upload":
            want = np.random.default_rng(a).integers(1, top + 1, size=o.nsites, dtype=np.int32)
            if model == "potts/pin" and u < 0.5:
                want[np.random.default_rng(b).random(o.nsites) < 0.1] = q + 1
            s.upload(want)
        elif op == "download":
            pass  # check() below downloads
        elif op == "energy":
            if s.energy() != o.total_energy(want):
                return tag + ": energy differs"
        elif op == "site_energy":
            e = s.site_energy()
            if int(e.sum()) != int(o.total_energy(want)):
                return tag + ": per-site energies do not add up to the total"
            for site in (0, o.nsites // 3, o.nsites - 1):
                if e[site] != O.lib().orc_site_energy(o.ctx, want.ctypes.data_as(O.C.POINTER(O.C.c_int32)), site):
                    return tag + ": per-site energy of site %d differs" % site
        elif op == "clusters":
            got, oc = s.cluster_stats(), o.clusters(want)
            if got[0] != oc[0] or abs(got[1] - oc[1]) > 1e-9 * oc[1] or abs(got[2] - oc[2]) > 1e-9 * oc[2]:
                return tag + ": cluster statistics %r, oracle %r" % (got, oc[:3])
        elif op == "cluster_ids":
            ids, vol = s.cluster_ids()
            oc = o.clusters(want)
            if not np.array_equal(ids, oc[3]) or not np.array_equal(vol, oc[4]):
                return tag + ": cluster ids / volumes differ"
        elif op == "temperature":
            T = [0.0, 0.3, 0.9, 2.0][a % 4]
            o = fresh_oracle(T)
            s.set_temperature(T)
        elif op == "seed":
            seed = 1 + (a * 7919 + b) % (2**40)
            s.set_seed(seed)
        elif op == "mask_on":
            if mask is None:
                mask = np.zeros(o.nsites, dtype=np.uint8)
            else:
                mask[:] = 0  # enabling again clears the masks (AppLattice::init: mask[i] = 0)
            s.mask_enable(True)
        elif op == "mask_off":
            mask = None
            s.mask_enable(False)
        elif op == "set_range":
            park = 1 + a % (O.IM - 2)
            lo = 1 + b % top
            hi = lo + (a % (top - lo + 1))
            state, am = park, 1.0 / O.IM
            vals = np.empty(o.nsites, dtype=np.int32)
            for i in range(o.nsites):
                state = state * 16807 % O.IM
                vals[i] = lo + min(int(state * am * (hi - lo + 1)), hi - lo)
            want = vals
            s.set_site_range(park, lo, hi)
        elif op == "pin" and model == "potts/pin":
            m = O.mars(1 + a % 900000000)
            park = int(O.lib().orc_mars_uniform(O.C.byref(m)) * O.IM) or 1
            frac = [0.02, 0.1, 0.3][b % 3]
            before = int((want > q).sum())
            if before > int(frac * o.nsites) + o.nsites // 2:
                continue
            o.pin_create(want, frac, 0, 0, O.mars(1 + a % 900000000))
            s.pin_create(park, frac)
        elif op == "randomize":
            s.randomize(a, 1, top)  # the device's own stream: take its result as the new truth
            want = s.download().copy()
            if want.min() < 1 or want.max() > top:
                return tag + ": randomize left the value range"
        elif op == "counters":
            pass
        elif op == "reset_counters":
            s.reset_counters()
            nacc = natt = 0
        elif op == "digest":
            d1 = s.digest()
            s.sweeps(sweep, 0)
            if s.digest() != d1:
                return tag + ": digest changed without a sweep"
        else:
            continue
        bad = check(tag)
        if bad:
            return bad
    s.close()
    return None


def run_case_ranks(c, dist, rank, world, local):
    """the same idea on several ranks (one process each): every call is made collectively, rank 0 compares the gathered
    slabs with the oracle after every step"""
    import numpy as np
    import torch
    import oracle_py as O
    from spparks_b200 import LatticeSweep
    keep = ("SPK_CUDA_LIB", "SPK_CUDA_ALLOW_EMULATOR", "SPK_CUDA_NCCL_LIB")
    for k in [k for k in os.environ if k.startswith("SPK_CUDA_") and k not in keep]:
        del os.environ[k]
    os.environ.update(TIER_ENV[c["tier"]])
    if c.get("halo") == "nccl":
        os.environ["SPK_CUDA_NO_P2P"] = "1"
    model, lat, q = c["model"], c["lat"], c["q"]
    n = list(c["n"])
    dim3 = lat.startswith("sc")
    per = max(2 if dim3 else 4, (n[2] if dim3 else n[1]) // 2 * 2)  # even slab thickness (2-D: >= the 4 ghost rows)
    if dim3:
        n[2] = per * world
    else:
        n[1] = per * world
    T, seed = c["T"], c["seed"]
    top = 2 if model.startswith("ising") else q
    o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
    colors = o.colors()
    want = np.random.default_rng(seed).integers(1, top + 1, size=o.nsites, dtype=np.int32)
    s = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=seed, device=local, rank=rank, nranks=world)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(LatticeSweep.comm_id()), dtype=torch.uint8)
    dist.broadcast(idt, 0)
    s.comm_init(idt.numpy().tobytes())
    lo, hi = rank * s.nlocal, (rank + 1) * s.nlocal
    s.upload(want[lo:hi])
    mask = None
    nacc = natt = 0
    sweep = 100
    siteseeds = None

    def gather(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return torch.cat(parts).numpy()

    def total(vals):
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t)
        return [float(x) for x in t]

    def check(what):
        got = gather(s.download())
        cnt = total(list(s.counters()))
        if rank != 0:
            return None
        if not np.array_equal(got, want):
            bad = np.flatnonzero(got != want)
            return "%s: spins differ at %d sites, first %d" % (what, bad.size, int(bad[0]))
        # (mask bytes are not compared here: a change unmasks its neighbours in the rank's OWN arrays, ghost copies
        # included, in the reference too, so the bytes next to a slab boundary depend on the decomposition; without
        # sectors every colour set is unmasked before it is swept, so the spins do not)
        if (int(cnt[0]), int(cnt[1])) != (nacc, natt):
            return "%s: counters %r, expected %r" % (what, (int(cnt[0]), int(cnt[1])), (nacc, natt))
        return None

    verdict = None
    for step, opd in enumerate(c["ops"]):
        op, a, b, u = opd["op"], opd["a"], opd["b"], opd["u"]
        tag = "step %d %s" % (step, op)
        bad = None
        if op == "sweeps":
            k = 1 + a % 3
            if mask is None:
                for i in range(k):
                    nacc += o.sweep_philox(want, seed, sweep + i)
            else:
                for i in range(k):
                    for col in range(1, o.ncolors + 1):
                        mask[colors == col] = 0
                        nacc += o.halfsweep_philox_mask(want, mask, seed, sweep + i, col, 0, None)
            natt += k * o.nsites
            s.sweeps(sweep, k)
            sweep += k
        elif op == "color":
            col, nloop = 1 + a % o.ncolors, 1 + b % 3
            s.halo_exchange()  # comm->all() before the set, as the plugin does (src/app_lattice.cpp:648-652)
            if mask is None:
                for l in range(nloop):
                    nacc += o.halfsweep_philox(want, seed, sweep, col, iloop=l)
                natt += nloop * int((colors == col).sum())
                s.sweep_color(sweep, col, None, nloop)
            else:
                mask[colors == col] = 0
                s.mask_clear_border(None, col)
                nacc += o.halfsweep_philox_mask(want, mask, seed, sweep, col, 0, None)
                natt += int((colors == col).sum())
                s.sweep_color(sweep, col, None, 1)
            sweep += 1
        elif op == "injected" and mask is None:
            col = 1 + a % o.ncolors
            kk = np.random.default_rng(a).integers(1, O.IM - 1, size=o.nsites, dtype=np.int32)
            u1, u2 = O.injected_uniforms(kk)
            nacc += o.halfsweep_injected(want, col, u1, u2)
            natt += int((colors == col).sum())
            s.halo_exchange()
            s.sweep_color_injected(col, u1[lo:hi], u2[lo:hi])
        elif op == "strict" and mask is None:
            if siteseeds is None:
                siteseeds = o.strict_siteseeds(O.mars(1 + a % 900000000))
                s.set_siteseeds(siteseeds[lo:hi])
            nacc += o.sweeps_strict(want, siteseeds, 1)
            natt += o.nsites
            for col in range(1, o.ncolors + 1):
                s.halo_exchange()
                s.sweep_color_strict(col, None, 1)
            if not np.array_equal(gather(s.get_siteseeds()), siteseeds) and rank == 0:
                bad = tag + ": site seeds differ"
        elif op == "upload":
            want = np.random.default_rng(a).integers(1, top + 1, size=o.nsites, dtype=np.int32)
            if model == "potts/pin" and u < 0.5:
                want[np.random.default_rng(b).random(o.nsites) < 0.1] = q + 1
            s.upload(want[lo:hi])
        elif op == "energy":
            e = total([s.energy()])[0]
            if rank == 0 and e != o.total_energy(want):
                bad = tag + ": energy differs"
        elif op == "site_energy":
            e = gather(s.site_energy())
            if rank == 0 and int(e.sum()) != int(o.total_energy(want)):
                bad = tag + ": per-site energies do not add up to the total"
        elif op == "clusters":
            got = s.cluster_stats()
            if rank == 0:
                oc = o.clusters(want)
                if got[0] != oc[0] or abs(got[1] - oc[1]) > 1e-9 * oc[1] or abs(got[2] - oc[2]) > 1e-9 * oc[2]:
                    bad = tag + ": cluster statistics %r, oracle %r" % (got, oc[:3])
        elif op == "temperature":
            T = [0.0, 0.3, 0.9, 2.0][a % 4]
            o = O.Oracle(model, lat, n[0], n[1], n[2], q, T)
            s.set_temperature(T)
        elif op == "seed":
            seed = 1 + (a * 7919 + b) % (2**40)
            s.set_seed(seed)
        elif op == "mask_on":
            mask = np.zeros(o.nsites, dtype=np.uint8)
            s.mask_enable(True)
        elif op == "mask_off":
            mask = None
            s.mask_enable(False)
        elif op == "set_range":
            park = 1 + a % (O.IM - 2)
            vlo = 1 + b % top
            vhi = vlo + (a % (top - vlo + 1))
            state, am = park, 1.0 / O.IM
            vals = np.empty(o.nsites, dtype=np.int32)
            for i in range(o.nsites):
                state = state * 16807 % O.IM
                vals[i] = vlo + min(int(state * am * (vhi - vlo + 1)), vhi - vlo)
            want = vals
            s.set_site_range(park, vlo, vhi)
        elif op == "pin" and model == "potts/pin":
            m = O.mars(1 + a % 900000000)
            park = int(O.lib().orc_mars_uniform(O.C.byref(m)) * O.IM) or 1
            frac = [0.02, 0.1, 0.3][b % 3]
            if int((want > q).sum()) <= int(frac * o.nsites) + o.nsites // 2:
                o.pin_create(want, frac, 0, 0, O.mars(1 + a % 900000000))
                s.pin_create(park, frac)
        elif op == "randomize":
            s.randomize(a, 1, top)
            want = gather(s.download()).copy()
        elif op == "reset_counters":
            s.reset_counters()
            nacc = natt = 0
        else:
            continue
        bad = bad or check(tag)
        flag = torch.tensor([1 if bad else 0], dtype=torch.int32)
        dist.all_reduce(flag)
        if int(flag.item()):
            verdict = bad or "failed on rank 0"
            break
    s.close()
    return verdict


def worker_ranks(cases):
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    for c in cases:
        try:
            r = run_case_ranks(c, dist, rank, world, local)
        except Exception as e:  # noqa: BLE001
            r = "exception on rank %d: %r" % (rank, e)
            print("FUZZ_NOTE " + r, flush=True)
        if rank == 0:
            print(("FUZZ_OK" if r is None else "FUZZ_FAIL " + json.dumps({"case": c, "what": r})), flush=True)
    dist.destroy_process_group()


def child(cases):
    for c in cases:
        try:
            r = run_case(c)
        except Exception as e:  # noqa: BLE001
            r = "exception: %r" % (e,)
        print(("FUZZ_OK" if r is None else "FUZZ_FAIL " + json.dumps({"case": c, "what": r})), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--ops", type=int, default=14)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--child", default=None)
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--port", type=int, default=29761)
    ap.add_argument("--worker-ranks", default=None)
    args = ap.parse_args()
    if args.child:
        child(json.loads(open(args.child).read()))
        return 0
    if args.worker_ranks:
        worker_ranks(json.loads(open(args.worker_ranks).read()))
        return 0
    import importlib.util
    import tempfile
    import numpy as np
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "host_shim", "build_emu.py"))
    be = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(be)
    lib = os.environ.get("SPK_EMU_USE_LIB") or be.build()
    env = dict(os.environ, SPK_CUDA_LIB=lib, SPK_CUDA_ALLOW_EMULATOR="1")
    rng = np.random.default_rng(args.seed)
    cases = [draw_case(rng, args.ops) for _ in range(args.cases)]
    if args.ranks > 1:
        for c in cases:
            c["halo"] = str(rng.choice(["p2p", "p2p", "nccl"]))
        arena = tempfile.mkdtemp(prefix="spk_emu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            cfile = os.path.join(arena, "cases.json")
            json.dump(cases, open(cfile, "w"))
            env.update(SPK_CUDA_NCCL_LIB=be.NCCL_LIB, SPK_EMU_ARENA_DIR=arena, SPK_EMU_WORKERS="2", OMP_NUM_THREADS="1")
            r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.ranks), "--master-addr",
                                "127.0.0.1", "--master-port", str(args.port), os.path.abspath(__file__), "--worker-ranks", cfile],
                               env=env, cwd=ROOT, capture_output=True, text=True)
        finally:
            import shutil
            shutil.rmtree(arena, ignore_errors=True)
        lines = [l for l in r.stdout.splitlines() if l.startswith("FUZZ_OK") or l.startswith("FUZZ_FAIL")]
        fails = [json.loads(l[10:]) for l in lines if l.startswith("FUZZ_FAIL")]
        print("%d ranks: %d / %d sequences ok" % (args.ranks, len(lines) - len(fails), len(cases)))
        for f in fails[:10]:
            print("FAIL", f["what"], "|", f["case"]["model"], f["case"]["lat"], f["case"]["n"], "q", f["case"]["q"], f["case"]["tier"], f["case"]["halo"],
                  "|", " ".join(o["op"] for o in f["case"]["ops"]))
        for l in r.stdout.splitlines():
            if l.startswith("FUZZ_NOTE"):
                print(l)
        if len(lines) != len(cases):
            print("the run ended early:\n" + r.stdout[-1000:] + r.stderr[-3000:])
            return 1
        return 1 if fails else 0
    fails, done = [], 0
    with tempfile.TemporaryDirectory() as tmp:
        for bnum in range(0, len(cases), args.batch):
            batch = cases[bnum:bnum + args.batch]
            cfile = os.path.join(tmp, "batch.json")
            json.dump(batch, open(cfile, "w"))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", cfile], env=env, cwd=ROOT, capture_output=True, text=True)
            lines = [l for l in r.stdout.splitlines() if l.startswith("FUZZ_")]
            fails += [json.loads(l[10:]) for l in lines if l.startswith("FUZZ_FAIL")]
            done += len(lines)
            if len(lines) != len(batch):
                fails.append({"case": batch[len(lines)], "what": "process died: " + r.stderr[-600:]})
                done = bnum + len(batch)
            print("%d / %d sequences, %d failures" % (done, len(cases), len(fails)), flush=True)
    for f in fails[:10]:
        ops = [o["op"] for o in f["case"]["ops"]]
        print("FAIL", f["what"], "|", f["case"]["model"], f["case"]["lat"], f["case"]["n"], "q", f["case"]["q"], f["case"]["tier"], "|", " ".join(ops))
    if fails:
        json.dump(fails, open("/tmp/emu_fuzz_api_fails.json", "w"))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
