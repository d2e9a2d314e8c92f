"""Randomised multi-rank test of the slab decomposition against the oracle, on the CPU (emulated build).

    python tools/emu_fuzz_ranks.py [--ranks N] [--cases C] [--seed S]

One process per rank (torchrun, gloo).  Every case draws an app, a lattice, extents whose slab
dimension splits into N even slabs, Q, T, a kernel tier, the halo mode (peer-memory stores or NCCL
exchange) and a call pattern of spk_cuda_sweeps interleaved with energy / cluster statistics /
counters; rank 0 gathers the slabs and compares spins, counters, energy and cluster statistics with
oracle/ on the whole lattice.  Test infrastructure only (tests/host_shim/README.md): CUDA IPC is a
shared file mapping, NCCL a shared-memory stand-in, the kernels run on the CUDA emulator.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

TIERS = {
    "tuned": {},
    "chain": {"SPK_CUDA_NEIGH_MODE": "chain"},
    "adaptive": {"SPK_CUDA_NEIGH_MODE": "adaptive"},
    "tuned32": {"SPK_CUDA_NO_PK16": "1"},
    "strip": {"SPK_CUDA_NO_STREAM": "1", "SPK_CUDA_NO_FUSED": "1"},
    "generic": {"SPK_CUDA_FORCE_GENERIC": "1"},
    "int32stream": {"SPK_CUDA_NO_SHADOW16": "1"},
}
LATS = ["sq/4n", "sq/8n", "tri", "sc/6n", "sc/26n"]
MODELS = ["ising", "ising/single", "potts", "potts/pin", "potts/neighonly", "potts/neigh"]


def draw_case(rng, world):
    lat = LATS[int(rng.integers(len(LATS)))]
    model = MODELS[int(rng.integers(len(MODELS)))]
    dim3 = lat.startswith("sc")
    nx = int(rng.choice([4, 8, 10, 16, 24, 32, 40, 64, 72, 128, 132, 136, 256, 264]))
    if dim3:
        ny = int(rng.choice([2, 4, 6, 10, 16, 18, 32, 34, 36]))
        nz = 2 * world * int(rng.integers(1, 5))            # even slabs of 2 .. 8 planes
    else:
        ny = 2 * world * int(rng.integers(2, 9))            # even slabs of 4 .. 16 rows (>= the 4 ghost rows)
        nz = 1
    q = 2 if model.startswith("ising") else int(rng.choice([2, 3, 5, 50, 10000, 65533, 65534, 70000]))
    tier = "tuned" if rng.random() < 0.5 else list(TIERS)[int(rng.integers(len(TIERS)))]
    halo = str(rng.choice(["p2p", "p2p", "nccl"]))
    if rng.random() < 0.15:  # the 16-bit shadow on several ranks (neighbours store into each other's shadow ghost planes)
        lat, model, tier, halo = "sc/26n", str(rng.choice(["potts", "potts/pin"])), "tuned", "p2p"
        nx, ny, nz = int(rng.choice([8, 16, 24, 264])), int(rng.choice([2, 6, 16, 32, 34])), 2 * world * int(rng.integers(1, 5))
        q = int(rng.choice([3, 50, 10000, 65533]))
    return {"lat": lat, "model": model, "n": [nx, ny, nz], "q": q, "T": float(rng.choice([0.0, 0.4, 1.0, 5.0])),
            "tier": tier, "halo": halo, "seed": int(rng.integers(1, 2**40)),
            "calls": [int(x) for x in rng.integers(1, 4, size=int(rng.integers(1, 4)))],   # sweeps per spk_cuda_sweeps call
            "between": [str(x) for x in rng.choice(["none", "energy", "clusters", "download"], size=3)],
            "start": str(rng.choice(["random", "few", "grains"])),
            # potts/pin: half of the cases get their pins from the device-side `pin f 0 0` (collective) instead of the upload
            "pin": float(rng.choice([0.0, 0.0, 0.02, 0.3])),
            "chunks": int(rng.choice([0, 0, 1, 2])), "ctas": int(rng.choice([0, 0, 1, 3]))}


def worker(cases):
    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle_py as O
    from spparks_b200 import LatticeSweep
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    keep = ("SPK_CUDA_LIB", "SPK_CUDA_ALLOW_EMULATOR", "SPK_CUDA_NCCL_LIB")
    for c in cases:
        for k in [k for k in os.environ if k.startswith("SPK_CUDA_") and k not in keep]:
            del os.environ[k]
        os.environ.update(TIERS[c["tier"]])
        if c["halo"] == "nccl":
            os.environ["SPK_CUDA_NO_P2P"] = "1"
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
        else:
            idx = np.arange(o.nsites) // o.nbasis
            i, j, k = idx % n[0], (idx // n[0]) % n[1], idx // (n[0] * n[1])
            spins = (1 + ((i // 5) * 7 + (j // 3) * 13 + (k // 2) * 29) % top).astype(np.int32)
        device_pin = model == "potts/pin" and c.get("pin", 0.0) > 0
        if model == "potts/pin" and not device_pin:
            spins[rng.random(o.nsites) < 0.07] = q + 1
        what = None
        try:
            s = LatticeSweep(model, lat, n[0], n[1], n[2], nspins=q, temperature=T, seed=c["seed"], device=local, rank=rank, nranks=world)
            idt = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                idt = torch.frombuffer(bytearray(LatticeSweep.comm_id()), dtype=torch.uint8)
            dist.broadcast(idt, 0)
            s.comm_init(idt.numpy().tobytes())
            s.upload(spins[rank * s.nlocal:(rank + 1) * s.nlocal])
            want = spins.copy()
            if device_pin:
                m = O.mars(c["seed"] % 900000000 + 1)
                park = int(O.lib().orc_mars_uniform(O.C.byref(m)) * O.IM) or 1
                npin = s.pin_create(park, c["pin"])
                onpin = o.pin_create(want, c["pin"], 0, 0, O.mars(c["seed"] % 900000000 + 1))
                if npin != onpin:
                    what = "pin_create: %d sites pinned, oracle %d" % (npin, onpin)
            nacc, sweep = 0, 7
            energies, clusters = [], []
            for ci, ns in enumerate(c["calls"]):
                s.sweeps(sweep, ns)
                for sw in range(ns):
                    nacc += o.sweep_philox(want, c["seed"], sweep + sw) if rank == 0 else 0
                sweep += ns
                b = c["between"][ci]
                if b == "energy":
                    energies.append((s.energy(), o.total_energy(want) if rank == 0 else 0))
                elif b == "clusters":
                    clusters.append((s.cluster_stats(), o.clusters(want)[:3] if rank == 0 else None))
                elif b == "download":
                    s.download()
            mine = torch.from_numpy(s.download())
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            acc = torch.tensor(list(s.counters()) + [s.energy()], dtype=torch.float64)
            dist.all_reduce(acc)
            clus = s.cluster_stats()
            path, mode = s.sweep_path()[0], s.halo_mode()
            s.close()
            if rank == 0:
                got = torch.cat(parts).numpy()
                if what is not None:
                    pass
                elif not np.array_equal(got, want):
                    bad = np.flatnonzero(got != want)
                    what = "spins differ at %d sites, first %d" % (bad.size, int(bad[0]))
                elif (int(acc[0]), int(acc[1])) != (nacc, sum(c["calls"]) * o.nsites):
                    what = "counters %r, oracle %r" % ((int(acc[0]), int(acc[1])), (nacc, sum(c["calls"]) * o.nsites))
                elif float(acc[2]) != o.total_energy(want):
                    what = "energy differs"
                else:
                    # energy() returns the slab's share: summed over ranks above; the in-between values are per rank
                    oc = o.clusters(want)
                    if clus[0] != oc[0] or abs(clus[1] - oc[1]) > 1e-9 * oc[1] or abs(clus[2] - oc[2]) > 1e-9 * oc[2]:
                        what = "cluster statistics %r, oracle %r" % (clus, oc[:3])
                    for cs, ocs in clusters:
                        if cs[0] != ocs[0] or abs(cs[2] - ocs[2]) > 1e-9 * ocs[2]:
                            what = "cluster statistics between sweeps %r, oracle %r" % (cs, ocs)
        except Exception as e:  # noqa: BLE001
            what = "exception on rank %d: %r" % (rank, e)
            path, mode = "?", -1
        flag = torch.tensor([0 if what is None else 1], dtype=torch.int32)
        dist.all_reduce(flag)
        if rank == 0:
            if int(flag.item()) == 0:
                print("FUZZ_OK " + json.dumps({"path": path, "halo_mode": mode}), flush=True)
            else:
                print("FUZZ_FAIL " + json.dumps({"case": c, "what": what or "failed on another rank"}), flush=True)
        elif what is not None:
            print("FUZZ_NOTE rank %d: %s" % (rank, what), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=2)
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--port", type=int, default=29771)
    ap.add_argument("--worker", default=None)
    args = ap.parse_args()
    if args.worker:
        worker(json.load(open(args.worker)))
        return 0
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "host_shim", "build_emu.py"))
    be = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(be)
    lib = os.environ.get("SPK_EMU_USE_LIB") or be.build()  # (the override: tools/emu_mutants.py)
    rng = np.random.default_rng(args.seed)
    cases = [draw_case(rng, args.ranks) for _ in range(args.cases)]
    arena = tempfile.mkdtemp(prefix="spk_emu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        cfile = os.path.join(arena, "cases.json")
        json.dump(cases, open(cfile, "w"))
        env = dict(os.environ, SPK_CUDA_LIB=lib, SPK_CUDA_ALLOW_EMULATOR="1", SPK_CUDA_NCCL_LIB=be.NCCL_LIB, SPK_EMU_ARENA_DIR=arena,
                   SPK_EMU_WORKERS="2", OMP_NUM_THREADS="1")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.ranks), "--master-addr",
               "127.0.0.1", "--master-port", str(args.port), os.path.abspath(__file__), "--worker", cfile]
        r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True)
    finally:
        shutil.rmtree(arena, ignore_errors=True)
    ok = [l for l in r.stdout.splitlines() if l.startswith("FUZZ_OK")]
    fails = [l for l in r.stdout.splitlines() if l.startswith("FUZZ_FAIL") or l.startswith("FUZZ_NOTE")]
    paths = {}
    for l in ok:
        d = json.loads(l[8:])
        key = "%s halo_mode=%d" % (d["path"], d["halo_mode"])
        paths[key] = paths.get(key, 0) + 1
    print("%d ranks: %d / %d cases ok" % (args.ranks, len(ok), len(cases)))
    print("swept with:", json.dumps(paths, sort_keys=True))
    for l in fails:
        print(l)
    if len(ok) + len([l for l in fails if l.startswith("FUZZ_FAIL")]) != len(cases):
        print("the run ended early:\n" + r.stdout[-1500:] + r.stderr[-3000:])
        return 1
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
