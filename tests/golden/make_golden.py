#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/ref_driver).

Run in the build container (needs /root/reference to have been compiled by
`make -C oracle/ref`):   python tests/golden/make_golden.py
The fixtures are committed; the GPU box and CI never run this script.

Each fixture pins one aspect of the hot path against the reference's own objects:
  geom_<lattice>.npz      neighbour table, colours, sector/colour sets (create_sites / create_set)
  half_<model>_<lattice>_T<t>.npz   injected-uniform colour half-sweeps, one after another
  sets_<model>_<lattice>.npz        one iterate_rejection pass over sector->colour sets, injected
  strict_<model>_<lattice>_T<t>.npz `sweep color/strict` multi-sweep run (per-site Park-Miller)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
IM = 2147483647

DIM = {"sq/4n": 2, "sq/8n": 2, "tri": 2, "sc/6n": 3, "sc/26n": 3}
NBASIS = {"sq/4n": 1, "sq/8n": 1, "tri": 2, "sc/6n": 1, "sc/26n": 1}
NCOLORS = {"sq/4n": 2, "sq/8n": 4, "tri": 4, "sc/6n": 2, "sc/26n": 8}
DT_SWEEP = {"ising": lambda q, mn: 0.5, "potts": lambda q, mn: 1.0 / q,
            "potts/pin": lambda q, mn: 1.0 / q, "potts/neighonly": lambda q, mn: 1.0 / mn,
            # src/app_ising_single.cpp:29, src/app_potts_neigh.cpp:47
            "ising/single": lambda q, mn: 1.0, "potts/neigh": lambda q, mn: 1.0 / mn}
ISING_LIKE = ("ising", "ising/single")
MAXNEIGH = {"sq/4n": 4, "sq/8n": 8, "tri": 6, "sc/6n": 6, "sc/26n": 26}


def tag(s):
    return s.replace("/", "_")


def header(model, q, lattice, n, seed=12345):
    nx, ny, nz = n
    app = model if model in ISING_LIKE else "%s %d" % (model, q)
    zreg = "0 %d" % nz if DIM[lattice] == 3 else "-0.5 0.5"
    return ["seed %d" % seed, "app_style %s" % app, "dimension %d" % DIM[lattice],
            "lattice %s 1.0" % lattice, "region box block 0 %d 0 %d %s" % (nx, ny, zreg),
            "create_box box", "create_sites box", "set site range 1 %d" % (2 if model in ISING_LIKE else q)]


def run_driver(lines, tmp):
    path = os.path.join(tmp, "drv.in")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    out = subprocess.run([DRIVER, path, "-quiet"], cwd=tmp, capture_output=True, text=True)
    if out.returncode != 0 or "ERROR" in out.stdout:
        raise RuntimeError("ref_driver failed:\n" + out.stdout + out.stderr)
    return out.stdout


def ints(tmp, name):
    return np.fromfile(os.path.join(tmp, name), dtype=np.int32)


def nsites(lattice, n):
    return n[0] * n[1] * (n[2] if DIM[lattice] == 3 else 1) * NBASIS[lattice]


def gen_geom(lattice, n):
    with tempfile.TemporaryDirectory() as tmp:
        lines = header("potts", 7, lattice, n) + ["sweep color", "sector yes", "temperature 0.5", "run 0.0",
                                                  "@neigh neigh.bin", "@colors colors.bin", "@sets sets.bin"]
        run_driver(lines, tmp)
        ng = ints(tmp, "neigh.bin")
        nlocal, maxneigh = int(ng[0]), int(ng[1])
        neigh = ng[2:].reshape(nlocal, maxneigh)
        colors = ints(tmp, "colors.bin")
        sets = ints(tmp, "sets.bin")
    np.savez_compressed(os.path.join(HERE, "geom_%s.npz" % tag(lattice)), n=np.array(n, dtype=np.int32),
                        neigh=neigh, colors=colors, sets=sets)


def random_spins(rng, model, q, ns, pinfrac=0.0, blocky=False):
    hi = 2 if model in ISING_LIKE else q
    s = rng.integers(1, hi + 1, size=ns, dtype=np.int32)
    if blocky:  # long runs of equal spins so that neighonly meets few/zero candidates
        s = np.repeat(rng.integers(1, hi + 1, size=(ns + 15) // 16, dtype=np.int32), 16)[:ns].copy()
    if pinfrac > 0:
        s[rng.random(ns) < pinfrac] = q + 1
    return s


def gen_half(model, q, lattice, n, T, rng, blocky=False):
    ns = nsites(lattice, n)
    nc = NCOLORS[lattice]
    with tempfile.TemporaryDirectory() as tmp:
        spins0 = random_spins(rng, model, q, ns, 0.08 if model == "potts/pin" else 0.0, blocky)
        spins0.tofile(os.path.join(tmp, "s0.bin"))
        ks = rng.integers(1, IM - 1, size=(nc, ns), dtype=np.int32)
        # force a few extreme uniforms (first/last admissible Park-Miller states)
        ks[:, 0] = 1
        ks[:, -1] = IM - 1
        lines = header(model, q, lattice, n) + ["sweep color", "sector no", "temperature %r" % T, "run 0.0",
                                                "@load s0.bin"]
        for c in range(nc):
            ks[c].tofile(os.path.join(tmp, "k%d.bin" % c))
            lines += ["@halfsweep %d k%d.bin" % (c + 1, c), "@save s%d.bin" % (c + 1)]
        lines += ["@energy"]
        out = run_driver(lines, tmp)
        after = np.stack([ints(tmp, "s%d.bin" % (c + 1)) for c in range(nc)])
        energy = float([l for l in out.splitlines() if l.startswith("energy")][-1].split()[1])
    name = "half_%s_%s_T%s%s.npz" % (tag(model), tag(lattice), str(T).replace(".", "p"), "_blocky" if blocky else "")
    np.savez_compressed(os.path.join(HERE, name), n=np.array(n, dtype=np.int32), q=q, T=T, spins0=spins0, k=ks,
                        after=after, energy=energy)


def gen_sets(model, q, lattice, n, T, rng):
    """sector yes + sweep color: one pass over the sets in iterate_rejection order."""
    ns = nsites(lattice, n)
    with tempfile.TemporaryDirectory() as tmp:
        spins0 = random_spins(rng, model, q, ns)
        spins0.tofile(os.path.join(tmp, "s0.bin"))
        k = rng.integers(1, IM - 1, size=ns, dtype=np.int32)
        k.tofile(os.path.join(tmp, "k.bin"))
        lines = header(model, q, lattice, n) + ["sweep color", "sector yes", "temperature %r" % T, "run 0.0",
                                                "@load s0.bin", "@setsweep k.bin", "@save s1.bin"]
        run_driver(lines, tmp)
        after = ints(tmp, "s1.bin")
    np.savez_compressed(os.path.join(HERE, "sets_%s_%s.npz" % (tag(model), tag(lattice))),
                        n=np.array(n, dtype=np.int32), q=q, T=T, spins0=spins0, k=k, after=after)


def gen_masksets(model, q, lattice, n, T, npass, sector, rng, blocky=True):
    """`sweep color mask yes` + sectors: npass passes over the sets in iterate_rejection order with the
    reference's boundary_clear_mask / mask rule, injected uniforms; spins and mask bytes after every pass."""
    ns = nsites(lattice, n)
    with tempfile.TemporaryDirectory() as tmp:
        spins0 = random_spins(rng, model, q, ns, 0.05 if model == "potts/pin" else 0.0, blocky)
        spins0.tofile(os.path.join(tmp, "s0.bin"))
        ks = rng.integers(1, IM - 1, size=(npass, ns), dtype=np.int32)
        lines = header(model, q, lattice, n) + ["sweep color mask yes", "sector %s" % sector, "temperature %r" % T,
                                                "run 0.0", "@load s0.bin"]
        for p in range(npass):
            ks[p].tofile(os.path.join(tmp, "k%d.bin" % p))
            lines += ["@masksweep k%d.bin m%d.bin" % (p, p), "@save s%d.bin" % (p + 1)]
        run_driver(lines, tmp)
        after = np.stack([ints(tmp, "s%d.bin" % (p + 1)) for p in range(npass)])
        masks = np.stack([ints(tmp, "m%d.bin" % p) for p in range(npass)]).astype(np.uint8)
    nsec = {"yes": 4 if DIM[lattice] == 2 else 8}.get(sector, None) or int(sector)
    name = "masksets_%s_%s_s%d.npz" % (tag(model), tag(lattice), nsec)
    np.savez_compressed(os.path.join(HERE, name), n=np.array(n, dtype=np.int32), q=q, T=T, nsector=nsec,
                        spins0=spins0, k=ks, after=after, mask=masks)


def main_mask():
    """SURVEY §8f rank 3 tail: `sweep ... mask yes`:  python tests/golden/make_golden.py mask"""
    rng = np.random.default_rng(20261019)
    gen_masksets("potts", 4, "sc/26n", (8, 8, 8), 0.0, 4, "yes", rng)
    gen_masksets("potts/neighonly", 5, "sc/26n", (8, 6, 8), 0.0, 4, "yes", rng)
    gen_masksets("potts/pin", 4, "sc/26n", (8, 8, 6), 0.0, 3, "2", rng)
    gen_masksets("ising", 2, "sq/4n", (16, 12, 1), 0.0, 4, "yes", rng)
    gen_masksets("ising", 2, "sq/4n", (12, 12, 1), 0.8, 3, "2", rng, blocky=False)
    gen_masksets("potts", 6, "tri", (8, 10, 1), 0.0, 4, "yes", rng, blocky=False)
    gen_masksets("potts/neigh", 3, "sq/8n", (12, 10, 1), 0.0, 3, "yes", rng)
    gen_masksets("ising/single", 2, "sc/6n", (8, 6, 6), 0.0, 3, "4", rng)


def gen_strict(model, q, lattice, n, T, nsweeps, rng):
    ns = nsites(lattice, n)
    dt = DT_SWEEP[model](q, MAXNEIGH[lattice])
    with tempfile.TemporaryDirectory() as tmp:
        spins0 = random_spins(rng, model, q, ns, 0.05 if model == "potts/pin" else 0.0)
        spins0.tofile(os.path.join(tmp, "s0.bin"))
        lines = header(model, q, lattice, n, seed=777) + [
            "sweep color/strict", "sector no", "temperature %r" % T, "run 0.0", "@load s0.bin",
            "@siteseeds seeds0.bin", "run %.17g" % (dt * (nsweeps - 0.5)), "@save s1.bin",
            "@siteseeds seeds1.bin", "@counters", "@energy"]
        out = run_driver(lines, tmp)
        seeds0, seeds1, after = ints(tmp, "seeds0.bin"), ints(tmp, "seeds1.bin"), ints(tmp, "s1.bin")
        cnt = [l for l in out.splitlines() if l.startswith("counters")][-1].split()
        energy = float([l for l in out.splitlines() if l.startswith("energy")][-1].split()[1])
        assert int(cnt[3]) == nsweeps, cnt
    name = "strict_%s_%s_T%s.npz" % (tag(model), tag(lattice), str(T).replace(".", "p"))
    np.savez_compressed(os.path.join(HERE, name), n=np.array(n, dtype=np.int32), q=q, T=T, nsweeps=nsweeps,
                        spins0=spins0, seeds0=seeds0, seeds1=seeds1, after=after, naccept=int(cnt[1]),
                        nattempt=int(cnt[2]), energy=energy, master_seed=777)


def main():
    if not os.path.exists(DRIVER):
        sys.exit("build the reference first: make -C oracle/ref")
    rng = np.random.default_rng(20261017)
    sizes = {"sq/4n": (8, 6, 1), "sq/8n": (6, 8, 1), "tri": (6, 4, 1), "sc/6n": (4, 6, 4), "sc/26n": (6, 4, 4)}
    for lat, n in sizes.items():
        gen_geom(lat, n)
    big = {"sq/4n": (16, 12, 1), "sq/8n": (12, 10, 1), "tri": (8, 10, 1), "sc/6n": (8, 6, 6), "sc/26n": (8, 6, 10)}
    for T in (0.0, 1.0):
        gen_half("ising", 2, "sq/4n", big["sq/4n"], T, rng)
        for lat in big:
            gen_half("potts", 5, lat, big[lat], T, rng)
        gen_half("potts/neighonly", 6, "sc/26n", big["sc/26n"], T, rng)
        gen_half("potts/neighonly", 4, "tri", big["tri"], T, rng)
        gen_half("potts/pin", 5, "sc/26n", big["sc/26n"], T, rng)
    gen_half("ising", 2, "sc/6n", big["sc/6n"], 2.2, rng)
    gen_half("ising", 2, "sq/8n", big["sq/8n"], 3.0, rng)
    gen_half("potts/neighonly", 10000, "sc/26n", big["sc/26n"], 0.5, rng)
    gen_half("potts/neighonly", 3, "sc/26n", big["sc/26n"], 0.7, rng, blocky=True)
    gen_half("potts/neighonly", 3, "sq/4n", big["sq/4n"], 0.0, rng, blocky=True)
    gen_half("potts/pin", 50, "sq/8n", big["sq/8n"], 0.9, rng)
    gen_sets("potts", 5, "sc/26n", (8, 6, 10), 0.8, rng)
    gen_sets("ising", 2, "sq/4n", (10, 14, 1), 1.5, rng)
    gen_sets("potts", 4, "tri", (6, 10, 1), 0.6, rng)
    gen_strict("ising", 2, "sq/4n", (16, 12, 1), 1.1346, 5, rng)
    gen_strict("potts", 7, "sc/26n", (6, 6, 8), 0.9, 4, rng)
    gen_strict("potts/neighonly", 9, "sc/26n", (6, 6, 8), 0.5, 4, rng)
    gen_strict("potts/pin", 6, "sc/26n", (6, 6, 8), 1.2, 4, rng)
    gen_strict("potts", 12, "tri", (8, 8, 1), 0.0, 6, rng)
    print("wrote", len([f for f in os.listdir(HERE) if f.endswith(".npz")]), "fixtures")


def main_variants():
    """ising/single and potts/neigh (own random stream, so that the fixtures of
    main() stay byte-identical):  python tests/golden/make_golden.py variants"""
    rng = np.random.default_rng(20261018)
    for T in (0.0, 1.0):
        gen_half("ising/single", 2, "sq/4n", (16, 12, 1), T, rng)
        gen_half("potts/neigh", 6, "sc/26n", (8, 6, 10), T, rng)
    gen_half("ising/single", 2, "sc/26n", (8, 6, 10), 6.0, rng)
    gen_half("potts/neigh", 4, "tri", (8, 10, 1), 0.6, rng)
    gen_half("potts/neigh", 3, "sc/26n", (8, 6, 10), 0.7, rng, blocky=True)
    gen_half("potts/neigh", 3, "sq/8n", (12, 10, 1), 0.5, rng)
    gen_sets("potts/neigh", 5, "sc/26n", (8, 6, 10), 0.8, rng)
    gen_strict("ising/single", 2, "sq/4n", (16, 12, 1), 1.1346, 5, rng)
    gen_strict("potts/neigh", 9, "sc/26n", (6, 6, 8), 0.5, 4, rng)
    gen_strict("potts/neigh", 5, "tri", (8, 8, 1), 0.4, 6, rng)


def gen_pins(lattice, n, q, fraction, multi, nthresh, seed, rng, blocky):
    """`pin fraction multi nthresh` of app_style potts/pin (src/app_potts_pin.cpp:216-349) run by the
    reference itself: spins before and after.  The lattice is loaded (blocky = grains, so that nthresh > 0
    finds boundaries), the master stream has served `set site range` once before the pin command."""
    ns = nsites(lattice, n)
    s0 = random_spins(rng, "potts/pin", q, ns, blocky=blocky)
    with tempfile.TemporaryDirectory() as tmp:
        s0.tofile(os.path.join(tmp, "s0.bin"))
        lines = header("potts/pin", q, lattice, n, seed=seed) + ["@load s0.bin", "pin %r %d %d" % (fraction, multi, nthresh),
                                                                 "@save s1.bin"]
        run_driver(lines, tmp)
        s1 = ints(tmp, "s1.bin")
    name = "pins_%s_m%d_t%d.npz" % (tag(lattice), multi, nthresh)
    np.savez_compressed(os.path.join(HERE, name), s0=s0, s1=s1, n=np.array(n), q=q, fraction=fraction, multi=multi,
                        nthresh=nthresh, seed=seed, lattice=lattice)
    print(name, "pinned", int((s1 > q).sum()), "of", ns)


def main_pins():
    rng = np.random.default_rng(20261017)
    gen_pins("sc/26n", (12, 10, 8), 50, 0.05, 0, 0, 4711, rng, False)
    gen_pins("sq/8n", (40, 30, 1), 9, 0.3, 0, 0, 99, rng, False)      # many rounds: most picks hit pinned sites
    gen_pins("tri", (16, 12, 1), 20, 0.1, 0, 0, 31337, rng, False)
    gen_pins("sc/26n", (12, 10, 8), 6, 0.04, 0, 10, 4711, rng, True)   # only at grain boundaries
    gen_pins("sq/4n", (30, 30, 1), 5, 0.1, 1, 0, 777, rng, False)      # site + its neighbours
    gen_pins("sc/6n", (10, 10, 10), 5, 0.15, 1, 2, 555, rng, True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        main_variants()
    elif len(sys.argv) > 1 and sys.argv[1] == "pins":
        main_pins()
    elif len(sys.argv) > 1 and sys.argv[1] == "mask":
        main_mask()
    else:
        main()
