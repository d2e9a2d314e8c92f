"""Mutation check of the parity tests, on the CPU (emulated build of the library).

    python tools/emu_mutants.py [first [last]]

Each mutant is the library's source with ONE small defect planted in a kernel or in the host logic
(a wrong halo column, a ring slot off by one, a dropped barrier, a wrong compare cap, ...).  It is
built for the host (tests/host_shim/build_emu.py with SPK_EMU_CSRC / SPK_EMU_BUILD_DIR, the tree is
not touched) and tests/test_zz_more_parity.py + tests/test_gpu_parity.py run against it (a survivor once more
under shuffled threads and lazy copies): a mutant is KILLED when a test fails.
A survivor of the one-rank files also meets 60 random two-rank cases (tools/emu_fuzz_ranks.py).
A survivor is a hole in the tests (or an equivalent mutant) and is printed as such.  This says how
much a green run of the emulated tier is worth; it is test infrastructure, not product.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "spparks_b200", "csrc")

# (file, text to replace (must occur exactly once), replacement, what breaks)
MUTANTS = [
    ("sweep_stream.cuh", "pp.col_l = x0 == 0 ? 3 : -1;", "pp.col_l = x0 == 0 ? 2 : -1;", "streaming kernel: left periodic column patched one element off"),
    ("sweep_stream.cuh", "pp.gx_l = rowlen - 1;", "pp.gx_l = rowlen - 2;", "streaming kernel: left periodic image taken from the wrong lattice column"),
    ("sweep_stream.cuh", "const int rb = Y - y0 + 1;\n  pp.row_b = rb <= C::RP - 1 ? rb : -1;\n  return pp;", "const int rb = Y - y0 + 1;\n  pp.row_b = rb <= C::RP - 2 ? rb : -1;\n  return pp;", "streaming kernel: bottom periodic row not patched when it is the tile's last halo row"),
    ("sweep_stream.cuh", "if (s + 1 < nsteps) slot = adv(slot, C::KSTEP);", "if (s + 1 < nsteps) slot = adv(slot, 1);", "streaming kernel: ring advanced by one plane instead of KSTEP"),
    ("sweep_stream.cuh", "      __syncthreads();  // planes k-1 (and k) may now be overwritten by the next step's loads", "      // (barrier dropped)", "streaming kernel: barrier before the ring slots are reused dropped"),
    ("sweep_stream.cuh", "const unsigned long long leader = gsite0 + P - (upper ? 4 : 0);", "const unsigned long long leader = gsite0 + P - (upper ? 2 : 0);", "streaming kernel: Philox block leader of the upper ising half-block wrong"),
    ("sweep_stream16.cuh", "pp.col_l = x0 == 0 ? 6 : -1;  // word (x0-2, x0-1)", "pp.col_l = x0 == 0 ? 4 : -1;  // word (x0-2, x0-1)", "16-bit shadow kernel: left periodic word patched one word off"),
    ("sweep_stream16.cuh", "rr == 4 ? (P ? CAP_LO : CAP_HI) : CAP_BOTH);", "rr == 4 ? (P ? CAP_HI : CAP_LO) : CAP_BOTH);", "16-bit shadow kernel: own-row cap counts the site itself instead of its neighbour"),
    ("sweep_stream16.cuh", "if (e0 == 26) return trial;", "if (e0 >= 25) return trial;", "16-bit shadow kernel: accepts without the second count one energy too early"),
    ("sweep_bits2d.cuh", "const uint32_t shifted = px ? __funnelshift_r(same, adj, 1) : __funnelshift_l(adj, same, 1);", "const uint32_t shifted = px ? __funnelshift_l(adj, same, 1) : __funnelshift_r(same, adj, 1);", "bit-shadow kernel: horizontal neighbour shifted the wrong way"),
    ("sweep_bits2d.cuh", "const uint32_t ge2 = c12 | c34 | cc;", "const uint32_t ge2 = c12 | c34;", "bit-shadow kernel: carry of the bit-sliced adder dropped"),
    ("sweep_bits2d.cuh", "const unsigned long long leader = rowbase + 8ull * (unsigned long long)(8 * k + q);", "const unsigned long long leader = rowbase + 8ull * (unsigned long long)(8 * k + q) + 2ull;", "bit-shadow kernel: Philox counter of a block off by one site"),
    ("sweep_tile2d.cuh", "mbar_wait(bar, phase);\n    phase ^= 1;", "mbar_wait(bar, phase);", "colour-fused 2-D kernel: barrier phase not toggled (second tile waits on the completed phase)"),
    ("event.cuh", "pick = (int)__umulhi((uint32_t)(NULLBIN ? NN : nevent), d.w1);", "pick = (int)__umulhi((uint32_t)(NULLBIN ? NN : nevent + 1), d.w1);", "candidate pick scaled by nevent + 1"),
    ("event.cuh", "constexpr uint32_t NB26_VALID = 0x1FFF1FFFu;", "constexpr uint32_t NB26_VALID = 0x1FFF0FFFu;", "candidate elimination: one neighbour slot dropped from the offset table"),
    ("event.cuh", "if (e1 <= e0) return trial;\n  if (p.t_zero) return old;\n  const int dE = e1 - e0;", "if (e1 < e0) return trial;\n  if (p.t_zero) return old;\n  const int dE = e1 - e0;", "events: dE = 0 goes through the acceptance draw instead of being accepted"),
    ("aux_kernels.cuh", "border |= ghost || ii < b.ilo || ii >= b.ihi || jj < b.jlo || jj >= b.jhi || kk < b.klo || kk >= b.khi;", "border |= ghost || ii < b.ilo || ii >= b.ihi || jj < b.jlo || jj >= b.jhi || kk < b.klo || kk > b.khi;", "mask border: upper z face of a sector not recognised"),
    ("cluster.cuh", "const uint32_t old = atomicMin(label + a, b);", "const uint32_t old = atomicMin(label + b, a);", "clusters: union hangs the smaller root under the larger"),
    ("model_entry.cuh", "    h->nattempt += class_count<SPK_SC_26N>(h, b, c0);\n    if (nq <= 0) return 0;", "    h->nattempt += class_count<SPK_SC_26N>(h, b, c0) + 1;\n    if (nq <= 0) return 0;", "host: attempt counter of the 16-bit path off by one per launch"),
    # second batch: strip / colour-fused / generic kernels, conversions, host logic, clusters
    ("sweep_tiled.cuh", "const int eL = e0 == 0 ? rowlen - 4 : e0 - 4;", "const int eL = e0 == 0 ? rowlen - 8 : e0 - 4;", "strip kernel: left periodic halo chunk taken one chunk early"),
    ("sweep_tiled.cuh", "const int eR = e0 + 8 == rowlen ? 0 : e0 + 8;", "const int eR = e0 + 8 == rowlen ? 4 : e0 + 8;", "strip kernel: right periodic halo chunk off by one chunk"),
    ("sweep_tile2d.cuh", "      __syncthreads();\n    }\n\n    // store the interior", "    }\n\n    // store the interior", "colour-fused 2-D kernel: barrier between colour steps dropped"),
    ("sweep_tile2d.cuh", "const int ylo = c0 + 1, yhi = C::BH - c0 - 1;", "const int ylo = c0 + 1, yhi = C::BH - c0;", "[equivalent] colour-fused 2-D kernel: a colour step updates one halo row too many (the row is never stored and only later halo rows read it)"),
    ("sweep_tile2d.cuh", "if (++jg == g.gY) jg = 0;", "if (++jg == g.gY + 1) jg = 0;", "colour-fused 2-D kernel: global row of the random stream does not wrap"),
    ("sweep_tile2d.cuh", "    __syncthreads();  // the tile buffer is reused by the next item", "    // (barrier dropped)", "colour-fused 2-D kernel: barrier before the tile buffer is reused dropped"),
    ("sweep_generic.cuh", "int px = (c0 + j + k + g.g0) & 1;           // (i + j + kglobal) % 2 == c0", "int px = (c0 + j + k) & 1;           // (i + j + kglobal) % 2 == c0", "[equivalent] generic kernel: sc/6n colour parity ignores the slab offset (slabs are even, so is every offset)"),
    ("sweep_generic.cuh", "const unsigned long long leader = gsite - (unsigned long long)(philox_leader_shift<MODEL>(i) * NB);", "const unsigned long long leader = gsite - (unsigned long long)(philox_leader_shift<MODEL>(i));", "generic kernel: Philox block leader ignores the basis (tri)"),
    ("sweep_stream16.cuh", "dst[i] = make_uint4((uint32_t)a.x | ((uint32_t)a.y << 16), (uint32_t)a.z | ((uint32_t)a.w << 16),", "dst[i] = make_uint4((uint32_t)a.y | ((uint32_t)a.x << 16), (uint32_t)a.z | ((uint32_t)a.w << 16),", "shadow pack: first two sites of a chunk swapped"),
    ("sweep_bits2d.cuh", "const uint32_t even = p0 ? w1 : w0, odd = p0 ? w0 : w1;", "const uint32_t even = p0 ? w0 : w1, odd = p0 ? w1 : w0;", "bit-shadow unpack: colours swapped in odd rows"),
    ("aux_kernels.cuh", "if (i > range) i = range;\n    spin[s] = i - 1 + lo;", "if (i > range) i = range;\n    spin[s] = i + lo;", "set site range: values shifted by one"),
    ("aux_kernels.cuh", "uint32_t state = pm_mulmod(seed0, pm_powmod(16807u, first_draw + (unsigned long long)(first + lane) + 1ull));", "uint32_t state = pm_mulmod(seed0, pm_powmod(16807u, first_draw + (unsigned long long)(first + lane)));", "device pin: draw numbers shifted by one"),
    ("aux_kernels.cuh", "e += (spin[zk[o.dz + 1] * sz + yj[o.dy + 1] * sy + (long long)xi[o.dx + 1] * NB + o.m] != me);\n      }\n    }\n  }\n  e = __reduce_add_sync", "e += (spin[zk[o.dz + 1] * sz + yj[o.dy + 1] * sy + (long long)xi[o.dx + 1] * NB] != me);\n      }\n    }\n  }\n  e = __reduce_add_sync", "energy kernel: neighbour basis ignored (tri)"),
    ("cluster.cuh", "const unsigned peers = __match_any_sync(0xffffffffu, r);", "const unsigned peers = __match_any_sync(0xffffffffu, r) | 1u;", "clusters: warp-aggregated volume counts lane 0 everywhere"),
    ("spk_cuda.cu", "if (h->kstat_mean < h->neigh_switch - 0.25) h->neigh_adaptive = true;", "if (h->kstat_mean < h->neigh_switch - 0.25) h->neigh_adaptive = false;", "[equivalent] host: automatic variant switch never switches (both variants give the same spins; speed only)"),
    ("spk_cuda.cu", "  h->canon_current = false;\n    // the conversion above rewrites", "  // (stale flag dropped)\n    // the conversion above rewrites", "host: int32 lattice not marked stale after shadow sweeps on several ranks (only the multi-rank cases can see it)"),
    ("spk_cuda.cu", "    cudaFree(h->mask_base);\n    h->mask_base = h->mask = nullptr;", "    cudaFree(h->mask_base);\n    cudaFree(h->shadow_base);\n    cudaFree(h->bits);\n    h->mask_base = h->mask = nullptr;", "host: switching the masks off frees the shadows (the defect the device-lattice mode found)"),
    ("spk_cuda.cu", "ev.thr[dE] = t > 4294967295.0 ? 4294967295u : (uint32_t)t;", "ev.thr[dE] = t > 4294967295.0 ? 4294967295u : (uint32_t)t + 1u;", "host: acceptance thresholds one too high"),
]

TESTS = {"cluster.cuh": ["tests/test_cluster.py", "-k", "not large"]}
ALSO = ["tests/test_zz_device_pin.py", "-k", "not script and not large"]  # run after the parity files for aux_kernels.cuh


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests", "host_shim"))
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    last = int(sys.argv[2]) if len(sys.argv) > 2 else len(MUTANTS) - 1
    survivors = []
    for k, (fname, old, new, what) in enumerate(MUTANTS):
        if k < first or k > last:
            continue
        tmp = tempfile.mkdtemp(prefix="spk_mutant_")
        try:
            src = os.path.join(tmp, "csrc")
            shutil.copytree(CSRC, src)
            text = open(os.path.join(src, fname)).read()
            if text.count(old) != 1:
                print("mutant %2d  NOT APPLICABLE (%d matches): %s" % (k, text.count(old), what), flush=True)
                survivors.append(k)
                continue
            open(os.path.join(src, fname), "w").write(text.replace(old, new))
            env = dict(os.environ, SPK_EMU_CSRC=src, SPK_EMU_BUILD_DIR=os.path.join(tmp, "build"))
            b = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_shim", "build_emu.py")], env=env, capture_output=True, text=True)
            if b.returncode != 0:
                print("mutant %2d  KILLED (does not build): %s" % (k, what), flush=True)
                continue
            env = dict(os.environ, SPK_CUDA_LIB=os.path.join(tmp, "build", "libspparks_cuda_emu.so"), SPK_CUDA_ALLOW_EMULATOR="1")
            args = TESTS.get(fname, ["tests/test_zz_more_parity.py", "tests/test_gpu_parity.py"])
            if fname == "aux_kernels.cuh":
                args = ["tests/test_zz_device_pin.py"] + args + ["-k", "not script and not large"]
            dead, how = False, ""
            # plain first; a survivor gets a second chance under the emulator's schedule checks
            for extra in ({}, {"SPK_EMU_ORDER": "shuffle:7", "SPK_EMU_TMA": "lazy"}):
                try:
                    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"] + args, cwd=ROOT,
                                       env=dict(env, **extra), capture_output=True, text=True, timeout=300)
                    dead = r.returncode != 0
                    how = (r.stdout.strip().splitlines() or ["?"])[-1] + (" under shuffled threads + lazy copies" if extra else "")
                except subprocess.TimeoutExpired:
                    dead, how = True, "timeout (hang)"
                if dead:
                    break
            if not dead:  # last chance: several ranks (random cases of tools/emu_fuzz_ranks.py, both halo modes)
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_fuzz_ranks.py"), "--ranks", "2", "--cases", "60", "--seed", "5"],
                                   cwd=ROOT, env=dict(os.environ, SPK_EMU_USE_LIB=env["SPK_CUDA_LIB"]), capture_output=True, text=True)
                dead = r.returncode != 0
                how = "2 ranks: " + (r.stdout.strip().splitlines() or ["?"])[0]
            print("mutant %2d  %s: %s   [%s]" % (k, "KILLED" if dead else "SURVIVED", what, how[:100]), flush=True)
            if not dead and not what.startswith("[equivalent]"):
                survivors.append(k)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    print("%d mutants, %d unexpected survivors %s" % (last - first + 1, len(survivors), survivors))
    return 1 if survivors else 0


if __name__ == "__main__":
    sys.exit(main())
