"""TEST INFRASTRUCTURE.  Builds libspparks_cuda for the HOST: tests/host_shim/_build/libspparks_cuda_emu.so.

The sources of spparks_b200/csrc are compiled with g++ against the small CUDA emulator of
tests/host_shim/emu (cuda_emu.h says what it does and does not model).  The sources are used as they
are except for the textual substitutions listed in this file, each checked to apply exactly as
often as expected, so a change of the product source that the table does not know fails the build
instead of going unnoticed:

* `kernel<<<grid, block, smem, stream>>>(args);`  ->  `emu::launch("kernel", grid, block, smem, [&] { kernel(args); });`
* the five PTX wrappers of the mbarrier / bulk-tensor-copy instructions -> the emulator's functions
* `ld.shared` through a 32-bit shared-memory address -> a load from the emulated shared memory
* `fence.*` -> a host fence; `%globaltimer` -> the host clock
* `extern __shared__ ... smem_raw[]` -> the emulated dynamic shared memory

Arithmetic written in PTX (event.cuh: unlike(); dedupe_chain.cuh; sweep_bits2d.cuh: shift_in_reject)
has a plain C++ twin under SPK_HOST_SHIM next to it in the product source.

Usage: python tests/host_shim/build_emu.py [-j N]   (tests/test_emulated_library.py calls build()).
"""
import concurrent.futures
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# (the two overrides are for mutation experiments: build a patched copy of the sources somewhere else)
CSRC = os.environ.get("SPK_EMU_CSRC") or os.path.join(ROOT, "spparks_b200", "csrc")
EMU = os.path.join(HERE, "emu")
BUILD = os.environ.get("SPK_EMU_BUILD_DIR") or os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libspparks_cuda_emu.so")
NCCL_LIB = os.path.join(BUILD, "libnccl_emu.so")  # for SPK_CUDA_NCCL_LIB: ranks = processes of this host
UNITS = ["spk_cuda.cu", "model_ising.cu", "model_potts.cu", "model_potts_neighonly.cu", "model_potts_pin.cu",
         "model_ising_single.cu", "model_potts_neigh.cu"]
CXXFLAGS = ["-std=c++17", "-O1", "-fPIC", "-pthread", "-w", "-fno-strict-aliasing"]


def _match_back(text, pos):
    """start of the kernel expression that ends at text[pos-1] (identifier, optionally with template arguments)"""
    i = pos
    while text[i - 1].isspace():
        i -= 1
    if text[i - 1] == ">":
        depth = 0
        while True:
            i -= 1
            if text[i] == ">":
                depth += 1
            elif text[i] == "<":
                depth -= 1
                if depth == 0:
                    break
    while re.match(r"[A-Za-z0-9_:]", text[i - 1]):
        i -= 1
    return i


def _match_paren(text, pos):
    """index just past the parenthesis that closes the one at text[pos]"""
    assert text[pos] == "(", text[pos:pos + 40]
    depth = 0
    for i in range(pos, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced parentheses")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(text):
    """kernel<<<cfg>>>(args);  ->  ::emu::launch(..., [&]() { kernel(args); });   returns (text, count)"""
    n = 0
    while True:
        p = text.find("<<<")
        if p < 0:
            return text, n
        start = _match_back(text, p)
        q = text.index(">>>", p)
        cfg = _split_top(text[p + 3:q])
        assert 2 <= len(cfg) <= 4, cfg
        a = q + 3
        while text[a].isspace():
            a += 1
        end = _match_paren(text, a)
        kernel, args = text[start:p].strip(), text[a:end]
        smem = cfg[2] if len(cfg) > 2 else "0"
        label = re.sub(r"\s+", " ", kernel).replace('"', "'")
        repl = '::emu::launch("%s", ::emu::fn_addr(%s), dim3(%s), dim3(%s), (size_t)(%s), [&]() { %s%s; })' % (label, kernel, cfg[0], cfg[1], smem, kernel, args)
        text = text[:start] + repl + text[end:]
        n += 1


TMA_WRAPPERS = """\
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
inline void mbar_init(uint64_t *bar, int count) { ::emu::mbar_init(bar, count); }
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { ::emu::mbar_expect_tx(bar, bytes); }
inline void mbar_wait(uint64_t *bar, uint32_t parity) { ::emu::mbar_wait(bar, parity); }
inline void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
  const int c[3] = {x, y, z};
  ::emu::tma_load(dst, map, bar, c, 3);
}

"""

# (file, how, pattern, replacement, expected count).  how: "str" exact text, "re" regular expression (DOTALL),
# "block" = everything from the line containing pattern[0] up to (not including) the line containing pattern[1]
SUBSTITUTIONS = [
    ("sweep_stream.cuh", "block", ("// ---- mbarrier / TMA wrappers", "// Periodic images of the halo cells of one smem plane"),
     TMA_WRAPPERS, 1),
    ("sweep_stream.cuh", "re", r'asm\("ld\.shared\.s32 %0, \[%1\];" : "=r"\((\w+)\) : "r"\((.*?)\)\);',
     r"\1 = *reinterpret_cast<const int *>(::emu::dynamic_smem() + (\2));", 6),
    ("sweep_tile2d.cuh", "re",
     r'asm volatile\(\s*"cp\.async\.bulk\.tensor\.2d[^;]*?;"\s*::"r"\(smem_u32\(dst\)\), "l"\(map\), "r"\(smem_u32\(bar\)\), "r"\(x\), "r"\(y\) : "memory"\);',
     "{ const int c[2] = {x, y}; ::emu::tma_load(dst, map, bar, c, 2); }", 1),
    ("aux_kernels.cuh", "re", r'asm volatile\("mov\.u64 %0, %%globaltimer;" : "=l"\((\w+)\)\);', r"\1 = ::emu::globaltimer();", 2),
    ("*", "re", r'asm volatile\("fence\.(?:mbarrier_init\.release\.cluster|proxy\.async\.shared::cta);" ::: "memory"\);',
     "__threadfence_block();", 8),
    ("*", "str", "extern __shared__ __align__(128) unsigned char smem_raw[];", "unsigned char *smem_raw = ::emu::dynamic_smem();", 3),
    ("*", "str", '#include "../../include/spparks_cuda.h"', '#include "%s"' % os.path.join(ROOT, "include", "spparks_cuda.h"), None),
]


def rewrite_sources(outdir):
    os.makedirs(outdir, exist_ok=True)
    counts = [0] * len(SUBSTITUTIONS)
    launches = 0
    for name in sorted(os.listdir(CSRC)):
        if not name.endswith((".cu", ".cuh", ".h")):
            continue
        text = open(os.path.join(CSRC, name)).read()
        for k, (fname, how, pat, repl, _) in enumerate(SUBSTITUTIONS):
            if fname not in ("*", name):
                continue
            if how == "str":
                counts[k] += text.count(pat)
                text = text.replace(pat, repl)
            elif how == "re":
                text, c = re.subn(pat, repl, text, flags=re.DOTALL)
                counts[k] += c
            else:
                lines = text.split("\n")
                a = [i for i, l in enumerate(lines) if pat[0] in l]
                b = [i for i, l in enumerate(lines) if pat[1] in l]
                if len(a) == 1 and len(b) == 1 and a[0] < b[0]:
                    text = "\n".join(lines[:a[0]] + repl.split("\n") + lines[b[0]:])
                    counts[k] += 1
        text, c = rewrite_launches(text)
        launches += c
        if "asm" in re.sub(r"//[^\n]*", "", text) and "SPK_HOST_SHIM" not in text:
            raise SystemExit("build_emu: %s still holds inline assembly after the substitutions" % name)
        out = os.path.join(outdir, name if not name.endswith(".cu") else name[:-3] + ".cpp")
        if not os.path.exists(out) or open(out).read() != text:
            open(out, "w").write(text)
    for k, (fname, how, pat, repl, want) in enumerate(SUBSTITUTIONS):
        if want is not None and counts[k] != want:
            raise SystemExit("build_emu: substitution %d (%s in %s) applied %d times, expected %d: the product source "
                             "changed, update tests/host_shim/build_emu.py" % (k, how, fname, counts[k], want))
    if launches < 40:
        raise SystemExit("build_emu: only %d kernel launches found" % launches)
    return launches


def _compile(src, obj, extra):
    cmd = ["g++"] + CXXFLAGS + ["-include", os.path.join(EMU, "cuda_emu.h"), "-I", EMU] + extra + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit("build_emu: %s\n%s" % (" ".join(cmd), r.stderr[-6000:]))


def _stamp():
    h = hashlib.sha256()
    for d in (CSRC, EMU, HERE):
        for name in sorted(os.listdir(d)):
            p = os.path.join(d, name)
            if os.path.isfile(p) and name.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                h.update(name.encode())
                h.update(open(p, "rb").read())
    h.update(open(os.path.join(ROOT, "include", "spparks_cuda.h"), "rb").read())
    return h.hexdigest()


def build(jobs=None, verbose=False):
    """build (if anything changed) and return the path of the emulated library"""
    stamp_file = os.path.join(BUILD, "stamp")
    stamp = _stamp()
    if os.path.exists(LIB) and os.path.exists(NCCL_LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    src = os.path.join(BUILD, "src")
    n = rewrite_sources(src)
    if verbose:
        print("build_emu: %d kernel launches rewritten" % n)
    work = [(os.path.join(src, u[:-3] + ".cpp"), os.path.join(BUILD, u[:-3] + ".o"), ["-I", src]) for u in UNITS]
    work.append((os.path.join(EMU, "cuda_emu.cpp"), os.path.join(BUILD, "cuda_emu.o"), []))
    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs or os.cpu_count()) as ex:
        list(ex.map(lambda w: _compile(*w), work))
    cmd = ["g++", "-shared", "-pthread", "-o", LIB] + [w[1] for w in work] + ["-ldl"]
    subprocess.check_call(cmd)
    subprocess.check_call(["g++"] + CXXFLAGS + ["-shared", "-I", EMU, os.path.join(EMU, "nccl_emu.cpp"), "-o", NCCL_LIB])
    open(stamp_file, "w").write(stamp)
    return LIB


if __name__ == "__main__":
    j = int(sys.argv[sys.argv.index("-j") + 1]) if "-j" in sys.argv else None
    print(build(j, verbose=True))
