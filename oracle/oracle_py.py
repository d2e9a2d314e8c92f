"""TEST INFRASTRUCTURE — ctypes binding of oracle/liboracle.so (the CPU oracle, spk_oracle.c).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs.
The product package spparks_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

SQ_4N, SQ_8N, TRI, SC_6N, SC_26N = 2, 3, 4, 5, 6
ISING, POTTS, POTTS_NEIGHONLY, POTTS_PIN, ISING_SINGLE, POTTS_NEIGH = 0, 1, 2, 3, 4, 5
LATTICES = {"sq/4n": SQ_4N, "sq/8n": SQ_8N, "tri": TRI, "sc/6n": SC_6N, "sc/26n": SC_26N}
MODELS = {"ising": ISING, "potts": POTTS, "potts/neighonly": POTTS_NEIGHONLY, "potts/pin": POTTS_PIN,
          "ising/single": ISING_SINGLE, "potts/neigh": POTTS_NEIGH}

IM = 2147483647
IA = 16807


class Cfg(C.Structure):
    _fields_ = [("model", C.c_int), ("lattice", C.c_int), ("nx", C.c_int), ("ny", C.c_int),
                ("nz", C.c_int), ("nspins", C.c_int), ("temperature", C.c_double)]


class Mars(C.Structure):
    _fields_ = [("u", C.c_double * 98), ("i97", C.c_int), ("j97", C.c_int),
                ("c", C.c_double), ("cd", C.c_double), ("cm", C.c_double)]


class StatsRow(C.Structure):
    _fields_ = [("time", C.c_double), ("naccept", C.c_int64), ("nreject", C.c_int64),
                ("nsweeps", C.c_int), ("energy", C.c_double)]


def build():
    """Compile the oracle if the shared object is missing or stale."""
    src = os.path.join(_HERE, "spk_oracle.c")
    if (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        i32p = C.POINTER(C.c_int32)
        i64p = C.POINTER(C.c_int64)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(Cfg)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_nsites.restype = C.c_int64
        L.orc_nsites.argtypes = [C.c_void_p]
        L.orc_neighbors.argtypes = [C.c_void_p, C.c_int64, i64p]
        L.orc_site_color.argtypes = [C.c_void_p, C.c_int64]
        L.orc_site_sector.argtypes = [C.c_void_p, C.c_int, C.c_int64]
        L.orc_sector_box.argtypes = [C.c_void_p, C.c_int, C.c_int, ip]
        L.orc_site_energy.argtypes = [C.c_void_p, i32p, C.c_int64]
        L.orc_total_energy.restype = C.c_double
        L.orc_total_energy.argtypes = [C.c_void_p, i32p]
        L.orc_halfsweep_injected.restype = C.c_int64
        L.orc_halfsweep_injected.argtypes = [C.c_void_p, i32p, C.c_int, ip, dp, dp]
        L.orc_halfsweep_philox.restype = C.c_int64
        L.orc_halfsweep_philox.argtypes = [C.c_void_p, i32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, ip]
        L.orc_halfsweep_philox_double.restype = C.c_int64
        L.orc_halfsweep_philox_double.argtypes = [C.c_void_p, i32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, ip]
        L.orc_site_event_words.argtypes = [C.c_void_p, i32p, C.c_int64, C.c_uint32, C.c_uint32]
        L.orc_site_event_words_double.argtypes = [C.c_void_p, i32p, C.c_int64, C.c_uint32, C.c_uint32]
        L.orc_boundary_clear_mask.argtypes = [C.c_void_p, C.c_void_p, ip]
        L.orc_halfsweep_injected_mask.restype = C.c_int64
        L.orc_halfsweep_injected_mask.argtypes = [C.c_void_p, i32p, C.c_void_p, C.c_int, ip, dp, dp]
        L.orc_halfsweep_philox_mask.restype = C.c_int64
        L.orc_halfsweep_philox_mask.argtypes = [C.c_void_p, i32p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, ip]
        L.orc_sweeps_strict.restype = C.c_int64
        L.orc_sweeps_strict.argtypes = [C.c_void_p, i32p, i32p, C.c_int]
        L.orc_strict_siteseeds.argtypes = [C.c_void_p, C.POINTER(Mars), i32p]
        L.orc_set_site_range.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, C.POINTER(Mars)]
        L.orc_pin_create.restype = C.c_int64
        L.orc_pin_create.argtypes = [C.c_void_p, i32p, C.c_double, C.c_int, C.c_int, C.POINTER(Mars)]
        L.orc_mars_init.argtypes = [C.POINTER(Mars), C.c_int]
        L.orc_mars_uniform.restype = C.c_double
        L.orc_mars_uniform.argtypes = [C.POINTER(Mars)]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_accept_threshold.restype = C.c_uint32
        L.orc_accept_threshold.argtypes = [C.c_double, C.c_int]
        L.orc_run_sweep_random.argtypes = [C.c_void_p, i32p, C.POINTER(Mars), C.c_int, C.c_double,
                                           C.c_double, C.POINTER(StatsRow), C.c_int]
        L.orc_clusters.restype = C.c_int64
        L.orc_clusters.argtypes = [C.c_void_p, i32p, i32p, i64p, dp, dp]
        L.orc_nbasis.argtypes = [C.c_int]
        L.orc_maxneigh.argtypes = [C.c_int]
        L.orc_ncolors.argtypes = [C.c_int]
        _lib = L
    return _lib


def _i32(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _box(box):
    if box is None:
        return None
    return (C.c_int * 6)(*[int(b) for b in box])


class Oracle:
    """One lattice + application; spins are numpy int32 arrays in reference site order."""

    def __init__(self, model, lattice, nx, ny, nz=1, nspins=2, temperature=0.0):
        self.model = MODELS[model] if isinstance(model, str) else model
        self.lattice = LATTICES[lattice] if isinstance(lattice, str) else lattice
        if self.model in (ISING, ISING_SINGLE):
            nspins = 2
        self.cfg = Cfg(self.model, self.lattice, nx, ny, nz, nspins, temperature)
        self.L = lib()
        self.ctx = self.L.orc_create(C.byref(self.cfg))
        self.nsites = self.L.orc_nsites(self.ctx)
        self.maxneigh = self.L.orc_maxneigh(self.lattice)
        self.ncolors = self.L.orc_ncolors(self.lattice)
        self.nbasis = self.L.orc_nbasis(self.lattice)

    def __del__(self):
        try:
            self.L.orc_destroy(self.ctx)
        except Exception:
            pass

    def neighbors(self, site):
        out = (C.c_int64 * self.maxneigh)()
        self.L.orc_neighbors(self.ctx, site, out)
        return list(out)

    def neighbor_table(self):
        t = np.empty((self.nsites, self.maxneigh), dtype=np.int64)
        for s in range(self.nsites):
            t[s] = self.neighbors(s)
        return t

    def colors(self):
        return np.array([self.L.orc_site_color(self.ctx, s) for s in range(self.nsites)], dtype=np.int32)

    def sectors(self, nsector):
        return np.array([self.L.orc_site_sector(self.ctx, nsector, s) for s in range(self.nsites)],
                        dtype=np.int32)

    def sector_box(self, nsector, isector):
        b = (C.c_int * 6)()
        self.L.orc_sector_box(self.ctx, nsector, isector, b)
        return list(b)

    def total_energy(self, spin):
        return self.L.orc_total_energy(self.ctx, _i32(spin))

    def clusters(self, spin):
        """diag_style cluster: (ncluster, <N>, <R>, cluster_ids[nsites] (1-based, reference numbering),
        volumes[ncluster])."""
        ids = np.zeros(self.nsites, dtype=np.int32)
        vol = np.zeros(self.nsites, dtype=np.int64)
        mv, mr = C.c_double(), C.c_double()
        n = int(self.L.orc_clusters(self.ctx, _i32(spin), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                    vol.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(mv), C.byref(mr)))
        return n, mv.value, mr.value, ids, vol[:n].copy()

    def halfsweep_injected(self, spin, color, u1, u2, box=None):
        return self.L.orc_halfsweep_injected(self.ctx, _i32(spin), color, _box(box), _f64(u1), _f64(u2))

    def boundary_clear_mask(self, mask, box):
        assert mask.dtype == np.uint8 and mask.flags.c_contiguous
        self.L.orc_boundary_clear_mask(self.ctx, mask.ctypes.data, _box(box))

    def halfsweep_injected_mask(self, spin, mask, color, u1, u2, box=None):
        assert mask.dtype == np.uint8 and mask.flags.c_contiguous
        return self.L.orc_halfsweep_injected_mask(self.ctx, _i32(spin), mask.ctypes.data, color, _box(box), _f64(u1), _f64(u2))

    def halfsweep_philox_mask(self, spin, mask, seed, sweep, color, iloop=0, box=None):
        assert mask.dtype == np.uint8 and mask.flags.c_contiguous
        return self.L.orc_halfsweep_philox_mask(self.ctx, _i32(spin), mask.ctypes.data, seed, sweep, color, iloop, _box(box))

    def halfsweep_philox(self, spin, seed, sweep, color, iloop=0, box=None):
        return self.L.orc_halfsweep_philox(self.ctx, _i32(spin), seed, sweep, color, iloop, _box(box))

    def halfsweep_philox_double(self, spin, seed, sweep, color, iloop=0, box=None):
        return self.L.orc_halfsweep_philox_double(self.ctx, _i32(spin), seed, sweep, color, iloop, _box(box))

    def site_event_words(self, spin, site, w0, w1, double=False):
        f = self.L.orc_site_event_words_double if double else self.L.orc_site_event_words
        return f(self.ctx, _i32(spin), site, w0, w1)

    def sweep_philox(self, spin, seed, sweep):
        acc = 0
        for color in range(1, self.ncolors + 1):
            acc += self.halfsweep_philox(spin, seed, sweep, color)
        return acc

    def sweeps_strict(self, spin, siteseeds, nsweeps):
        return self.L.orc_sweeps_strict(self.ctx, _i32(spin), _i32(siteseeds), nsweeps)

    def strict_siteseeds(self, mars):
        ss = np.empty(self.nsites, dtype=np.int32)
        self.L.orc_strict_siteseeds(self.ctx, C.byref(mars), _i32(ss))
        return ss

    def set_site_range(self, lo, hi, mars):
        spin = np.empty(self.nsites, dtype=np.int32)
        self.L.orc_set_site_range(self.ctx, _i32(spin), lo, hi, C.byref(mars))
        return spin

    def pin_create(self, spin, fraction, multi, nthresh, mars):
        return self.L.orc_pin_create(self.ctx, _i32(spin), fraction, multi, nthresh, C.byref(mars))

    def run_sweep_random(self, spin, mars, sectors, stats_delta, run_time, maxrows=4096):
        rows = (StatsRow * maxrows)()
        n = self.L.orc_run_sweep_random(self.ctx, _i32(spin), C.byref(mars), int(sectors), stats_delta,
                                        run_time, rows, maxrows)
        return [(r.time, r.naccept, r.nreject, r.nsweeps, r.energy) for r in rows[:n]]


def mars(seed):
    m = Mars()
    lib().orc_mars_init(C.byref(m), seed)
    return m


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def accept_threshold(temperature, dE):
    return lib().orc_accept_threshold(temperature, dE)


def injected_uniforms(k):
    """Park-Miller pair for injected seeds k (int array in [1, IM-1]): u1 = k/IM and
    u2 = (16807 k mod IM)/IM, computed as the reference does (AM*seed, src/random_park.cpp:84)."""
    k = np.asarray(k, dtype=np.int64)
    am = 1.0 / IM
    u1 = am * k.astype(np.float64)
    u2 = am * ((k * IA) % IM).astype(np.float64)
    return u1, u2
