"""LatticeSweep: thin object wrapper over the libspparks_cuda C ABI (one slab on one GPU)."""
import ctypes as C

import numpy as np

from . import _lib

MODELS = {"ising": _lib.ISING, "potts": _lib.POTTS, "potts/neighonly": _lib.POTTS_NEIGHONLY,
          "potts/pin": _lib.POTTS_PIN, "ising/single": _lib.ISING_SINGLE, "potts/neigh": _lib.POTTS_NEIGH}
LATTICES = {"sq/4n": _lib.SQ_4N, "sq/8n": _lib.SQ_8N, "tri": _lib.TRI, "sc/6n": _lib.SC_6N,
            "sc/26n": _lib.SC_26N}


def _box(box):
    if box is None:
        return None
    return (C.c_int32 * 6)(*[int(b) for b in box])


class LatticeSweep:
    """Device-resident lattice of one app_style on one lattice style.

    Spins travel as numpy int32 arrays in reference site order (x fastest, basis fastest of all;
    src/create_sites.cpp:344-350).  With nranks > 1 the object owns slab `rank` of the last
    dimension and needs comm_init() before halo exchanges.
    """

    def __init__(self, model, lattice, nx, ny, nz=1, nspins=2, temperature=0.0, seed=1, device=0,
                 rank=0, nranks=1):
        self.L = _lib.load()
        self.model = MODELS[model] if isinstance(model, str) else int(model)
        self.lattice = LATTICES[lattice] if isinstance(lattice, str) else int(lattice)
        self.desc = _lib.Desc(self.model, self.lattice, nx, ny, nz, nspins, temperature, seed, device,
                              rank, nranks)
        self.h = C.c_void_p()
        _lib.check(self.L.spk_cuda_create(C.byref(self.desc), C.byref(self.h)))
        self.nlocal = self.L.spk_cuda_nlocal(self.h)
        self.nglobal = self.L.spk_cuda_nglobal(self.h)
        self.ncolors = self.L.spk_cuda_ncolors(self.lattice)
        self.maxneigh = self.L.spk_cuda_maxneigh(self.lattice)
        self.nbasis = self.L.spk_cuda_nbasis(self.lattice)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.spk_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data movement -------------------------------------------------------------------------
    def upload(self, spins):
        a = np.ascontiguousarray(spins, dtype=np.int32)
        _lib.check(self.L.spk_cuda_upload_spins(self.h, a.ctypes.data, a.size))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.nlocal, dtype=np.int32)
        _lib.check(self.L.spk_cuda_download_spins(self.h, out.ctypes.data, out.size))
        return out

    def upload_ptr(self, ptr, n):
        _lib.check(self.L.spk_cuda_upload_spins(self.h, ptr, n))

    def download_ptr(self, ptr, n):
        _lib.check(self.L.spk_cuda_download_spins(self.h, ptr, n))

    def randomize(self, seed, lo, hi):
        _lib.check(self.L.spk_cuda_randomize(self.h, seed, lo, hi))

    def set_site_range(self, park_seed, lo, hi):
        """`set site range lo hi` with the reference's own Park-Miller stream (state park_seed)."""
        _lib.check(self.L.spk_cuda_set_site_range(self.h, int(park_seed), int(lo), int(hi)))

    def site_energy(self):
        out = np.empty(self.nlocal, dtype=np.uint8)
        _lib.check(self.L.spk_cuda_site_energy(self.h, out.ctypes.data_as(C.c_void_p), self.nlocal))
        return out

    def set_seed(self, seed):
        _lib.check(self.L.spk_cuda_set_seed(self.h, seed))

    def accept_thresholds(self, n=27):
        out = (C.c_uint32 * n)()
        _lib.check(self.L.spk_cuda_accept_thresholds(self.h, out, n))
        return list(out)

    def pin_random(self, seed, fraction):
        _lib.check(self.L.spk_cuda_pin_random(self.h, seed, fraction))

    def pin_create(self, park_seed, fraction, multi=0, nthresh=0):
        """`pin fraction multi nthresh` with the reference's own stream; returns the pinned sites of the lattice"""
        n = C.c_int64()
        _lib.check(self.L.spk_cuda_pin_create(self.h, fraction, multi, nthresh, park_seed, C.byref(n)))
        return n.value

    def set_temperature(self, t):
        _lib.check(self.L.spk_cuda_set_temperature(self.h, t))

    # -- sweeps --------------------------------------------------------------------------------
    def sweep_color(self, sweep, color, box=None, nloop=1):
        _lib.check(self.L.spk_cuda_sweep_color(self.h, sweep, color, _box(box), nloop))

    def sweep_color_injected(self, color, u1, u2, box=None):
        u1 = np.ascontiguousarray(u1, dtype=np.float64)
        u2 = np.ascontiguousarray(u2, dtype=np.float64)
        assert u1.size == self.nlocal and u2.size == self.nlocal
        dp = C.POINTER(C.c_double)
        _lib.check(self.L.spk_cuda_sweep_color_injected(self.h, color, _box(box), u1.ctypes.data_as(dp),
                                                        u2.ctypes.data_as(dp)))

    def set_siteseeds(self, seeds):
        a = np.ascontiguousarray(seeds, dtype=np.int32)
        _lib.check(self.L.spk_cuda_set_siteseeds(self.h, a.ctypes.data, a.size))

    def get_siteseeds(self):
        out = np.empty(self.nlocal, dtype=np.int32)
        _lib.check(self.L.spk_cuda_get_siteseeds(self.h, out.ctypes.data, out.size))
        return out

    def sweep_color_strict(self, color, box=None, nloop=1):
        _lib.check(self.L.spk_cuda_sweep_color_strict(self.h, color, _box(box), nloop))

    def sweeps(self, first_sweep, nsweeps):
        _lib.check(self.L.spk_cuda_sweeps(self.h, first_sweep, nsweeps))

    # -- observables ---------------------------------------------------------------------------
    def counters(self):
        a, b = C.c_int64(), C.c_int64()
        _lib.check(self.L.spk_cuda_counters(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def reset_counters(self):
        _lib.check(self.L.spk_cuda_reset_counters(self.h))

    def energy(self):
        e = C.c_double()
        _lib.check(self.L.spk_cuda_energy(self.h, C.byref(e)))
        return e.value

    def cluster_stats(self):
        """diag_style cluster on the device: (Nclust, <N>, <R>) of the reference's stats line."""
        n, mv, mr = C.c_int64(), C.c_double(), C.c_double()
        _lib.check(self.L.spk_cuda_cluster_stats(self.h, C.byref(n), C.byref(mv), C.byref(mr)))
        return int(n.value), mv.value, mr.value

    def cluster_ids(self):
        """(ids[nlocal] int32 in the reference's numbering, volumes[ncluster] int64)."""
        nc = self.cluster_stats()[0]
        ids = np.empty(self.nlocal, dtype=np.int32)
        vol = np.empty(nc, dtype=np.int64)
        _lib.check(self.L.spk_cuda_cluster_ids(self.h, ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                               vol.ctypes.data_as(C.POINTER(C.c_int64)), nc))
        return ids, vol

    # -- multi-GPU -----------------------------------------------------------------------------
    @staticmethod
    def comm_id():
        buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
        _lib.check(_lib.load().spk_cuda_comm_id(buf))
        return buf.raw

    def comm_init(self, comm_id):
        _lib.check(self.L.spk_cuda_comm_init(self.h, comm_id))

    def halo_exchange(self):
        _lib.check(self.L.spk_cuda_halo_exchange(self.h))

    def halo_mode(self):
        """0 = single GPU, 1 = NCCL exchange per colour step, 2 = peer-memory stores (no exchange)."""
        return int(self.L.spk_cuda_halo_mode(self.h))

    def mask_enable(self, on=True):
        _lib.check(self.L.spk_cuda_mask_enable(self.h, 1 if on else 0))

    def mask_clear_border(self, box=None, color=0):
        _lib.check(self.L.spk_cuda_mask_clear_border(self.h, _box(box), color))

    def upload_mask(self, mask):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        assert mask.size == self.nlocal
        _lib.check(self.L.spk_cuda_upload_mask(self.h, mask.ctypes.data_as(C.c_void_p), self.nlocal))

    def download_mask(self):
        out = np.empty(self.nlocal, dtype=np.uint8)
        _lib.check(self.L.spk_cuda_download_mask(self.h, out.ctypes.data_as(C.c_void_p), self.nlocal))
        return out

    def sweep_path(self):
        """(kernel name, launches per sweep, bytes per spin of the copy it sweeps on; 0 = one bit)."""
        buf = C.create_string_buffer(128)
        per, width = C.c_int(), C.c_int()
        _lib.check(self.L.spk_cuda_sweep_path(self.h, buf, 128, C.byref(per), C.byref(width)))
        return buf.value.decode(), per.value, width.value

    def digest(self):
        """Order-independent 64-bit digest of the owned spins (global site index, value)."""
        d = C.c_uint64()
        _lib.check(self.L.spk_cuda_digest(self.h, C.byref(d)))
        return d.value

    def neigh_variant(self, with_mean=False):
        """Candidate-list apps on sc/26n: "chain" | "adaptive" (None when not applicable)."""
        v, m = C.c_int(), C.c_double()
        _lib.check(self.L.spk_cuda_neigh_variant(self.h, C.byref(v), C.byref(m)))
        name = {0: "chain", 1: "adaptive"}.get(v.value)
        return (name, m.value) if with_mean else name

    # -- stream / timing -----------------------------------------------------------------------
    def sync(self):
        _lib.check(self.L.spk_cuda_sync(self.h))

    def timer_start(self):
        _lib.check(self.L.spk_cuda_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        _lib.check(self.L.spk_cuda_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def timer_lap(self):
        _lib.check(self.L.spk_cuda_timer_lap(self.h))

    def timer_laps(self):
        """Device times (ms) between consecutive laps since timer_start."""
        n = C.c_int()
        _lib.check(self.L.spk_cuda_timer_laps(self.h, None, 0, C.byref(n)))
        buf = (C.c_float * max(n.value, 1))()
        _lib.check(self.L.spk_cuda_timer_laps(self.h, buf, n.value, C.byref(n)))
        return [float(buf[i]) for i in range(n.value)]

    def launch_count(self):
        n = C.c_int64()
        _lib.check(self.L.spk_cuda_launch_count(self.h, C.byref(n)))
        return n.value

    def device_ptr(self):
        p = C.c_void_p()
        _lib.check(self.L.spk_cuda_device_spins(self.h, C.byref(p)))
        return p.value
