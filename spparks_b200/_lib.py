"""ctypes binding of libspparks_cuda (include/spparks_cuda.h).

There is no CPU fallback: if the shared library is missing, or a call fails (e.g. no CUDA
device), a SpkCudaError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPK_CUDA_LIB selects another build of the same library (kernel A/B experiments only)
LIB_PATH = os.environ.get("SPK_CUDA_LIB") or os.path.join(_HERE, "lib", "libspparks_cuda.so")

ISING, POTTS, POTTS_NEIGHONLY, POTTS_PIN, ISING_SINGLE, POTTS_NEIGH = 0, 1, 2, 3, 4, 5
SQ_4N, SQ_8N, TRI, SC_6N, SC_26N = 2, 3, 4, 5, 6
COMM_ID_BYTES = 128

# every symbol include/spparks_cuda.h declares (tests check that the library exports them all)
SYMBOLS = [
    "spk_cuda_last_error", "spk_cuda_abi_version", "spk_cuda_device_count", "spk_cuda_create",
    "spk_cuda_destroy", "spk_cuda_ncolors", "spk_cuda_maxneigh", "spk_cuda_nbasis", "spk_cuda_nlocal",
    "spk_cuda_nglobal", "spk_cuda_upload_spins", "spk_cuda_download_spins", "spk_cuda_set_temperature",
    "spk_cuda_randomize", "spk_cuda_pin_random", "spk_cuda_sweep_color", "spk_cuda_sweep_color_injected",
    "spk_cuda_set_siteseeds", "spk_cuda_get_siteseeds", "spk_cuda_sweep_color_strict", "spk_cuda_sweeps",
    "spk_cuda_counters", "spk_cuda_reset_counters", "spk_cuda_energy", "spk_cuda_comm_id",
    "spk_cuda_comm_init", "spk_cuda_halo_exchange", "spk_cuda_halo_mode", "spk_cuda_sync", "spk_cuda_timer_start",
    "spk_cuda_timer_stop", "spk_cuda_launch_count", "spk_cuda_device_spins", "spk_cuda_cluster_stats",
    "spk_cuda_cluster_ids", "spk_cuda_neigh_variant", "spk_cuda_timer_lap", "spk_cuda_timer_laps",
    "spk_cuda_digest", "spk_cuda_mask_enable", "spk_cuda_mask_clear_border", "spk_cuda_download_mask",
    "spk_cuda_host_register", "spk_cuda_host_unregister", "spk_cuda_set_site_range", "spk_cuda_site_energy",
    "spk_cuda_sweep_path", "spk_cuda_pin_create", "spk_cuda_accept_thresholds", "spk_cuda_set_seed", "spk_cuda_validate_spins", "spk_cuda_upload_mask",
]


class SpkCudaError(RuntimeError):
    pass


class Desc(C.Structure):
    _fields_ = [("model", C.c_int32), ("lattice", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
                ("nz", C.c_int32), ("nspins", C.c_int32), ("temperature", C.c_double),
                ("seed", C.c_uint64), ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32)]


_lib = None


def load():
    """Load libspparks_cuda.so; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpkCudaError("libspparks_cuda.so not built (%s); run `make -C spparks_b200/csrc` or "
                           "__graft_entry__.build()" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    # tests/host_shim builds the library's sources for the host against a CUDA emulator (test
    # infrastructure: it lets the CPU test tier run the kernels).  It is not a way to run the product:
    # only a test that says so may load it.
    if hasattr(L, "spk_cuda_emulated") and os.environ.get("SPK_CUDA_ALLOW_EMULATOR") != "1":
        raise SpkCudaError("%s is the emulated test build of the library; the product needs the CUDA build "
                           "(spparks_b200/lib/libspparks_cuda.so) and a GPU" % LIB_PATH)
    vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    dp = C.POINTER(C.c_double)
    L.spk_cuda_last_error.restype = C.c_char_p
    L.spk_cuda_device_count.argtypes = [C.POINTER(C.c_int)]
    L.spk_cuda_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
    L.spk_cuda_destroy.argtypes = [vp]
    L.spk_cuda_nlocal.restype = C.c_int64
    L.spk_cuda_nlocal.argtypes = [vp]
    L.spk_cuda_nglobal.restype = C.c_int64
    L.spk_cuda_nglobal.argtypes = [vp]
    L.spk_cuda_upload_spins.argtypes = [vp, vp, C.c_int64]
    L.spk_cuda_download_spins.argtypes = [vp, vp, C.c_int64]
    L.spk_cuda_set_temperature.argtypes = [vp, C.c_double]
    L.spk_cuda_set_seed.argtypes = [vp, C.c_uint64]
    L.spk_cuda_validate_spins.argtypes = [vp]
    L.spk_cuda_accept_thresholds.argtypes = [vp, C.POINTER(C.c_uint32), C.c_int]
    L.spk_cuda_randomize.argtypes = [vp, C.c_uint64, C.c_int32, C.c_int32]
    L.spk_cuda_pin_random.argtypes = [vp, C.c_uint64, C.c_double]
    L.spk_cuda_pin_create.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_int32, i64p]
    L.spk_cuda_sweep_color.argtypes = [vp, C.c_uint32, C.c_int, i32p, C.c_int]
    L.spk_cuda_sweep_color_injected.argtypes = [vp, C.c_int, i32p, dp, dp]
    L.spk_cuda_set_siteseeds.argtypes = [vp, vp, C.c_int64]
    L.spk_cuda_get_siteseeds.argtypes = [vp, vp, C.c_int64]
    L.spk_cuda_sweep_color_strict.argtypes = [vp, C.c_int, i32p, C.c_int]
    L.spk_cuda_sweeps.argtypes = [vp, C.c_uint32, C.c_int]
    L.spk_cuda_counters.argtypes = [vp, i64p, i64p]
    L.spk_cuda_reset_counters.argtypes = [vp]
    L.spk_cuda_energy.argtypes = [vp, dp]
    L.spk_cuda_comm_id.argtypes = [C.c_char_p]
    L.spk_cuda_comm_init.argtypes = [vp, C.c_char_p]
    L.spk_cuda_halo_exchange.argtypes = [vp]
    L.spk_cuda_halo_mode.argtypes = [vp]
    L.spk_cuda_cluster_stats.argtypes = [vp, i64p, dp, dp]
    L.spk_cuda_cluster_ids.argtypes = [vp, i32p, i64p, C.c_int64]
    L.spk_cuda_neigh_variant.argtypes = [vp, C.POINTER(C.c_int), dp]
    L.spk_cuda_sync.argtypes = [vp]
    L.spk_cuda_timer_start.argtypes = [vp]
    L.spk_cuda_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.spk_cuda_mask_enable.argtypes = [vp, C.c_int]
    L.spk_cuda_mask_clear_border.argtypes = [vp, i32p, C.c_int]
    L.spk_cuda_download_mask.argtypes = [vp, vp, C.c_int64]
    L.spk_cuda_upload_mask.argtypes = [vp, vp, C.c_int64]
    L.spk_cuda_host_register.argtypes = [vp, C.c_size_t]
    L.spk_cuda_host_unregister.argtypes = [vp]
    L.spk_cuda_set_site_range.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
    L.spk_cuda_site_energy.argtypes = [vp, vp, C.c_int64]
    L.spk_cuda_sweep_path.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.spk_cuda_digest.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.spk_cuda_timer_lap.argtypes = [vp]
    L.spk_cuda_timer_laps.argtypes = [vp, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
    L.spk_cuda_launch_count.argtypes = [vp, i64p]
    L.spk_cuda_device_spins.argtypes = [vp, C.POINTER(vp)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SpkCudaError(load().spk_cuda_last_error().decode())
