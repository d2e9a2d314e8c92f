/* libspparks_cuda — C ABI of the B200-native rejection-KMC colour sweep for SPPARKS lattices.
 *
 * This is the drop-in boundary.  The reference has no FFI for this path: the
 * "binding" is a C++ app-style plugin (plugin/app_*_cuda.{h,cpp}, registered through
 * style_app.h exactly like src/app_potts_neighonly.h:18-42) whose init_app()/iterate_rejection()
 * call the functions below instead of running AppLattice::sweep_*() on the host.  Each entry
 * point cites the reference interface it replaces (paths relative to the SPPARKS tree).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero on error, and
 *     spk_cuda_last_error() returns the message (the plugin forwards it to error->one(),
 *     src/error.cpp:105-115).
 *   - host pointers are only read/written during the call; device memory belongs to the handle.
 *   - sites are in reference order: id-1 = ((k*ny + j)*nx + i)*nbasis + m
 *     (src/create_sites.cpp:344-350); colours are 1-based as Lattice::id2color returns them
 *     (src/lattice.cpp:234-283).
 *   - a slab handle (nranks > 1) owns the cells  [slab_lo, slab_hi)  of the LAST dimension
 *     (z in 3-D = `processors 1 1 N`, y in 2-D = `processors 1 N 1`, src/domain.cpp:204-328) and
 *     one ghost plane on each side (delpropensity = 1, src/app_potts.cpp:35).
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef SPPARKS_CUDA_H
#define SPPARKS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPK_CUDA_ABI_VERSION 1

/* app_style on the path: AppIsing (src/app_ising.cpp:104-136), AppPotts (src/app_potts.cpp:112-146),
 * AppPottsNeighOnly (src/app_potts_neighonly.cpp:71-117), AppPottsPin (src/app_potts_pin.cpp:93-131) */
enum { SPK_ISING = 0, SPK_POTTS = 1, SPK_POTTS_NEIGHONLY = 2, SPK_POTTS_PIN = 3,
       /* variants on the same template: ising/single flips the spin and draws
        * only for the acceptance (src/app_ising_single.cpp:39-56); potts/neigh picks among maxneigh
        * bins of which only the distinct neighbour values are events (src/app_potts_neigh.cpp:71-93) */
       SPK_ISING_SINGLE = 4, SPK_POTTS_NEIGH = 5 };
/* lattice styles with the reference's enum values (src/lattice.cpp:32-33) */
enum { SPK_SQ_4N = 2, SPK_SQ_8N = 3, SPK_TRI = 4, SPK_SC_6N = 5, SPK_SC_26N = 6 };

typedef struct {
  int32_t model;        /* SPK_ISING ... */
  int32_t lattice;      /* SPK_SQ_4N ... */
  int32_t nx, ny, nz;   /* GLOBAL extents in unit cells, even (tri: only nx must be; src/lattice.cpp:187-228); nz = 1 in 2-D */
  int32_t nspins;       /* Q of `app_style potts Q`; ignored for ising */
  double temperature;   /* `temperature` command (src/app_lattice.cpp:889-894) */
  uint64_t seed;        /* key of the counter-based Philox stream (replaces ranapp, src/app_lattice.cpp:274-278) */
  int32_t device;       /* CUDA device ordinal */
  int32_t rank, nranks; /* slab index / number of slabs along the last dimension; 0,1 = whole lattice */
} spk_cuda_desc;

typedef struct spk_cuda_handle spk_cuda_handle;

const char *spk_cuda_last_error(void);
int spk_cuda_abi_version(void);
int spk_cuda_device_count(int *count);

/* lifetime: AppLattice::init()/init_app() (src/app_lattice.cpp:133-327) creates, ~App destroys */
int spk_cuda_create(const spk_cuda_desc *desc, spk_cuda_handle **out);
int spk_cuda_destroy(spk_cuda_handle *h);

/* geometry queries (Lattice::ncolors src/lattice.cpp:187-228; CreateSites maxneigh src/create_sites.cpp:535-540) */
int spk_cuda_ncolors(int lattice);
int spk_cuda_maxneigh(int lattice);
int spk_cuda_nbasis(int lattice);
int64_t spk_cuda_nlocal(const spk_cuda_handle *h);  /* owned sites of this slab */
int64_t spk_cuda_nglobal(const spk_cuda_handle *h);

/* spin = iarray[0] (src/app_potts.cpp:68-71): host <-> device, n = nlocal owned sites */
int spk_cuda_upload_spins(spk_cuda_handle *h, const int32_t *host, int64_t n);
int spk_cuda_download_spins(spk_cuda_handle *h, int32_t *host, int64_t n);
/* `temperature` (src/app_lattice.cpp:889-894): rebuilds the exp() acceptance tables on the host */
int spk_cuda_set_temperature(spk_cuda_handle *h, double temperature);
/* init_app's check of the site values (src/app_potts.cpp:83-90, src/app_ising.cpp:74-81,
 * src/app_potts_pin.cpp:78-84) on the device lattice: error if a spin lies outside 1..Q (Q+1: pinned) */
int spk_cuda_validate_spins(spk_cuda_handle *h);
/* key of the counter-based random stream of the Philox-mode sweeps (spk_cuda_desc::seed), changed after
 * creation: a lattice created before the point where a script's `seed`-derived draws are due (the
 * plugin's cuda_create_sites) gets the seed the host path would have used */
int spk_cuda_set_seed(spk_cuda_handle *h, uint64_t seed);
/* the integer acceptance thresholds the Philox-mode kernels compare the second word with:
 * thr[dE] = floor(exp(-dE/T) * 2^32) for dE = 0 .. n-1 (n <= 27), from the host's libm
 * (exp((einitial-efinal)*t_inverse), src/app_potts.cpp:130); all 0 beyond dE = 0 at T = 0.  For tests. */
int spk_cuda_accept_thresholds(const spk_cuda_handle *h, uint32_t *thr, int n);

/* `set site range lo hi` ON THE DEVICE, bit for bit the reference's values (src/set.cpp:582-627,
 * default loop all / no region, fraction or condition): the one Park-Miller stream that Set walks
 * over the global site ids is evaluated per site with a modular power.  park_seed is the state
 * of `new RandomPark(ranmaster->uniform())` (src/set.cpp:598: (int)(u*(2^31-1)), 0 -> 1). */
int spk_cuda_set_site_range(spk_cuda_handle *h, int32_t park_seed, int32_t lo, int32_t hi);
/* site_energy(i) (src/app_potts.cpp:97-104) of every owned site, one byte each, for
 * `dump ... energy` (DumpText::pack_energy, src/dump_text.cpp:670) without the host's 2*maxneigh
 * gathers per site */
int spk_cuda_site_energy(spk_cuda_handle *h, unsigned char *host, int64_t n);
/* device-side synthetic initialisation for lattices too large for the host path
 * (distribution of `set site range lo hi`, src/set.cpp:582-627, and of `pin f 0 0`,
 * src/app_potts_pin.cpp:262-283; NOT the reference's random stream) */
int spk_cuda_randomize(spk_cuda_handle *h, uint64_t seed, int32_t lo, int32_t hi);
int spk_cuda_pin_random(spk_cuda_handle *h, uint64_t seed, double fraction);
/* `pin fraction multi nthresh` of app_style potts/pin ON THE DEVICE, bit for bit the reference's
 * pinned sites (AppPottsPin::pin_create, src/app_potts_pin.cpp:242-349): park_seed is the state of
 * `new RandomPark(ranmaster->uniform())` (:248), as for spk_cuda_set_site_range.  multi = 0, nthresh = 0
 * (pins anywhere; the attempts of a round commute) runs in parallel and on any number of ranks
 * (collective; after spk_cuda_comm_init); multi = 1 or nthresh > 0 depend on the order of the attempts and
 * are walked by one device thread in the reference's order (one GPU, one-shot).  *npinned = pinned sites
 * of the lattice afterwards.  Where the reference would loop forever (no site can be pinned any more) an
 * error is returned instead.  Lattices beyond 2^31 sites: the reference's 32-bit tagint cannot address
 * them; the same formula is evaluated in 64 bits. */
int spk_cuda_pin_create(spk_cuda_handle *h, double fraction, int multi, int nthresh, int32_t park_seed, int64_t *npinned);

/* One colour set = one kernel: replaces (this->*sweep)(set.nlocal,set.site2i) x nloop
 * (src/app_lattice.cpp:674-691) for the set {sites of `color` inside `box`}.
 * box = {ilo,ihi,jlo,jhi,klo,khi} half-open GLOBAL cell ranges = the sector of create_set()
 * (src/app_lattice.cpp:969-984), or NULL for the whole lattice.  Randoms: Philox4x32-10 keyed
 * on (seed; global site, sweep, colour, iloop).  Ghost planes must be current (spk_cuda_halo_exchange). */
int spk_cuda_sweep_color(spk_cuda_handle *h, uint32_t sweep, int color, const int32_t *box, int nloop);

/* Parity mode: the same set with the two uniforms of site_event_rejection INJECTED per owned
 * site (host arrays of nlocal doubles: u1 = first random->uniform(), u2 = second).  Uses the
 * reference's double-precision formulas and a host-computed exp() table. */
int spk_cuda_sweep_color_injected(spk_cuda_handle *h, int color, const int32_t *box,
                                  const double *u1, const double *u2);

/* Parity mode `sweep color/strict`: per-site Park-Miller streams, sweep_nomask_strict
 * (src/app_lattice.cpp:287-295, 796-805; RandomPark::uniform src/random_park.cpp:79-86) */
int spk_cuda_set_siteseeds(spk_cuda_handle *h, const int32_t *seeds, int64_t n);
int spk_cuda_get_siteseeds(spk_cuda_handle *h, int32_t *seeds, int64_t n);
int spk_cuda_sweep_color_strict(spk_cuda_handle *h, int color, const int32_t *box, int nloop);

/* `sweep ... mask yes` (src/app_lattice.cpp:258-267, 784-792, 809-819): one mask byte per site on
 * the device, zeroed when enabled.  While enabled, every sweep entry point skips masked sites and
 * the events update the masks as the reference's do (src/app_potts.cpp:136-146: mask the site if
 * einitial < numneigh/2, unmask all neighbours of a changed site); the sweeps then run through
 * the generic kernel.  mask_clear_border is AppLattice::boundary_clear_mask (:655, :1158-1164) for
 * the loop set the caller is about to sweep: color = 0 -> the set is the sector `box` (NULL =
 * whole slab), its border sites are unmasked; color = 1..ncolors -> the set is that colour class
 * inside `box`, all of it border.  download_mask copies the owned sites' mask bytes (tests). */
int spk_cuda_mask_enable(spk_cuda_handle *h, int on);
int spk_cuda_mask_clear_border(spk_cuda_handle *h, const int32_t *box, int color);
int spk_cuda_download_mask(spk_cuda_handle *h, unsigned char *host, int64_t n);
/* the way back: the reference keeps its mask array from run to run (created and zeroed once,
 * src/app_lattice.cpp:258-267), so a plugin that builds a new device lattice per run carries the bytes over */
int spk_cuda_upload_mask(spk_cuda_handle *h, const unsigned char *host, int64_t n);

/* nsweeps full sweeps of iterate_rejection with `sweep color`, `sector no`
 * (src/app_lattice.cpp:646-705): for each sweep, for each colour 1..ncolors: ghost refresh
 * (comm->all(), src/app_lattice.cpp:650-651) then the colour kernel.  Sweep numbers
 * first_sweep .. first_sweep+nsweeps-1 key the random stream.  Bit-identical to calling
 * spk_cuda_halo_exchange + spk_cuda_sweep_color per colour, but free to fuse colours. */
int spk_cuda_sweeps(spk_cuda_handle *h, uint32_t first_sweep, int nsweeps);

/* AppLattice::stats (src/app_lattice.cpp:917-934): cumulative accepted / attempted events of
 * this slab since create or the last reset (the caller sums over ranks, as MPI_Allreduce does) */
int spk_cuda_counters(spk_cuda_handle *h, int64_t *naccept, int64_t *nattempt);
int spk_cuda_reset_counters(spk_cuda_handle *h);

/* Which colour kernel spk_cuda_sweeps launches for this handle, how many launches of it make one
 * sweep, and the width of the lattice copy it sweeps on (4 = the int32 lattice, 2 = the 16-bit
 * shadow of sc/26n potts / potts/pin, 0 = the 1-bit shadow of ising on sq/4n; the narrower copies
 * are converted from / to the int32 lattice lazily at the boundaries of spk_cuda_sweeps).
 * bench.py reports it next to the roofline so that narrower device types are visible. */
int spk_cuda_sweep_path(spk_cuda_handle *h, char *name, int capacity, int *launches_per_sweep, int *bytes_per_spin);

/* Candidate-list apps (potts/neighonly, potts/neigh) on sc/26n: which of the two equivalent
 * first-occurrence algorithms (src/app_potts_neighonly.cpp:81-88) the colour kernels currently use:
 * 0 = all-pairs chain (cost independent of the spins), 1 = candidate elimination (cost
 * proportional to the distinct neighbour values; the production path once grains have formed),
 * -1 = not applicable to this handle.  mean_distinct (may be NULL) receives the last mean number
 * of candidates per attempted site the kernels reported, or -1.  The library switches by itself
 * (SPK_CUDA_NEIGH_MODE=chain|adaptive|auto, default auto); results are bit-identical either way. */
int spk_cuda_neigh_variant(spk_cuda_handle *h, int *variant, double *mean_distinct);

/* order-independent 64-bit digest of the owned spins keyed on the GLOBAL site index (the slabs'
 * digests add up, mod 2^64, to the whole lattice's): compares runs at sizes where no host copy
 * is wanted (tests/test_full_size.py: tuned kernels against the reference-pinned generic kernel) */
int spk_cuda_digest(spk_cuda_handle *h, uint64_t *digest);

/* diag_style energy (src/diag_energy.cpp:53-66): sum of site_energy over owned sites */
int spk_cuda_energy(spk_cuda_handle *h, double *energy);

/* diag_style cluster (src/diag_cluster.cpp:230-366, 585-612; AppLattice::push_connected_neighbors,
 * src/app_lattice.cpp:1187-1201): connected components of equal-spin sites over the lattice's
 * neighbour stencil, computed on the device.  ncluster, mean_volume (<N>) and mean_radius (<R> =
 * mean of volume^(1/dimension)) are the three columns DiagCluster::stats prints
 * (src/diag_cluster.cpp:894-897).  Fewer than 2^32 sites per GPU.  With several slabs (after
 * spk_cuda_comm_init; collective: every rank calls it) the per-slab clusters are merged across the
 * slab boundaries with 64-bit global labels (src/diag_cluster.cpp:259-285, 372-396) and every rank
 * receives the statistics of the whole lattice. */
int spk_cuda_cluster_stats(spk_cuda_handle *h, int64_t *ncluster, double *mean_volume, double *mean_radius);
/* the per-site cluster ids (1-based, the reference's numbering: clusters in the order of their
 * lowest site index, as the scan of analyze_clusters opens them) and the volume of every cluster
 * in that order; ids has room for nlocal entries, volumes for nvolumes entries (>= ncluster);
 * either pointer may be NULL.  Runs the analysis itself. */
int spk_cuda_cluster_ids(spk_cuda_handle *h, int32_t *ids, int64_t *volumes, int64_t nvolumes);

/* Multi-GPU calls are COLLECTIVE, like their MPI counterparts in the reference: every rank makes
 * the same calls in the same order.  That includes the calls that replace or expose the spins
 * (upload_spins, randomize, set_site_range, pin_random, device_spins, sweep_color*): they mark the
 * ghost layers stale, and spk_cuda_sweeps re-seeds them with one exchange in which all ranks take
 * part; and the calls that read the ghost layers (energy, site_energy, cluster_stats), which end
 * with a handshake so that no neighbour's next sweep can overtake a slower rank's read.
 * spk_cuda_destroy of a handle that took part in spk_cuda_comm_init is collective too: nobody
 * frees lattice memory that a neighbour still has mapped.
 *
 * multi-GPU: one process per GPU; rank 0 creates the id, the host broadcasts it (MPI_Bcast in
 * the plugin, torch.distributed in bench.py), every rank joins.  Replaces CommLattice::init
 * (src/comm_lattice.cpp:74-135). */
#define SPK_CUDA_COMM_ID_BYTES 128
int spk_cuda_comm_id(char id[SPK_CUDA_COMM_ID_BYTES]);
int spk_cuda_comm_init(spk_cuda_handle *h, const char id[SPK_CUDA_COMM_ID_BYTES]);
/* CommLattice::all() -> perform_swap_site (src/comm_lattice.cpp:137-143, 912-946): refresh both
 * ghost planes from the ring neighbours (or from the periodic image when nranks == 1) */
int spk_cuda_halo_exchange(spk_cuda_handle *h);
/* how spk_cuda_sweeps refreshes the ghosts of this handle: 0 = one GPU (periodic images read in
 * place), 1 = NCCL send/recv per colour step, 2 = peer memory (3-D slabs: the boundary launches
 * store into the neighbours' ghost planes over NVLink, no exchange step; needs one process per
 * GPU with CUDA IPC between them, chosen by spk_cuda_comm_init, SPK_CUDA_NO_P2P=1 keeps NCCL) */
int spk_cuda_halo_mode(const spk_cuda_handle *h);

/* stream control and device timing (CUDA events recorded on the handle's own stream) */
int spk_cuda_sync(spk_cuda_handle *h);
int spk_cuda_timer_start(spk_cuda_handle *h);
int spk_cuda_timer_stop(spk_cuda_handle *h, float *milliseconds);
/* lap events between timer_start and timer_stop (bench.py: per-sweep device times of every rank);
 * timer_laps returns the time from the previous lap (or the start) to each lap */
int spk_cuda_timer_lap(spk_cuda_handle *h);
int spk_cuda_timer_laps(spk_cuda_handle *h, float *milliseconds, int capacity, int *nlaps);
/* number of kernels this handle has launched since create (bench.py's gpu_launches) */
int spk_cuda_launch_count(spk_cuda_handle *h, int64_t *launches);
/* page-lock / release a host array that spk_cuda_upload_spins / _download_spins will be called
 * with repeatedly (the plugin does it for iarray[0], re-registering in grow_app when
 * Memory::grow moves it, src/app_potts.cpp:68-71): copies then run at full PCIe speed */
int spk_cuda_host_register(void *ptr, size_t bytes);
int spk_cuda_host_unregister(void *ptr);
/* raw device pointer to the first owned site (row-major, reference order) for zero-copy hosts */
int spk_cuda_device_spins(spk_cuda_handle *h, void **dptr);

#ifdef __cplusplus
}
#endif
#endif
