/* TEST INFRASTRUCTURE — see spk_oracle.h.  CPU restatement of the reference hot path
 * (SPPARKS "08 Oct 2025").  Citations are relative to /root/reference/.
 * Never linked into or called from the product path. */
#include "spk_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXNEIGH 26
#define ORC_MAXBASIS 2

struct orc_ctx {
  orc_cfg cfg;
  int nbasis, maxneigh, ncolors, dim;
  /* neighbour offset map [basis][slot] -> (di,dj,dk,m'), in the enumeration order of
   * CreateSites::offsets_2d/_3d (src/create_sites.cpp:1303-1371) */
  int cmap[ORC_MAXBASIS][ORC_MAXNEIGH][4];
  double basis[ORC_MAXBASIS][3];
  double xlat, ylat, zlat;
};

/* ---------------------------------------------------------------- lattice tables */

int orc_nbasis(int lattice) { return lattice == ORC_TRI ? 2 : 1; } /* src/lattice.cpp:98-104 */

int orc_maxneigh(int lattice) { /* src/create_sites.cpp:535-540 */
  switch (lattice) {
    case ORC_SQ_4N: return 4;
    case ORC_SQ_8N: return 8;
    case ORC_TRI: return 6;
    case ORC_SC_6N: return 6;
    case ORC_SC_26N: return 26;
  }
  return 0;
}

int orc_ncolors(int lattice) { /* src/lattice.cpp:187-228 with delcolor = 1 */
  switch (lattice) {
    case ORC_SQ_4N: return 2;
    case ORC_SQ_8N: return 4;
    case ORC_TRI: return 4;
    case ORC_SC_6N: return 2;
    case ORC_SC_26N: return 8;
  }
  return 0;
}

static int lattice_dim(int lattice) { return (lattice == ORC_SC_6N || lattice == ORC_SC_26N) ? 3 : 2; }

/* Build the neighbour offset map by the reference's rule: scan cell offsets -1..1 with x
 * outermost, then y, (then z,) basis innermost; keep images whose distance from the
 * central basis site lies in [cutlo,cuthi] (src/create_sites.cpp:1252-1262, 1303-1371). */
static void build_cmap(orc_ctx *c) {
  const double eps = 1.0e-6;
  double cutlo = 1.0, cuthi = 1.0;
  if (c->cfg.lattice == ORC_SQ_8N) cuthi = sqrt(2.0);
  if (c->cfg.lattice == ORC_SC_26N) cuthi = sqrt(3.0);
  int k0 = c->dim == 3 ? -1 : 0, k1 = c->dim == 3 ? 1 : 0;
  for (int b = 0; b < c->nbasis; b++) {
    int n = 0;
    double x0 = c->basis[b][0] * c->xlat, y0 = c->basis[b][1] * c->ylat, z0 = c->basis[b][2] * c->zlat;
    for (int i = -1; i <= 1; i++)
      for (int j = -1; j <= 1; j++)
        for (int k = k0; k <= k1; k++)
          for (int m = 0; m < c->nbasis; m++) {
            double dx = (i + c->basis[m][0]) * c->xlat - x0;
            double dy = (j + c->basis[m][1]) * c->ylat - y0;
            double dz = (k + c->basis[m][2]) * c->zlat - z0;
            double r = sqrt(dx * dx + dy * dy + dz * dz);
            if (r > cutlo - eps && r < cuthi + eps && n < ORC_MAXNEIGH) {
              c->cmap[b][n][0] = i; c->cmap[b][n][1] = j; c->cmap[b][n][2] = k; c->cmap[b][n][3] = m;
              n++;
            }
          }
  }
}

orc_ctx *orc_create(const orc_cfg *cfg) {
  orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
  c->cfg = *cfg;
  c->nbasis = orc_nbasis(cfg->lattice);
  c->maxneigh = orc_maxneigh(cfg->lattice);
  c->ncolors = orc_ncolors(cfg->lattice);
  c->dim = lattice_dim(cfg->lattice);
  if (c->dim == 2) c->cfg.nz = 1;
  /* lattice constant 1.0; tri has a 2-site rectangular cell 1 x sqrt(3) (src/lattice.cpp:98-104,153) */
  c->xlat = 1.0; c->ylat = cfg->lattice == ORC_TRI ? sqrt(3.0) : 1.0; c->zlat = 1.0;
  if (cfg->lattice == ORC_TRI) { c->basis[1][0] = 0.5; c->basis[1][1] = 0.5; }
  build_cmap(c);
  return c;
}

void orc_destroy(orc_ctx *c) { free(c); }

int64_t orc_nsites(const orc_ctx *c) {
  return (int64_t)c->cfg.nx * c->cfg.ny * c->cfg.nz * c->nbasis;
}

static void site_ijkm(const orc_ctx *c, int64_t s, int *i, int *j, int *k, int *m) {
  /* id-1 = ((k*ny + j)*nx + i)*nbasis + m   (src/create_sites.cpp:344-350) */
  *m = (int)(s % c->nbasis); s /= c->nbasis;
  *i = (int)(s % c->cfg.nx); s /= c->cfg.nx;
  *j = (int)(s % c->cfg.ny); *k = (int)(s / c->cfg.ny);
}

void orc_neighbors(const orc_ctx *c, int64_t site, int64_t *out) {
  /* periodic remap of cell indices, then the id formula (src/create_sites.cpp:566-640) */
  int i, j, k, m;
  site_ijkm(c, site, &i, &j, &k, &m);
  int nx = c->cfg.nx, ny = c->cfg.ny, nz = c->cfg.nz;
  for (int n = 0; n < c->maxneigh; n++) {
    int in = i + c->cmap[m][n][0], jn = j + c->cmap[m][n][1], kn = k + c->cmap[m][n][2];
    if (in < 0) in += nx; if (in >= nx) in -= nx;
    if (jn < 0) jn += ny; if (jn >= ny) jn -= ny;
    if (kn < 0) kn += nz; if (kn >= nz) kn -= nz;
    out[n] = (((int64_t)kn * ny + jn) * nx + in) * c->nbasis + c->cmap[m][n][3];
  }
}

int orc_site_color(const orc_ctx *c, int64_t s) { /* src/lattice.cpp:234-283, delcolor = 1 */
  int64_t nx = c->cfg.nx, ny = c->cfg.ny;
  int64_t i, j, k;
  switch (c->cfg.lattice) {
    case ORC_SQ_4N: i = s % nx; j = s / nx; return (int)((i + j) % 2) + 1;
    case ORC_SQ_8N: i = s % nx; j = s / nx; return (int)(2 * (j % 2) + i % 2) + 1;
    case ORC_TRI: return (int)(s % 4) + 1;
    case ORC_SC_6N:
      i = s % nx; j = (s % (nx * ny)) / nx; k = s / (nx * ny);
      return (int)((i + j + k) % 2) + 1;
    case ORC_SC_26N:
      i = s % nx; j = (s % (nx * ny)) / nx; k = s / (nx * ny);
      return (int)(4 * (k % 2) + 2 * (j % 2) + i % 2) + 1;
  }
  return 0;
}

int orc_site_sector(const orc_ctx *c, int nsector, int64_t site) {
  /* create_set(): compare site coordinates with the sub-domain midpoint
   * (src/app_lattice.cpp:969-984); serial run => sub-domain = box = [0,n*lat) */
  int i, j, k, m;
  site_ijkm(c, site, &i, &j, &k, &m);
  double x = (i + c->basis[m][0]) * c->xlat, y = (j + c->basis[m][1]) * c->ylat;
  double z = c->dim == 3 ? (k + c->basis[m][2]) * c->zlat : 0.0;
  double xmid = 0.5 * (0.0 + c->cfg.nx * c->xlat), ymid = 0.5 * (0.0 + c->cfg.ny * c->ylat);
  double zmid = c->dim == 3 ? 0.5 * (0.0 + c->cfg.nz * c->zlat) : 0.0;
  int iw = x < xmid ? 0 : 1, jw = y < ymid ? 0 : 1, kw = z < zmid ? 0 : 1;
  if (nsector == 2) return iw + 1;
  if (nsector == 4) return 2 * jw + iw + 1;
  return 4 * kw + 2 * jw + iw + 1;
}

void orc_sector_box(const orc_ctx *c, int nsector, int isector, int box[6]) {
  /* the same partition expressed as half-open cell ranges: i < ceil(nx/2) <=> x < xmid
   * for every basis site (checked against orc_site_sector in tests) */
  int nx = c->cfg.nx, ny = c->cfg.ny, nz = c->cfg.nz;
  int hx = (nx + 1) / 2, hy = (ny + 1) / 2, hz = (nz + 1) / 2;
  int s = isector - 1;
  int iw = s & 1, jw = nsector >= 4 ? (s >> 1) & 1 : -1, kw = nsector == 8 ? (s >> 2) & 1 : -1;
  box[0] = iw ? hx : 0; box[1] = iw ? nx : hx;
  box[2] = jw < 0 ? 0 : (jw ? hy : 0); box[3] = jw < 0 ? ny : (jw ? ny : hy);
  box[4] = kw < 0 ? 0 : (kw ? hz : 0); box[5] = kw < 0 ? nz : (kw ? nz : hz);
}

/* ---------------------------------------------------------------- RNGs */

#define PM_IA 16807
#define PM_IM 2147483647
#define PM_AM (1.0 / PM_IM)
#define PM_IQ 127773
#define PM_IR 2836

double orc_park_uniform(orc_park *r) { /* src/random_park.cpp:79-86 (Schrage form) */
  int k = r->seed / PM_IQ;
  r->seed = PM_IA * (r->seed - k * PM_IQ) - PM_IR * k;
  if (r->seed < 0) r->seed += PM_IM;
  return PM_AM * r->seed;
}

void orc_park_init_double(orc_park *r, double rseed) { /* src/random_park.cpp:46-50 */
  r->seed = (int)(rseed * PM_IM);
  if (r->seed == 0) r->seed = 1;
}

void orc_park_reset(orc_park *r, double rseed, int64_t offset, int warmup) {
  /* src/random_park.cpp:60-74 (reset and tagreset are the same arithmetic) */
  r->seed = (int)fmod(rseed * PM_IM + offset, PM_IM);
  if (r->seed < 0) r->seed = -r->seed;
  if (r->seed == 0) r->seed = 1;
  for (int i = 0; i < warmup; i++) orc_park_uniform(r);
}

int orc_park_irandom(orc_park *r, int n) { /* src/random_park.cpp:92-97 */
  int i = (int)(orc_park_uniform(r) * n) + 1;
  return i > n ? n : i;
}

void orc_mars_init(orc_mars *r, int seed) { /* src/random_mars.cpp:43-82 */
  while (seed > 900000000) seed -= 900000000;
  int ij = (seed - 1) / 30082, kl = (seed - 1) - 30082 * ij;
  int i = (ij / 177) % 177 + 2, j = ij % 177 + 2, k = (kl / 169) % 178 + 1, l = kl % 169;
  for (int ii = 1; ii <= 97; ii++) {
    double s = 0.0, t = 0.5;
    for (int jj = 1; jj <= 24; jj++) {
      int m = ((i * j) % 179) * k % 179;
      i = j; j = k; k = m;
      l = (53 * l + 1) % 169;
      if ((l * m) % 64 >= 32) s += t;
      t *= 0.5;
    }
    r->u[ii] = s;
  }
  r->c = 362436.0 / 16777216.0; r->cd = 7654321.0 / 16777216.0; r->cm = 16777213.0 / 16777216.0;
  r->i97 = 97; r->j97 = 33;
  orc_mars_uniform(r);
}

double orc_mars_uniform(orc_mars *r) { /* src/random_mars.cpp:88-105 */
  double uni = r->u[r->i97] - r->u[r->j97];
  if (uni < 0.0) uni += 1.0;
  r->u[r->i97] = uni;
  if (--r->i97 == 0) r->i97 = 97;
  if (--r->j97 == 0) r->j97 = 97;
  r->c -= r->cd;
  if (r->c < 0.0) r->c += r->cm;
  uni -= r->c;
  if (uni < 0.0) uni += 1.0;
  return uni;
}

/* Philox4x32-10 (Salmon et al., SC'11).  Not in the reference: it is the counter-based
 * generator the CUDA path uses in production mode; restated here so that GPU half-sweeps in
 * Philox mode can be checked bit-for-bit.  Known-answer vectors in tests/test_oracle_rng.py. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---------------------------------------------------------------- per-site kernels */

int orc_site_energy(const orc_ctx *c, const int32_t *spin, int64_t site) {
  /* AppIsing::site_energy / AppPotts::site_energy: number of unlike neighbours
   * (src/app_ising.cpp:88-95, src/app_potts.cpp:97-104) */
  int64_t nb[ORC_MAXNEIGH];
  orc_neighbors(c, site, nb);
  int e = 0;
  for (int n = 0; n < c->maxneigh; n++) e += spin[site] != spin[nb[n]];
  return e;
}

double orc_total_energy(const orc_ctx *c, const int32_t *spin) { /* src/diag_energy.cpp:53-66 */
  double e = 0.0;
  int64_t n = orc_nsites(c);
  for (int64_t s = 0; s < n; s++) e += orc_site_energy(c, spin, s);
  return e;
}

/* DiagCluster::analyze_clusters (src/diag_cluster.cpp:286-354): sites in increasing local index;
 * an unvisited site opens cluster ncluster+1 (AppLattice::push_new_site, src/app_lattice.cpp:1170-1178)
 * and a stack flood fill marks every equal-spin neighbour (push_connected_neighbors,
 * src/app_lattice.cpp:1187-1201).  Serial, fully periodic: no ghost merging pass. */
int64_t orc_clusters(const orc_ctx *c, const int32_t *spin, int32_t *cluster_ids, int64_t *volumes,
                     double *mean_volume, double *mean_radius) {
  const int64_t n = orc_nsites(c);
  const int nn = c->maxneigh;
  int64_t *stack = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  int64_t nb[ORC_MAXNEIGH];
  int64_t ncluster = 0;
  memset(cluster_ids, 0, (size_t)n * sizeof(int32_t));
  for (int64_t i = 0; i < n; i++) {
    if (cluster_ids[i] != 0) continue;
    const int32_t id = (int32_t)(ncluster + 1);
    int64_t top = 0, vol = 0;
    stack[top++] = i;
    cluster_ids[i] = id;
    ncluster++;
    while (top) {
      const int64_t ii = stack[--top];
      vol++;
      orc_neighbors(c, ii, nb);
      for (int j = 0; j < nn; j++)
        if (spin[nb[j]] == spin[ii] && cluster_ids[nb[j]] == 0) {
          stack[top++] = nb[j];
          cluster_ids[nb[j]] = id;
        }
    }
    volumes[ncluster - 1] = vol;
  }
  free(stack);
  /* stats (src/diag_cluster.cpp:595-605) */
  double volsum = 0.0, rsum = 0.0;
  const double invdim = 1.0 / c->dim;
  for (int64_t k = 0; k < ncluster; k++) {
    volsum += (double)volumes[k];
    rsum += pow((double)volumes[k], invdim);
  }
  if (mean_volume) *mean_volume = volsum / (double)ncluster;
  if (mean_radius) *mean_radius = rsum / (double)ncluster;
  return ncluster;
}

static int count_unlike(const int32_t *spin, const int64_t *nb, int nn, int value) {
  int e = 0;
  for (int n = 0; n < nn; n++) e += value != spin[nb[n]];
  return e;
}

static int site_event_mask(const orc_ctx *c, int32_t *spin, unsigned char *mask, int64_t site, orc_draw_fn draw, void *state);
int orc_site_event(const orc_ctx *c, int32_t *spin, int64_t site, orc_draw_fn draw, void *state) {
  return site_event_mask(c, spin, NULL, site, draw, state);
}

static int site_event_mask(const orc_ctx *c, int32_t *spin, unsigned char *mask, int64_t site, orc_draw_fn draw, void *state) {
  /* AppIsing (src/app_ising.cpp:104-136), AppPotts (src/app_potts.cpp:112-146),
   * AppPottsNeighOnly (src/app_potts_neighonly.cpp:71-117), AppPottsPin
   * (src/app_potts_pin.cpp:93-131) ::site_event_rejection, masks off */
  const int Q = c->cfg.nspins, nn = c->maxneigh;
  int old = spin[site];
  if (c->cfg.model == ORC_POTTS_PIN && old > Q) return 0; /* pinned: no draw, no event */

  int64_t nb[ORC_MAXNEIGH];
  orc_neighbors(c, site, nb);
  int e0 = count_unlike(spin, nb, nn, old);

  int trial;
  if (c->cfg.model == ORC_ISING) {
    trial = draw(state) < 0.5 ? 1 : 2;
  } else if (c->cfg.model == ORC_ISING_SINGLE) {
    /* AppIsingSingle::site_event_rejection (src/app_ising_single.cpp:39-56): flip, no draw here */
    trial = old == 1 ? 2 : 1;
  } else if (c->cfg.model == ORC_POTTS_NEIGHONLY || c->cfg.model == ORC_POTTS_NEIGH) {
    /* distinct neighbour values other than my own, in order of first appearance */
    int cand[ORC_MAXNEIGH], ncand = 0;
    for (int n = 0; n < nn; n++) {
      int v = spin[nb[n]], seen = (v == old);
      for (int q = 0; q < ncand && !seen; q++) seen = (cand[q] == v);
      if (!seen) cand[ncand++] = v;
    }
    if (c->cfg.model == ORC_POTTS_NEIGH) {
      /* AppPottsNeigh (src/app_potts_neigh.cpp:71-93): the draw is always made; the null bin
       * extends the candidate list to maxneigh entries */
      int pick = (int)(nn * draw(state));
      if (pick >= ncand) return 0;
      trial = cand[pick];
    } else {
      if (ncand == 0) return 0; /* no draw */
      int pick = (int)(ncand * draw(state));
      if (pick >= ncand) pick = ncand - 1;
      trial = cand[pick];
    }
  } else {
    trial = (int)(Q * draw(state)) + 1;
    if (trial > Q) trial = Q;
  }

  int e1 = count_unlike(spin, nb, nn, trial);
  int keep = 1;
  if (e1 > e0) {
    if (c->cfg.temperature == 0.0) keep = 0;
    else {
      double t_inverse = 1.0 / c->cfg.temperature; /* src/app_lattice.cpp:893 */
      if (draw(state) > exp(((double)e0 - (double)e1) * t_inverse)) keep = 0;
    }
  }
  spin[site] = keep ? trial : old;
  /* `sweep ... mask yes` (src/app_potts.cpp:136-146 and the same lines of the other apps): reached
   * only when the event ran to its end — the early returns above (pinned site, no candidate, null
   * bin) leave the masks alone, as in the reference */
  if (mask) {
    if (2 * e0 < nn) mask[site] = 1; /* einitial < 0.5*numneigh[i] */
    if (spin[site] != old)
      for (int n = 0; n < nn; n++) mask[nb[n]] = 0;
  }
  return spin[site] != old;
}

/* ---------------------------------------------------------------- sweeps */

static int in_box(const orc_ctx *c, int64_t site, const int *box) {
  if (!box) return 1;
  int i, j, k, m;
  site_ijkm(c, site, &i, &j, &k, &m);
  return i >= box[0] && i < box[1] && j >= box[2] && j < box[3] && k >= box[4] && k < box[5];
}

typedef struct { const double *u; int n; } inj_state;
static double inj_draw(void *p) { inj_state *s = (inj_state *)p; return s->u[s->n++]; }

int64_t orc_halfsweep_injected(const orc_ctx *c, int32_t *spin, int color, const int *box,
                               const double *u1, const double *u2) {
  /* sweep_nomask_nostrict over one colour set (src/app_lattice.cpp:776-780), the RNG
   * replaced by per-site injected uniforms */
  int64_t n = orc_nsites(c), acc = 0;
  for (int64_t s = 0; s < n; s++) {
    if (orc_site_color(c, s) != color || !in_box(c, s, box)) continue;
    double u[2] = {u1[s], u2[s]};
    inj_state st = {u, 0};
    acc += orc_site_event(c, spin, s, inj_draw, &st);
  }
  return acc;
}

/* AppLattice::boundary_clear_mask (src/app_lattice.cpp:1158-1164) for the set `box`: border =
 * sites of the set with a neighbour outside it (find_border_sites, :1092-1148, one layer) */
void orc_boundary_clear_mask(const orc_ctx *c, unsigned char *mask, const int *box) {
  int64_t n = orc_nsites(c), nb[ORC_MAXNEIGH];
  for (int64_t s = 0; s < n; s++) {
    if (!in_box(c, s, box)) continue;
    orc_neighbors(c, s, nb);
    for (int q = 0; q < c->maxneigh; q++)
      if (!in_box(c, nb[q], box)) { mask[s] = 0; break; }
  }
}

/* sweep_mask_nostrict over one colour set (src/app_lattice.cpp:784-792), injected uniforms */
int64_t orc_halfsweep_injected_mask(const orc_ctx *c, int32_t *spin, unsigned char *mask, int color, const int *box,
                                    const double *u1, const double *u2) {
  int64_t n = orc_nsites(c), acc = 0;
  for (int64_t s = 0; s < n; s++) {
    if (orc_site_color(c, s) != color || !in_box(c, s, box) || mask[s]) continue;
    double u[2] = {u1[s], u2[s]};
    inj_state st = {u, 0};
    acc += site_event_mask(c, spin, mask, s, inj_draw, &st);
  }
  return acc;
}

/* Philox production mode.  One Philox4x32-10 block serves the PAIR of same-colour sites whose
 * cell index i differs only in bit 1 (i and i^2, same j,k,basis): counter = (leader lo, leader hi,
 * sweep, iloop*16+color) with leader = index of the site with bit 1 of i cleared, key = seed.
 * The site with bit 1 clear consumes words (0,1), the other one words (2,3); the first word plays
 * u1 = w/2^32 and the second u2.  (ising: four sites per block, see philox_event.)  With those exactly representable uniforms the reference's
 * double formulas reduce to the integer forms below, which is what the CUDA kernels evaluate:
 *   ising:      u1 < 0.5            <=>  w < 2^31
 *   potts/pin:  (int)(Q*u1)+1        =   ((Q*w) >> 32) + 1          (Q < 2^31)
 *   neighonly:  (int)(nevent*u1)     =   (nevent*w) >> 32
 *   reject:     u2 > p               <=>  w' > floor(p*2^32),  p = exp(-dE/T) from libm on the host */
uint32_t orc_accept_threshold(double temperature, int dE) {
  double p = exp((0.0 - (double)dE) * (1.0 / temperature));
  double t = floor(p * 4294967296.0);
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}

/* the two Philox words that serve `site` in (sweep, colour, iloop) */
static void philox_site_words(const orc_ctx *c, int64_t site, uint64_t seed, uint32_t sweep, int color, int iloop,
                              uint32_t w[2]) {
  int ci, cj, ck, cm;
  site_ijkm(c, site, &ci, &cj, &ck, &cm);
  uint64_t leader;
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, blk[4];
  const int isinglike = c->cfg.model == ORC_ISING || c->cfg.model == ORC_ISING_SINGLE;
  if (isinglike) {
    /* ising needs one bit for the trial spin: one block serves the FOUR same-colour sites whose
     * cell index differs only in bits 1,2; site number (i>>1)&3 takes word w of the block:
     * u1 = top bit of w (trial = 1 + bit), u2 = (w mod 2^31)/2^31, i.e. second word = w << 1 */
    leader = (uint64_t)site - (uint64_t)(ci & 6) * (uint64_t)c->nbasis;
  } else {
    leader = (uint64_t)site - (((ci >> 1) & 1) ? 2u * (uint64_t)c->nbasis : 0u);
  }
  uint32_t ctr[4] = {(uint32_t)leader, (uint32_t)(leader >> 32), sweep, (uint32_t)iloop * 16u + (uint32_t)color};
  orc_philox4x32_10(ctr, key, blk);
  if (isinglike) { /* ising/single: same word, only its low 31 bits (the acceptance uniform) are used */
    w[0] = blk[(ci >> 1) & 3]; w[1] = w[0] << 1;
  } else {
    const int second = (ci >> 1) & 1;
    w[0] = blk[2 * second]; w[1] = blk[2 * second + 1];
  }
}

/* One event from two given 32-bit words, integer form (what the CUDA kernels evaluate) */
static int words_event_mask(const orc_ctx *c, int32_t *spin, unsigned char *mask, int64_t site, uint32_t w0, uint32_t w1);
int orc_site_event_words(const orc_ctx *c, int32_t *spin, int64_t site, uint32_t w0, uint32_t w1) {
  return words_event_mask(c, spin, NULL, site, w0, w1);
}
static int words_event_mask(const orc_ctx *c, int32_t *spin, unsigned char *mask, int64_t site, uint32_t w0, uint32_t w1) {
  const int Q = c->cfg.nspins, nn = c->maxneigh;
  const uint32_t w[2] = {w0, w1};
  int old = spin[site];
  if (c->cfg.model == ORC_POTTS_PIN && old > Q) return 0;
  int64_t nb[ORC_MAXNEIGH];
  orc_neighbors(c, site, nb);
  int e0 = count_unlike(spin, nb, nn, old);
  int trial;
  if (c->cfg.model == ORC_ISING) {
    trial = w[0] < 0x80000000u ? 1 : 2;
  } else if (c->cfg.model == ORC_ISING_SINGLE) {
    trial = old == 1 ? 2 : 1;
  } else if (c->cfg.model == ORC_POTTS_NEIGHONLY || c->cfg.model == ORC_POTTS_NEIGH) {
    int cand[ORC_MAXNEIGH], ncand = 0;
    for (int n = 0; n < nn; n++) {
      int v = spin[nb[n]], seen = (v == old);
      for (int q = 0; q < ncand && !seen; q++) seen = (cand[q] == v);
      if (!seen) cand[ncand++] = v;
    }
    if (c->cfg.model == ORC_POTTS_NEIGH) { /* (int)(maxneigh*u1) = (maxneigh*w) >> 32, null bin beyond ncand */
      const int pick = (int)(((uint64_t)nn * w[0]) >> 32);
      if (pick >= ncand) return 0;
      trial = cand[pick];
    } else {
      if (ncand == 0) return 0;
      trial = cand[(int)(((uint64_t)ncand * w[0]) >> 32)];
    }
  } else {
    trial = (int)(((uint64_t)Q * w[0]) >> 32) + 1;
  }
  int e1 = count_unlike(spin, nb, nn, trial);
  int keep = 1;
  if (e1 > e0) {
    if (c->cfg.temperature == 0.0) keep = 0;
    else if (w[1] > orc_accept_threshold(c->cfg.temperature, e1 - e0)) keep = 0;
  }
  spin[site] = keep ? trial : old;
  if (mask) { /* as in site_event_mask */
    if (2 * e0 < nn) mask[site] = 1;
    if (spin[site] != old)
      for (int n = 0; n < nn; n++) mask[nb[n]] = 0;
  }
  return spin[site] != old;
}


/* The same two words fed as uniforms u = w/2^32 through the reference-pinned DOUBLE formulas of
 * orc_site_event: the first word is the event's first draw, the second word its second one
 * (ising/single draws only the acceptance uniform: the second word).  tests/test_oracle_rng.py
 * checks that this and orc_site_event_words agree, i.e. that the integer reduction stated above
 * orc_accept_threshold is exact. */
typedef struct { double u[2]; int n; } word_state;
static double word_draw(void *p) { word_state *s = (word_state *)p; return s->u[s->n < 1 ? s->n++ : 1]; }
int orc_site_event_words_double(const orc_ctx *c, int32_t *spin, int64_t site, uint32_t w0, uint32_t w1) {
  word_state st = {{(double)w0 / 4294967296.0, (double)w1 / 4294967296.0}, 0};
  if (c->cfg.model == ORC_ISING_SINGLE) st.u[0] = st.u[1];
  return orc_site_event(c, spin, site, word_draw, &st);
}

static int philox_event(const orc_ctx *c, int32_t *spin, int64_t site, uint64_t seed,
                        uint32_t sweep, int color, int iloop) {
  uint32_t w[2];
  philox_site_words(c, site, seed, sweep, color, iloop, w);
  return orc_site_event_words(c, spin, site, w[0], w[1]);
}

/* sweep_mask_nostrict over one colour set in Philox mode */
int64_t orc_halfsweep_philox_mask(const orc_ctx *c, int32_t *spin, unsigned char *mask, uint64_t seed, uint32_t sweep,
                                  int color, int iloop, const int *box) {
  int64_t n = orc_nsites(c), acc = 0;
  for (int64_t s = 0; s < n; s++)
    if (orc_site_color(c, s) == color && in_box(c, s, box) && !mask[s]) {
      uint32_t w[2];
      philox_site_words(c, s, seed, sweep, color, iloop, w);
      acc += words_event_mask(c, spin, mask, s, w[0], w[1]);
    }
  return acc;
}

/* the Philox half-sweep evaluated through the double formulas (parity of the integer reduction) */
int64_t orc_halfsweep_philox_double(const orc_ctx *c, int32_t *spin, uint64_t seed, uint32_t sweep,
                                    int color, int iloop, const int *box) {
  int64_t n = orc_nsites(c), acc = 0;
  for (int64_t s = 0; s < n; s++)
    if (orc_site_color(c, s) == color && in_box(c, s, box)) {
      uint32_t w[2];
      philox_site_words(c, s, seed, sweep, color, iloop, w);
      acc += orc_site_event_words_double(c, spin, s, w[0], w[1]);
    }
  return acc;
}

int64_t orc_halfsweep_philox(const orc_ctx *c, int32_t *spin, uint64_t seed, uint32_t sweep,
                             int color, int iloop, const int *box) {
  int64_t n = orc_nsites(c), acc = 0;
  for (int64_t s = 0; s < n; s++)
    if (orc_site_color(c, s) == color && in_box(c, s, box))
      acc += philox_event(c, spin, s, seed, sweep, color, iloop);
  return acc;
}

static double park_draw(void *p) { return orc_park_uniform((orc_park *)p); }

void orc_strict_siteseeds(const orc_ctx *c, orc_mars *master, int32_t *siteseeds) {
  /* AppLattice::init, sweep color/strict (src/app_lattice.cpp:287-295) */
  orc_park r;
  orc_park_init_double(&r, orc_mars_uniform(master));
  double seed = orc_mars_uniform(master);
  int64_t n = orc_nsites(c);
  for (int64_t s = 0; s < n; s++) {
    orc_park_reset(&r, seed, s + 1, 100);
    siteseeds[s] = r.seed;
  }
}

int64_t orc_sweeps_strict(const orc_ctx *c, int32_t *spin, int32_t *siteseeds, int nsweeps) {
  /* iterate_rejection with colour sets and sweep_nomask_strict
   * (src/app_lattice.cpp:646-705, 796-805) */
  int64_t n = orc_nsites(c), acc = 0;
  for (int sw = 0; sw < nsweeps; sw++)
    for (int color = 1; color <= c->ncolors; color++)
      for (int64_t s = 0; s < n; s++) {
        if (orc_site_color(c, s) != color) continue;
        orc_park r = {siteseeds[s]};
        acc += orc_site_event(c, spin, s, park_draw, &r);
        siteseeds[s] = r.seed;
      }
  return acc;
}

/* ---------------------------------------------------------------- host-side one-shot pieces */

void orc_set_site_range(const orc_ctx *c, int32_t *spin, int lo, int hi, orc_mars *master) {
  /* `set site range lo hi`, default loop all, no region/fraction/condition
   * (src/set.cpp:582-627): one Park-Miller stream walked over global ids */
  orc_park r;
  orc_park_init_double(&r, orc_mars_uniform(master));
  int range = hi - lo + 1;
  int64_t n = orc_nsites(c);
  for (int64_t s = 0; s < n; s++) spin[s] = orc_park_irandom(&r, range) - 1 + lo;
}

int64_t orc_pin_create(const orc_ctx *c, int32_t *spin, double fraction, int multi, int nthresh,
                       orc_mars *master) {
  /* AppPottsPin::pin_create, serial (src/app_potts_pin.cpp:242-349) */
  const int Q = c->cfg.nspins, nn = c->maxneigh;
  int64_t n = orc_nsites(c);
  int ndesired = (int)(fraction * n);
  orc_park r;
  orc_park_init_double(&r, orc_mars_uniform(master));
  int64_t nb[ORC_MAXNEIGH];
  int64_t npin = 0;
  if (!multi) {
    while (npin < ndesired) {
      int64_t tries = ndesired - npin;
      for (int64_t t = 0; t < tries; t++) {
        int64_t s = orc_park_irandom(&r, (int)n) - 1;
        if (nthresh == 0) spin[s] = Q + 1;
        else if (orc_site_energy(c, spin, s) >= nthresh) spin[s] = Q + 1;
      }
      npin = 0;
      for (int64_t s = 0; s < n; s++) npin += spin[s] > Q;
    }
  } else {
    while (npin < ndesired) {
      int64_t s = orc_park_irandom(&r, (int)n) - 1;
      orc_neighbors(c, s, nb);
      int ok = spin[s] <= Q;
      for (int j = 0; j < nn; j++) ok &= spin[nb[j]] <= Q;
      if (ok && nthresh && orc_site_energy(c, spin, s) < nthresh) ok = 0;
      if (!ok) continue;
      spin[s] = Q + 1;
      for (int j = 0; j < nn; j++) spin[nb[j]] = Q + 1;
      npin = 0;
      for (int64_t q = 0; q < n; q++) npin += spin[q] > Q;
    }
  }
  return npin;
}

static double next_output_time(double tcur, double delta) { /* src/output.cpp:413-420, logfreq 0 */
  double t = ceil(tcur / delta) * delta;
  if (t == tcur) t = tcur + delta;
  return t;
}

int orc_run_sweep_random(const orc_ctx *c, int32_t *spin, orc_mars *master, int sectors,
                         double stats_delta, double run_time, orc_stats_row *rows, int maxrows) {
  /* App::run -> AppLattice::init/setup/iterate_rejection with sweepflag RANDOM, serial
   * (src/app.cpp:86-148, src/app_lattice.cpp:133-442, 632-711).  Sets are the sectors
   * (default 4 in 2-D, 8 in 3-D) or the whole lattice; each pass draws nselect = nlocal(set)
   * sites with ranapp->irandom and applies site_event_rejection with the same stream. */
  int64_t n = orc_nsites(c);
  int nsector = sectors ? (c->dim == 2 ? 4 : 8) : 1;
  int64_t **site2i = (int64_t **)calloc(nsector, sizeof(int64_t *));
  int64_t *nset = (int64_t *)calloc(nsector, sizeof(int64_t));
  for (int q = 0; q < nsector; q++) site2i[q] = (int64_t *)malloc(n * sizeof(int64_t));
  for (int64_t s = 0; s < n; s++) {
    int q = sectors ? orc_site_sector(c, nsector, s) - 1 : 0;
    site2i[q][nset[q]++] = s;
  }

  /* ranapp (src/app_lattice.cpp:274-278) */
  orc_park ranapp;
  orc_park_init_double(&ranapp, orc_mars_uniform(master));
  double rs = orc_mars_uniform(master);
  orc_park_reset(&ranapp, rs, 0, 100);

  /* dt_sweep: ising 1/2 (src/app_ising.cpp:41), potts and potts/pin 1/Q (src/app_potts.cpp:53),
   * neighonly 1/maxneigh (src/app_potts_neighonly.cpp:54); nstop = 1 => dt_rkmc = dt_sweep */
  double dt_sweep = c->cfg.model == ORC_ISING ? 0.5
                    : c->cfg.model == ORC_POTTS_NEIGHONLY ? 1.0 / c->maxneigh : 1.0 / c->cfg.nspins;
  double ntotal = 0.0;
  for (int q = 0; q < nsector; q++) ntotal += (double)(int)(1.0 * nset[q]);
  double time = 0.0, stoptime = run_time;
  double dt_rkmc = ntotal / n * dt_sweep;
  if (dt_rkmc > stoptime - time) dt_rkmc = stoptime - time;

  int nrows = 0, nsweeps = 0;
  int64_t naccept = 0, nattempt = 0;
  int64_t *sitelist = (int64_t *)malloc(n * sizeof(int64_t));
  /* Output::setup: stats at time 0, then next multiple of delta (src/output.cpp:82-140) */
  if (nrows < maxrows) {
    orc_stats_row r0 = {0.0, 0, 0, 0, orc_total_energy(c, spin)};
    rows[nrows++] = r0;
  }
  double stats_time = stats_delta > 0.0 ? next_output_time(time, stats_delta) : stoptime;
  double nextoutput = stats_time < stoptime ? stats_time : stoptime;

  int done = 0;
  while (!done) {
    for (int q = 0; q < nsector; q++) {
      int64_t nsel = (int64_t)(int)(1.0 * nset[q]);
      for (int64_t t = 0; t < nsel; t++) sitelist[t] = site2i[q][orc_park_irandom(&ranapp, (int)nset[q]) - 1];
      for (int64_t t = 0; t < nsel; t++) naccept += orc_site_event(c, spin, sitelist[t], park_draw, &ranapp);
      nattempt += nsel;
    }
    nsweeps++;
    time += dt_rkmc;
    if (time >= stoptime) done = 1;
    if (done || time >= nextoutput) {
      /* Output::compute (src/output.cpp:145-202) */
      if (time >= stats_time || done) {
        if (nrows < maxrows) {
          orc_stats_row r = {time, naccept, nattempt - naccept, nsweeps, orc_total_energy(c, spin)};
          rows[nrows++] = r;
        }
        stats_time = stoptime;
        if (stats_delta > 0.0) stats_time = next_output_time(time, stats_delta);
      }
      nextoutput = stats_time < stoptime ? stats_time : stoptime;
    }
  }
  for (int q = 0; q < nsector; q++) free(site2i[q]);
  free(site2i); free(nset); free(sitelist);
  return nrows;
}
