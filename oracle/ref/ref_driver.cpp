// TEST INFRASTRUCTURE — not part of the product path.
//
// ref_driver: a thin script-driven harness around the UNMODIFIED reference objects
// (oracle/_ref/libspparks_ref.a, built from /root/reference/src by oracle/ref/Makefile).
// It feeds ordinary SPPARKS input lines to Input::one() (src/input.cpp:244) and understands a
// few extra '@' lines that poke the reference's own per-site functions:
//
//   @load <file>              raw int32[nlocal] (site-id order) -> app->iarray[0]
//   @save <file>              app->iarray[0] -> raw int32[nlocal]
//   @neigh <file>             int32 header {nlocal,maxneigh}, then neighbor[i][0..maxneigh) (local idx)
//   @colors <file>            int32[nlocal]: Lattice::id2color(id[i],1)   (src/lattice.cpp:234)
//   @ids <file>               int32[nlocal]: global ids of the owned sites, local order
//   @sets <file>              int32 nset, then per set {nlocal, site2i[...]}  (src/app_lattice.cpp:954)
//   @siteseeds <file>         int32[nlocal] per-site Park-Miller seeds of sweep color/strict
//   @halfsweep <color> <kfile>
//        one colour half-sweep with INJECTED uniforms: for every owned site i, in increasing
//        local index, with id2color(id[i],1)==color, call the reference's own virtual
//        AppLattice::site_event_rejection(i,&rng) (src/app_lattice.h:74) with rng.seed chosen
//        so that the next RandomPark::uniform() (src/random_park.cpp:79-86) returns exactly
//        k_i/(2^31-1) and the following one (16807*k_i mod (2^31-1))/(2^31-1).
//        kfile = raw int32[nlocal], k_i in [1, 2^31-2].
//   @setsweep <kfile>         one pass over the sets in AppLattice::iterate_rejection order
//        (src/app_lattice.cpp:646-699: sectors -> colours -> sites, nloop=1), injected the same way
//   @masksweep <kfile> <maskfile>
//        one pass of iterate_rejection with `sweep color mask yes` (src/app_lattice.cpp:646-699):
//        per loop set the reference's own boundary_clear_mask(iset) (:655, :1158), then
//        sweep_mask_nostrict's rule `if (mask[i]) continue;` (:784-792) around the same injected
//        site_event_rejection calls as @setsweep (the mask updates are inside the reference's
//        events).  Afterwards the mask bytes of the owned sites are written to maskfile (int32 each).
//   @energy                   prints  "energy <sum_i site_energy(i)>"       (src/diag_energy.cpp:53-66)
//   @counters                 prints  "counters <naccept> <nattempt> <nsweeps>"
//   @time <label>             prints wall seconds since the previous @time
//
// Usage: ref_driver <driver-script>        (SPPARKS screen output goes to stdout unless -quiet)

#include "mpi.h"
#include "spparks.h"
#include "input.h"
#include "app.h"
#include "app_lattice.h"
#include "random_park.h"
#include "domain.h"
#include "lattice.h"
#include "error.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace SPPARKS_NS;

namespace {

const long long IM = 2147483647LL;
const long long IA_INV = 1407677000LL;  // 16807^-1 mod (2^31-1)

// legal access to protected members of AppLattice through a derived class
struct Peek : public AppLattice {
  using AppLattice::naccept;
  using AppLattice::nattempt;
  using AppLattice::nsweeps;
  using AppLattice::set;
  using AppLattice::nset;
  using AppLattice::nsector;
  using AppLattice::ncolors;
  using AppLattice::bothflag;
  using AppLattice::siteseeds;
  using AppLattice::Set;
  using AppLattice::Lmask;
  using AppLattice::mask;
  using AppLattice::boundary_clear_mask;
};

template <class T> T &peek(AppLattice *al, T AppLattice::*m) { return al->*m; }

std::vector<int> read_ints(const char *path, size_t n) {
  std::vector<int> v(n);
  FILE *fp = fopen(path, "rb");
  if (!fp) { fprintf(stderr, "ref_driver: cannot open %s\n", path); exit(2); }
  if (fread(v.data(), sizeof(int), n, fp) != n) { fprintf(stderr, "ref_driver: short read %s\n", path); exit(2); }
  fclose(fp);
  return v;
}

void write_ints(const char *path, const int *p, size_t n) {
  FILE *fp = fopen(path, "wb");
  if (!fp) { fprintf(stderr, "ref_driver: cannot write %s\n", path); exit(2); }
  fwrite(p, sizeof(int), n, fp);
  fclose(fp);
}

void inject_event(AppLattice *al, int i, int k) {
  RandomPark rng(1);
  rng.seed = (int)(((long long)k * IA_INV) % IM);
  al->site_event_rejection(i, &rng);
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: ref_driver <driver-script> [-quiet]\n"); return 2; }
  bool quiet = argc > 2 && strcmp(argv[2], "-quiet") == 0;

  MPI_Init(&argc, &argv);
  std::vector<char *> args;
  std::string a0 = "ref_driver", a1 = "-screen", a2 = "none", a3 = "-log", a4 = "none";
  args.push_back(&a0[0]);
  if (quiet) { args.push_back(&a1[0]); args.push_back(&a2[0]); }
  args.push_back(&a3[0]); args.push_back(&a4[0]);
  SPPARKS *spk = new SPPARKS((int)args.size(), args.data(), MPI_COMM_WORLD);

  FILE *fp = fopen(argv[1], "r");
  if (!fp) { fprintf(stderr, "ref_driver: cannot open %s\n", argv[1]); return 2; }

  auto tprev = std::chrono::steady_clock::now();
  char line[4096];
  while (fgets(line, sizeof line, fp)) {
    size_t len = strlen(line);
    while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
    if (line[0] != '@') { spk->input->one(line); continue; }

    char cmd[64] = "", s1[2048] = "", s2[2048] = "";
    sscanf(line, "%63s %2047s %2047s", cmd, s1, s2);
    AppLattice *al = (AppLattice *)spk->app;
    int nlocal = spk->app ? spk->app->nlocal : 0;

    if (!strcmp(cmd, "@load")) {
      std::vector<int> v = read_ints(s1, nlocal);
      for (int i = 0; i < nlocal; i++) al->iarray[0][i] = v[al->id[i] - 1];
    } else if (!strcmp(cmd, "@save")) {
      std::vector<int> v(nlocal);
      for (int i = 0; i < nlocal; i++) v[al->id[i] - 1] = al->iarray[0][i];
      write_ints(s1, v.data(), nlocal);
    } else if (!strcmp(cmd, "@neigh")) {
      std::vector<int> v;
      v.push_back(nlocal); v.push_back(al->maxneigh);
      for (int i = 0; i < nlocal; i++)
        for (int j = 0; j < al->maxneigh; j++) v.push_back(j < al->numneigh[i] ? al->neighbor[i][j] : -1);
      write_ints(s1, v.data(), v.size());
    } else if (!strcmp(cmd, "@colors")) {
      std::vector<int> v(nlocal);
      for (int i = 0; i < nlocal; i++) v[i] = spk->domain->lattice->id2color(al->id[i], 1);
      write_ints(s1, v.data(), nlocal);
    } else if (!strcmp(cmd, "@ids")) {   // int32[nlocal]: global site ids of the owned sites (several ranks: this rank's)
      std::vector<int> v(nlocal);
      for (int i = 0; i < nlocal; i++) v[i] = (int)al->id[i];
      write_ints(s1, v.data(), nlocal);
    } else if (!strcmp(cmd, "@sets")) {
      Peek::Set *set = peek(al, &Peek::set);
      int nset = peek(al, &Peek::nset);
      std::vector<int> v;
      v.push_back(nset);
      v.push_back(peek(al, &Peek::nsector));
      v.push_back(peek(al, &Peek::ncolors));
      for (int s = 0; s < nset; s++) {
        v.push_back(set[s].nlocal);
        for (int m = 0; m < set[s].nlocal; m++) v.push_back(set[s].site2i[m]);
      }
      write_ints(s1, v.data(), v.size());
    } else if (!strcmp(cmd, "@siteseeds")) {
      int *ss = peek(al, &Peek::siteseeds);
      if (!ss) { fprintf(stderr, "ref_driver: no siteseeds (need sweep color/strict + run)\n"); return 2; }
      write_ints(s1, ss, nlocal);
    } else if (!strcmp(cmd, "@halfsweep")) {
      int color = atoi(s1);
      std::vector<int> k = read_ints(s2, nlocal);
      for (int i = 0; i < nlocal; i++)
        if (spk->domain->lattice->id2color(al->id[i], 1) == color) inject_event(al, i, k[al->id[i] - 1]);
    } else if (!strcmp(cmd, "@setsweep")) {
      std::vector<int> k = read_ints(s1, nlocal);
      Peek::Set *set = peek(al, &Peek::set);
      int nset = peek(al, &Peek::nset), nsector = peek(al, &Peek::nsector);
      int ncolors = peek(al, &Peek::ncolors), both = peek(al, &Peek::bothflag);
      int nset_loop = both ? nsector : nset;
      for (int iset = 0; iset < nset_loop; iset++) {
        if (!both) {
          for (int m = 0; m < set[iset].nlocal; m++) {
            int i = set[iset].site2i[m];
            inject_event(al, i, k[al->id[i] - 1]);
          }
        } else {
          for (int ic = 0; ic < ncolors; ic++) {
            int jset = nsector + iset * ncolors + ic;
            for (int m = 0; m < set[jset].nlocal; m++) {
              int i = set[jset].site2i[m];
              inject_event(al, i, k[al->id[i] - 1]);
            }
          }
        }
      }
    } else if (!strcmp(cmd, "@masksweep")) {
      std::vector<int> k = read_ints(s1, nlocal);
      Peek::Set *set = peek(al, &Peek::set);
      int nset = peek(al, &Peek::nset), nsector = peek(al, &Peek::nsector);
      int ncolors = peek(al, &Peek::ncolors), both = peek(al, &Peek::bothflag);
      char *mask = peek(al, &Peek::mask);
      if (!peek(al, &Peek::Lmask) || !mask) { fprintf(stderr, "ref_driver: @masksweep needs sweep ... mask yes + run\n"); return 2; }
      void (AppLattice::*clear)(int) = &Peek::boundary_clear_mask;
      int nset_loop = both ? nsector : nset;
      for (int iset = 0; iset < nset_loop; iset++) {
        (al->*clear)(iset);
        for (int ic = 0; ic < (both ? ncolors : 1); ic++) {
          int jset = both ? nsector + iset * ncolors + ic : iset;
          for (int m = 0; m < set[jset].nlocal; m++) {
            int i = set[jset].site2i[m];
            if (mask[i]) continue;
            inject_event(al, i, k[al->id[i] - 1]);
          }
        }
      }
      std::vector<int> mv(nlocal);
      for (int i = 0; i < nlocal; i++) mv[al->id[i] - 1] = mask[i];
      write_ints(s2, mv.data(), nlocal);
    } else if (!strcmp(cmd, "@energy")) {
      double e = 0.0;
      for (int i = 0; i < nlocal; i++) e += al->site_energy(i);
      printf("energy %.17g\n", e);
    } else if (!strcmp(cmd, "@counters")) {
      printf("counters %lld %lld %d\n", (long long)peek(al, &Peek::naccept),
             (long long)peek(al, &Peek::nattempt), peek(al, &Peek::nsweeps));
    } else if (!strcmp(cmd, "@time")) {
      auto t = std::chrono::steady_clock::now();
      printf("time %s %.6f\n", s1, std::chrono::duration<double>(t - tprev).count());
      tprev = t;
    } else {
      fprintf(stderr, "ref_driver: unknown directive %s\n", cmd);
      return 2;
    }
    fflush(stdout);
  }
  fclose(fp);
  delete spk;
  MPI_Finalize();
  return 0;
}
