// Lattice stencils of the SPPARKS on-lattice apps, as compile-time tables.
//
// Neighbour enumeration order is part of the contract (potts/neighonly picks the n-th DISTINCT
// neighbour value in first-occurrence order, src/app_potts_neighonly.cpp:81-88): it is the order
// in which CreateSites::offsets_2d/_3d (src/create_sites.cpp:1303-1371) scan cell offsets —
// dx outermost, then dy, (then dz,) basis innermost.  tests/test_oracle_golden.py pins the same
// tables against the reference's own neighbor[][] arrays.
#pragma once
#include <stdint.h>
#include "../../include/spparks_cuda.h"

namespace spk {

struct Off { int8_t dx, dy, dz, m; };

template <int LAT> struct Lat;

template <> struct Lat<SPK_SQ_4N> {
  static constexpr int NB = 1, NN = 4, NC = 2, DIM = 2;
  __host__ __device__ static constexpr Off off(int /*m*/, int n) {
    constexpr Off t[4] = {{-1, 0, 0, 0}, {0, -1, 0, 0}, {0, 1, 0, 0}, {1, 0, 0, 0}};
    return t[n];
  }
};

template <> struct Lat<SPK_SQ_8N> {
  static constexpr int NB = 1, NN = 8, NC = 4, DIM = 2;
  __host__ __device__ static constexpr Off off(int /*m*/, int n) {
    constexpr Off t[8] = {{-1, -1, 0, 0}, {-1, 0, 0, 0}, {-1, 1, 0, 0}, {0, -1, 0, 0},
                          {0, 1, 0, 0},   {1, -1, 0, 0}, {1, 0, 0, 0},  {1, 1, 0, 0}};
    return t[n];
  }
};

// tri: rectangular 1 x sqrt(3) cell with basis (0,0) and (1/2,1/2) (src/lattice.cpp:98-104,153)
template <> struct Lat<SPK_TRI> {
  static constexpr int NB = 2, NN = 6, NC = 4, DIM = 2;
  __host__ __device__ static constexpr Off off(int m, int n) {
    constexpr Off t[2][6] = {
        {{-1, -1, 0, 1}, {-1, 0, 0, 0}, {-1, 0, 0, 1}, {0, -1, 0, 1}, {0, 0, 0, 1}, {1, 0, 0, 0}},
        {{-1, 0, 0, 1}, {0, 0, 0, 0}, {0, 1, 0, 0}, {1, 0, 0, 0}, {1, 0, 0, 1}, {1, 1, 0, 0}}};
    return t[m][n];
  }
};

template <> struct Lat<SPK_SC_6N> {
  static constexpr int NB = 1, NN = 6, NC = 2, DIM = 3;
  __host__ __device__ static constexpr Off off(int /*m*/, int n) {
    constexpr Off t[6] = {{-1, 0, 0, 0}, {0, -1, 0, 0}, {0, 0, -1, 0}, {0, 0, 1, 0}, {0, 1, 0, 0}, {1, 0, 0, 0}};
    return t[n];
  }
};

template <> struct Lat<SPK_SC_26N> {
  static constexpr int NB = 1, NN = 26, NC = 8, DIM = 3;
  __host__ __device__ static constexpr Off off(int /*m*/, int n) {
    // all (dx,dy,dz) in {-1,0,1}^3 \ {0}, dx-major, dz-minor
    int q = n < 13 ? n : n + 1;
    return Off{(int8_t)(q / 9 - 1), (int8_t)((q / 3) % 3 - 1), (int8_t)(q % 3 - 1), 0};
  }
};

// Geometry of one slab as the kernels see it.  Local storage is row-major in reference order
// with one ghost plane (3-D) / ghost row (2-D) on each side of the slab dimension:
//   element(i,j,k,m) = spin[((k*Y + j)*X + i)*NB + m],  k in [-1,Z] (3-D)  or  j in [-1,Y] (2-D)
struct Geo {
  int X, Y, Z;        // local extents in cells (Z = 1 in 2-D)
  int gX, gY, gZ;     // global extents
  int g0;             // global index of local 0 along the slab dimension
  int lo_wrap;        // local index holding the image of "-1"  (-1 = ghost, n-1 = periodic)
  int hi_wrap;        // local index holding the image of "n"   ( n = ghost, 0   = periodic)
};

// inclusive-exclusive LOCAL cell ranges of the set being swept (sector box ∩ slab)
struct Box { int ilo, ihi, jlo, jhi, klo, khi; };

}  // namespace spk
