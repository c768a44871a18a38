// pion_b200/csrc/grid.cuh -- device-resident grid layout shared by all kernels.
//
// The reference stores the grid as a linked list of heap-allocated `cell`
// objects (grid/cell_interface.h:83-121).  Here the same cells live in
// structure-of-arrays form in HBM: one FP64 plane set per variable,
//   A[v][k][j][i],   i fastest,  padded extents (NG + 2*nb) per active axis,
// i.e. the reference's cell-id order (grid/uniform_grid.cpp:449-451) per
// variable.  Rows are pitched to a multiple of 16 doubles and shifted by
// `xoff` so that the first INTERIOR cell of every row starts on a 128-byte
// boundary (coalesced, sector-aligned warp loads for the interior sweep).
#pragma once
#include <cuda_runtime.h>
#include "physics.cuh"

namespace pion {

#define PION_MAXVAR 16
#define PION_MAXTR 4

struct GridD {
  int ndim;
  int NG[3];   // interior cells
  int nb[3];   // ghost depth per axis (0 for unused axes)
  int NGa[3];  // padded extents
  int xoff;    // leading pad of each row (doubles)
  long sy, sz; // element strides of y and z (x stride is 1)
  long vs;     // variable stride
  double dx;
  // coordinate system (constants.h COORD_*: 1 Cartesian, 2 cylindrical (z,R) 2-D, 3 spherical 1-D) and the
  // position of the low edge of the first interior cell along the radial axis
  int coord;
  double r0;
};

__host__ __device__ __forceinline__ long gidx(const GridD& g, int i, int j, int k) {
  return (long)g.xoff + i + g.sy * j + g.sz * k;
}
__host__ __device__ __forceinline__ long axis_stride(const GridD& g, int ax) {
  return (ax == 0) ? 1L : (ax == 1) ? g.sy : g.sz;
}

// radial axis of a curvilinear grid: Rcyl = axis 1 of the 2-D (z,R) grid, Rsph = axis 0 in 1-D
__host__ __device__ __forceinline__ bool radial_axis(const GridD& g, int ax) {
  return (g.coord == 2 && ax == 1) || (g.coord == 3 && ax == 0);
}
// cell-centre radius of padded index q along the radial axis (cell_interface.cpp:506-512)
__host__ __device__ __forceinline__ double cell_R(const GridD& g, int ax, int q) {
  return g.r0 + (2 * (q - g.nb[ax]) + 1) * (0.5 * g.dx);
}
// centre-of-volume radius: cylindrical VectorOps.h:414-418, spherical VectorOps_spherical.h:188-197
__host__ __device__ __forceinline__ double cell_Rcom(const GridD& g, double R) {
  if (g.coord == 2) return R + g.dx * g.dx / 12. / R;
  double d2 = g.dx / R;
  d2 *= d2;
  return R * (1.0 + 0.25 * d2) / (1.0 + d2 / 12.0);
}

// number of non-tracer variables
__host__ __device__ __forceinline__ constexpr int nbase(int eq) { return (eq == EQ_EULER) ? 5 : (eq == EQ_MHD) ? 8 : 9; }

// Load a primitive state in the solver frame of axis `ax` (a1,a2 = the next two
// axes in cyclic order): the permutation of eqns_base::SetDirection done with
// address arithmetic.
template <int EQ>
__device__ __forceinline__ Prim load_prim(const double* __restrict__ A, long idx, long vs, int ax, int a1, int a2) {
  Prim p;
  p.ro = __ldg(A + idx);
  p.pg = __ldg(A + vs + idx);
  p.vn = __ldg(A + (2 + ax) * vs + idx);
  p.vt1 = __ldg(A + (2 + a1) * vs + idx);
  p.vt2 = __ldg(A + (2 + a2) * vs + idx);
  if (EQ != EQ_EULER) {
    p.bn = __ldg(A + (5 + ax) * vs + idx);
    p.bt1 = __ldg(A + (5 + a1) * vs + idx);
    p.bt2 = __ldg(A + (5 + a2) * vs + idx);
  } else {
    p.bn = p.bt1 = p.bt2 = 0.0;
  }
  p.psi = (EQ == EQ_GLM) ? __ldg(A + 8 * vs + idx) : 0.0;
  return p;
}
template <int EQ>
__device__ __forceinline__ void store_prim(double* __restrict__ A, long idx, long vs, const Prim& p) {
  A[idx] = p.ro;
  A[vs + idx] = p.pg;
  A[2 * vs + idx] = p.vn;
  A[3 * vs + idx] = p.vt1;
  A[4 * vs + idx] = p.vt2;
  if (EQ != EQ_EULER) {
    A[5 * vs + idx] = p.bn;
    A[6 * vs + idx] = p.bt1;
    A[7 * vs + idx] = p.bt2;
  }
  if (EQ == EQ_GLM) A[8 * vs + idx] = p.psi;
}

// rotate a solver-frame triple to the next axis: (n,t1,t2) <- (t1,t2,n)
__device__ __forceinline__ void rot3(double& n, double& t1, double& t2) {
  double tmp = n;
  n = t1;
  t1 = t2;
  t2 = tmp;
}

}  // namespace pion
