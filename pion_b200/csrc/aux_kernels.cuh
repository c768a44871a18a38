// pion_b200/csrc/aux_kernels.cuh -- the small kernels around the stage kernel:
// HLLD pre-processing flags, H-correction eta, the stand-alone state update
// (unfused API path), ghost-cell fills for every boundary type, halo pack/unpack
// for the NCCL exchange, and the CFL min-reduction.
//
// Reference path restated (paths relative to /root/reference/source):
//   spatial_solvers/solver_eqn_base.cpp:353-599  preprocess_data, calc_Hcorrection, set_Hcorrection
//   coord_sys/VectorOps.cpp:282-439              CentralDiff, GradZone, Divergence
//   sim_control/time_integrator.cpp:881-958      grid_update_state_vector
//   boundaries/{periodic,outflow,oneway_out,inflow,reflecting,fixed,double_Mach_ref}_boundaries.cpp
//   boundaries/MCMD_boundaries.cpp:57-236        BC_select_data2send / BC_update_BCMPI
//   sim_control/calc_timestep.cpp:271-333        calc_dynamics_dt
#pragma once
#include "stage_kernel.cuh"

namespace pion {

// ---------------------------------------------------------------------------
// HLLD shock switch: flag = (divV < 0 && sum_axes |dp|/min(p) > 5) per cell
// (solver_eqn_base.cpp:398-412, consumed at solver_eqn_mhd_adi.cpp:167-177).
// Evaluated for every cell that has both neighbours along every active axis;
// the outermost ghost layer only ever feeds ghost-cell dU, so it is left 0.
// ---------------------------------------------------------------------------
__global__ void k_hlld_flags(GridD g, const double* __restrict__ S, unsigned char* __restrict__ flag) {
  const int ex = g.NGa[0] - 2, ey = (g.ndim > 1) ? g.NGa[1] - 2 : 1, ez = (g.ndim > 2) ? g.NGa[2] - 2 : 1;
  const long n = (long)ex * ey * ez;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    int i = (int)(t % ex) + 1;
    int j = (int)((t / ex) % ey) + ((g.ndim > 1) ? 1 : 0);
    int k = (int)(t / ((long)ex * ey)) + ((g.ndim > 2) ? 1 : 0);
    long c = gidx(g, i, j, k);
    double divv = 0.0, gradp = 0.0;
    const double id2 = 1.0 / (2.0 * g.dx);
    for (int ax = 0; ax < g.ndim; ax++) {
      long st = axis_stride(g, ax);
      const double* V = S + (2 + ax) * g.vs;
      if (radial_axis(g, ax)) {  // VectorOps_Cyl::Divergence: d(R v_R)/(R dR) between centres of volume (VectorOps.cpp:948-954)
        const int q = (ax == 0) ? i : j;
        const double rn = cell_Rcom(g, cell_R(g, ax, q - 1)), rp = cell_Rcom(g, cell_R(g, ax, q + 1));
        divv += 2.0 * (rp * __ldg(V + c + st) - rn * __ldg(V + c - st)) / (rp * rp - rn * rn);
      } else {
        divv += (__ldg(V + c + st) - __ldg(V + c - st)) * id2;
      }
      double pp = __ldg(S + g.vs + c + st), pn = __ldg(S + g.vs + c - st);
      gradp += fabs(pp - pn) * fast_rcp(fmin(pp, pn));
    }
    flag[c] = (divv < 0. && gradp > 5.) ? 1 : 0;
  }
}

// 3-D Cartesian variant: a 32 x 8 tile marches in z keeping p and v_z of the planes k-1, k, k+1 in
// registers, so every plane of p, v_x, v_y, v_z is pulled from L2 once per tile (the x / y neighbours
// come from lines the same block has just loaded) instead of three times in the flat sweep above.
__global__ void __launch_bounds__(256) k_hlld_flags_3d(GridD g, const double* __restrict__ S, unsigned char* __restrict__ flag,
                                                       int kchunk) {
  const int ex = g.NGa[0] - 2, ey = g.NGa[1] - 2, ez = g.NGa[2] - 2;
  const int i = blockIdx.x * 32 + (threadIdx.x & 31) + 1, j = blockIdx.y * 8 + (threadIdx.x >> 5) + 1;
  if (i > ex || j > ey) return;
  const int k0 = blockIdx.z * kchunk + 1, k1 = min(k0 + kchunk, ez + 1);
  const double* Pg = S + g.vs;
  const double *Vx = S + 2 * g.vs, *Vy = S + 3 * g.vs, *Vz = S + 4 * g.vs;
  const double id2 = 1.0 / (2.0 * g.dx);
  long c = gidx(g, i, j, k0);
  double p_m = __ldg(Pg + c - g.sz), p_c = __ldg(Pg + c), vz_m = __ldg(Vz + c - g.sz), vz_c = __ldg(Vz + c);
  for (int k = k0; k < k1; k++, c += g.sz) {
    const double p_p = __ldg(Pg + c + g.sz), vz_p = __ldg(Vz + c + g.sz);
    const double pxp = __ldg(Pg + c + 1), pxn = __ldg(Pg + c - 1), pyp = __ldg(Pg + c + g.sy), pyn = __ldg(Pg + c - g.sy);
    // same summation order as the flat kernel: x, y, z
    double divv = (__ldg(Vx + c + 1) - __ldg(Vx + c - 1)) * id2;
    double gradp = fabs(pxp - pxn) * fast_rcp(fmin(pxp, pxn));
    divv += (__ldg(Vy + c + g.sy) - __ldg(Vy + c - g.sy)) * id2;
    gradp += fabs(pyp - pyn) * fast_rcp(fmin(pyp, pyn));
    divv += (vz_p - vz_m) * id2;
    gradp += fabs(p_p - p_m) * fast_rcp(fmin(p_p, p_m));
    flag[c] = (divv < 0. && gradp > 5.) ? 1 : 0;
    p_m = p_c; p_c = p_p; vz_m = vz_c; vz_c = vz_p;
  }
}

// Face form of the HLLD->HLL switch for the TMA sweep kernel: the flux through the LOW face of a cell along
// x / y / z runs HLL when either cell of that face is flagged (solver_eqn_mhd_adi.cpp:167-177), so one
// byte per cell (bit 0: x face, bit 1: y face, bit 2: z face) replaces two flag loads per face.
__global__ void k_hll_face_flags(GridD g, const unsigned char* __restrict__ flag, unsigned char* __restrict__ face) {
  // four cells per thread: rows are pitched to 16 doubles, so every row starts on a 4-byte boundary; the cell
  // flags are 0/1 bytes, so the per-byte ORs can be done on whole words.  Rows j = 0 / planes k = 0 have no
  // lower neighbour and no face anybody needs.
  const unsigned wpr = (unsigned)(g.sy / 4);
  const unsigned nrow = (unsigned)(g.NGa[1] - 1);
  const unsigned n = wpr * nrow * (unsigned)(g.NGa[2] - 1);
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const unsigned q = t % wpr, r = t / wpr;
    const unsigned j = r % nrow + 1, k = r / nrow + 1;
    const long o = (long)g.sy * j + (long)g.sz * k + 4L * q;
    const unsigned w = *reinterpret_cast<const unsigned*>(flag + o);
    const unsigned xw = (w << 8) | flag[o - 1];
    const unsigned yw = *reinterpret_cast<const unsigned*>(flag + o - g.sy);
    const unsigned zw = *reinterpret_cast<const unsigned*>(flag + o - g.sz);
    *reinterpret_cast<unsigned*>(face + o) = (w | xw) | ((w | yw) << 1) | ((w | zw) << 2);
  }
}

// ---------------------------------------------------------------------------
// H-correction eta for the interface on the + side of every cell, per axis
// (calc_Hcorrection / set_Hcorrection, solver_eqn_base.cpp:423-599):
//   eta = 0.5 (|u_R - u_L| + |c_max(R) - c_max(L)|) from the same edge states as
// the flux.  Column-end rule: the first and last cell of a column have zero slope.
// ---------------------------------------------------------------------------
template <int EQ>
__global__ void k_hcorr_eta(GridD g, const double* __restrict__ S, double* __restrict__ eta, int order, double gamma,
                            double tiny2) {
  const long n = (long)g.NGa[0] * g.NGa[1] * g.NGa[2];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    int ijk[3] = {(int)(t % g.NGa[0]), (int)((t / g.NGa[0]) % g.NGa[1]), (int)(t / ((long)g.NGa[0] * g.NGa[1]))};
    long c = gidx(g, ijk[0], ijk[1], ijk[2]);
    for (int ax = 0; ax < g.ndim; ax++) {
      const int a1 = (ax == 2) ? 0 : ax + 1, a2 = (a1 == 2) ? 0 : a1 + 1;
      const long st = axis_stride(g, ax);
      const int q = ijk[ax], nq = g.NGa[ax];
      if (q + 1 >= nq) continue;  // no interface beyond the last cell
      Prim C = load_prim<EQ>(S, c, g.vs, ax, a1, a2);
      Prim P1 = load_prim<EQ>(S, c + st, g.vs, ax, a1, a2);
      Prim eL = C, eR = P1;
      if (order == 2) {
        const bool slopeC = (q >= 1), slopeP = (q + 2 < nq);
        Prim M1 = slopeC ? load_prim<EQ>(S, c - st, g.vs, ax, a1, a2) : C;
        Prim P2 = slopeP ? load_prim<EQ>(S, c + 2 * st, g.vs, ax, a1, a2) : P1;
        if (!radial_axis(g, ax)) {
#define PION_HE(f)                                                                        \
  {                                                                                       \
    double sc = slopeC ? minmod(C.f - M1.f, P1.f - C.f, tiny2) : 0.0;                      \
    double sp = slopeP ? minmod(P1.f - C.f, P2.f - P1.f, tiny2) : 0.0;                     \
    eL.f = C.f + sc * 0.5;                                                                \
    eR.f = P1.f - sp * 0.5;                                                               \
  }
          PION_HE(ro) PION_HE(pg) PION_HE(vn)
          if (EQ != EQ_EULER) { PION_HE(bn) PION_HE(bt1) PION_HE(bt2) }
#undef PION_HE
        } else {
          // curvilinear radial axis: slopes between centres of volume, edge offsets from them
          double Rq[4], Rcm[4];
          for (int w = 0; w < 4; w++) { Rq[w] = cell_R(g, ax, q - 1 + w); Rcm[w] = cell_Rcom(g, Rq[w]); }
          const double i01 = 1.0 / (Rcm[1] - Rcm[0]), i12 = 1.0 / (Rcm[2] - Rcm[1]), i23 = 1.0 / (Rcm[3] - Rcm[2]);
          const double delL = Rq[1] + 0.5 * g.dx - Rcm[1], delR = Rq[2] - 0.5 * g.dx - Rcm[2];
#define PION_HER(f)                                                                                              \
  {                                                                                                              \
    double sc = slopeC ? minmod((C.f - M1.f) * i01, (P1.f - C.f) * i12, PION_VERY_TINY_VALUE) : 0.0;              \
    double sp = slopeP ? minmod((P1.f - C.f) * i12, (P2.f - P1.f) * i23, PION_VERY_TINY_VALUE) : 0.0;             \
    eL.f = C.f + sc * delL;                                                                                      \
    eR.f = P1.f + sp * delR;                                                                                     \
  }
          PION_HER(ro) PION_HER(pg) PION_HER(vn)
          if (EQ != EQ_EULER) { PION_HER(bn) PION_HER(bt1) PION_HER(bt2) }
#undef PION_HER
        }
      }
      double e = 0.5 * (fabs(eR.vn - eL.vn) + fabs(maxspeed<EQ>(eR, gamma) - maxspeed<EQ>(eL, gamma)));
      eta[(long)ax * g.vs + c] = e;
    }
  }
}

// ---------------------------------------------------------------------------
// Stand-alone grid_update_state_vector (unfused API path): Ph = UtoP(PtoU(P)+dU),
// dU = 0, P = Ph on the full step (time_integrator.cpp:881-958).
// ---------------------------------------------------------------------------
struct UpdateArgs {
  GridD g;
  PhysParams pp;
  double* P;
  double* Ph;
  double* dU;
  const unsigned char* mask;
  long long* counters;
  double glm_damp;
  int ntr;
  int full;  // step == ooa
};
template <int EQ>
__global__ void k_update_state(const __grid_constant__ UpdateArgs a) {
  const GridD& g = a.g;
  // the reference loops over ALL cells; ghost cells are !isdomain except periodic
  // ghosts, whose update is overwritten by the next boundary update, so only the
  // interior is advanced and dU is cleared everywhere.
  const long nall = (long)g.NGa[0] * g.NGa[1] * g.NGa[2];
  constexpr int NB = nbase(EQ);
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nall; t += (long)gridDim.x * blockDim.x) {
    int i = (int)(t % g.NGa[0]), j = (int)((t / g.NGa[0]) % g.NGa[1]), k = (int)(t / ((long)g.NGa[0] * g.NGa[1]));
    long c = gidx(g, i, j, k);
    bool interior = (i >= g.nb[0] && i < g.NGa[0] - g.nb[0] && j >= g.nb[1] && j < g.NGa[1] - g.nb[1] &&
                     k >= g.nb[2] && k < g.NGa[2] - g.nb[2]);
    bool domain = interior && (a.mask ? a.mask[c] != 0 : true);
    if (domain) {
      Prim Pb = load_prim<EQ>(a.P, c, g.vs, 0, 1, 2);
      Cons U;
      PtoU<EQ>(Pb, U, a.pp.gamma - 1.0);
      U.rho += a.dU[c]; U.erg += a.dU[g.vs + c]; U.mn += a.dU[2 * g.vs + c]; U.mt1 += a.dU[3 * g.vs + c];
      U.mt2 += a.dU[4 * g.vs + c];
      if (EQ != EQ_EULER) { U.bbn += a.dU[5 * g.vs + c]; U.bbt1 += a.dU[6 * g.vs + c]; U.bbt2 += a.dU[7 * g.vs + c]; }
      if (EQ == EQ_GLM) U.psi += a.dU[8 * g.vs + c];
      Prim Pn;
      int st = UtoP<EQ>(U, Pn, a.pp);
      if (EQ == EQ_GLM) Pn.psi *= a.glm_damp;
      if (a.pp.have_mp && (Pn.pg * a.pp.mu_tot_over_kB / Pn.ro > a.pp.max_temp))
        Pn.pg = Pn.ro * a.pp.max_temp / a.pp.mu_tot_over_kB;
      store_prim<EQ>(a.Ph, c, g.vs, Pn);
      if (a.full) store_prim<EQ>(a.P, c, g.vs, Pn);
      for (int q = 0; q < a.ntr; q++) {
        long o = (long)(NB + q) * g.vs + c;
        double pb = a.P[o];
        if (a.pp.have_mp) pb *= scma_corr(pb);
        double pn = (pb * Pb.ro + a.dU[o]) / U.rho;
        if (a.pp.have_mp) pn *= scma_corr(pn);
        a.Ph[o] = pn;
        if (a.full) a.P[o] = pn;
      }
      if (st && a.counters) {
        if (st & ST_NEG_RHO) atomicAdd((unsigned long long*)&a.counters[0], 1ULL);
        if (st & ST_NEG_PG) atomicAdd((unsigned long long*)&a.counters[1], 1ULL);
      }
    } else if (a.full && interior) {
      for (int v = 0; v < NB + a.ntr; v++) a.P[(long)v * g.vs + c] = a.Ph[(long)v * g.vs + c];
    }
    for (int v = 0; v < NB + a.ntr; v++) a.dU[(long)v * g.vs + c] = 0.0;
  }
}

// ---------------------------------------------------------------------------
// Boundary ghost fill, one launch per face, faces in the order XN,XP,YN,YP,ZN,ZP
// so that edge/corner ghosts inherit from the faces filled before them, exactly as
// the reference's per-boundary cell lists do (grid/uniform_grid.cpp:1009-1216:
// X faces hold interior y,z; Y faces add the x-ghost corners; Z faces whole planes).
// ---------------------------------------------------------------------------
struct BCArgs {
  GridD g;
  double* A[2];  // arrays to fill (Ph and/or P); unused entries null
  int narr;
  int face;      // 0..5
  int type;      // PION_BC_* code
  int nvar;
  int eq;
  int ftr;       // first tracer index
  double refval[PION_MAXVAR];
  double simtime;
  double sim_xmin[3];
};

struct BCRef {
  double v[PION_MAXVAR];
};

__device__ __forceinline__ void bc_face_extents(const GridD& g, int face, int* lo, int* hi) {
  const int ax = face >> 1;
  for (int q = 0; q < 3; q++) { lo[q] = 0; hi[q] = g.NGa[q]; }
  if (ax == 0) {
    for (int q = 1; q < 3; q++) { lo[q] = g.nb[q]; hi[q] = g.NGa[q] - g.nb[q]; }
  } else if (ax == 1) {
    lo[2] = g.nb[2]; hi[2] = g.NGa[2] - g.nb[2];
  }
  if (face & 1) { lo[ax] = g.NGa[ax] - g.nb[ax]; hi[ax] = g.NGa[ax]; }
  else { lo[ax] = 0; hi[ax] = g.nb[ax]; }
}

// ghost cell number t of face `face` (boundary type `type`, reference values `refval`)
__device__ __forceinline__ void bc_fill_cell(const BCArgs& a, int face, int type, const double* refval, long t) {
  const GridD& g = a.g;
  int lo[3], hi[3];
  bc_face_extents(g, face, lo, hi);
  const int ex = hi[0] - lo[0], ey = hi[1] - lo[1];
  const int ax = face >> 1, pos = face & 1;
  const long st = axis_stride(g, ax);
  int ijk[3] = {(int)(t % ex) + lo[0], (int)((t / ex) % ey) + lo[1], (int)(t / ((long)ex * ey)) + lo[2]};
  const long c = gidx(g, ijk[0], ijk[1], ijk[2]);
  // depth of this ghost cell (1 = adjacent to the grid) and its source cells
  const int q = ijk[ax];
  const int edge = pos ? g.NGa[ax] - g.nb[ax] - 1 : g.nb[ax];
  const int depth = pos ? q - edge : edge - q;
  const long c_edge = c + (long)(edge - q) * st;
  const long c_mirror = c_edge + (long)(pos ? -(depth - 1) : (depth - 1)) * st;
  const long c_per = c + (long)(pos ? -g.NG[ax] : g.NG[ax]) * st;
  for (int w = 0; w < a.narr; w++) {
    double* A = a.A[w];
    switch (type) {
      case 1:  // PERIODIC (periodic_boundaries.cpp:68-88)
        for (int v = 0; v < a.nvar; v++) A[(long)v * g.vs + c] = A[(long)v * g.vs + c_per];
        break;
      case 2:   // OUTFLOW (outflow_boundaries.cpp:109-160)
      case 13:  // ONEWAY_OUT (oneway_out_boundaries.cpp:38-115)
        for (int v = 0; v < a.nvar; v++) A[(long)v * g.vs + c] = A[(long)v * g.vs + c_edge];
        if (type == 13) {
          const double sgn = pos ? 1.0 : -1.0;
          const long o = (long)(2 + ax) * g.vs + c;
          A[o] = sgn * fmax(0.0, A[o] * sgn);
        }
        if (a.eq == EQ_GLM) A[8 * g.vs + c] = -A[8 * g.vs + c_mirror];  // GLM_NEGATIVE_BOUNDARY
        break;
      case 4:  // REFLECTING (reflecting_boundaries.cpp:123-145): both layers copy the edge cell
        for (int v = 0; v < a.nvar; v++) A[(long)v * g.vs + c] = A[(long)v * g.vs + c_edge] * refval[v];
        break;
      case 3:  // INFLOW (inflow_boundaries.cpp:83-100)
      case 5:  // FIXED (fixed_boundaries.cpp:91-107)
        for (int v = 0; v < a.nvar; v++) A[(long)v * g.vs + c] = refval[v];
        break;
      case 8: {  // DMACH (double_Mach_ref_boundaries.cpp:169-208)
        const double dxo2 = 0.5 * g.dx;
        const double xpos = a.sim_xmin[0] + (2 * (ijk[0] - g.nb[0]) + 1) * dxo2;
        const double ypos = a.sim_xmin[1] + (2 * (ijk[1] - g.nb[1]) + 1) * dxo2;
        const double bpos = 10.0 * a.simtime / sin(M_PI / 3.0) + 1.0 / 6.0 + ypos / tan(M_PI / 3.0);
        if (xpos <= bpos) {
          A[c] = 8.0; A[g.vs + c] = 116.5; A[2 * g.vs + c] = 7.14470958; A[3 * g.vs + c] = -4.125; A[4 * g.vs + c] = 0.0;
          for (int v = a.ftr; v < a.nvar; v++) A[(long)v * g.vs + c] = 1.0;
        } else {
          for (int v = 0; v < a.nvar; v++) A[(long)v * g.vs + c] = refval[v];
        }
      } break;
      default:
        break;
    }
  }
}

__global__ void k_bc_face(const __grid_constant__ BCArgs a) {
  int lo[3], hi[3];
  bc_face_extents(a.g, a.face, lo, hi);
  const long n = (long)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x)
    bc_fill_cell(a, a.face, a.type, a.refval, t);
}

// Both faces of ONE axis in one launch (they never read each other's ghost cells -- a periodic face copies
// from interior cells): three ghost-fill launches per boundary update instead of six.  The order ACROSS axes
// (x, y, z: edges and corners inherit) stays with the host.  `a.face` is the LOW face; type2 / refval2 describe
// the high face; a type of 0 (or PION_BC_MPI: filled by the halo exchange) skips that face.
__global__ void k_bc_axis(const __grid_constant__ BCArgs a, const int type2, const __grid_constant__ BCRef ref2) {
  int lo[3], hi[3];
  bc_face_extents(a.g, a.face, lo, hi);
  const long n = (long)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);  // same count on both faces
  const bool do_lo = a.type != 0 && a.type != 10, do_hi = type2 != 0 && type2 != 10;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * n; t += (long)gridDim.x * blockDim.x) {
    if (t < n) {
      if (do_lo) bc_fill_cell(a, a.face, a.type, a.refval, t);
    } else if (do_hi) {
      bc_fill_cell(a, a.face + 1, type2, ref2.v, t - n);
    }
  }
}

// DMACH2 internal boundary: fixed post-shock state in the y<0 ghost rows for
// x <= 1/6 (double_Mach_ref_boundaries.cpp:90-150, :214-230).
__global__ void k_bc_dmach2(const __grid_constant__ BCArgs a) {
  const GridD& g = a.g;
  const long n = (long)g.NG[0] * g.nb[1];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    int i = (int)(t % g.NG[0]) + g.nb[0], j = (int)(t / g.NG[0]);
    const double xpos = a.sim_xmin[0] + (2 * (i - g.nb[0]) + 1) * (0.5 * g.dx);
    if (xpos <= 1. / 6.) {
      const long c = gidx(g, i, j, g.nb[2]);
      for (int w = 0; w < a.narr; w++)
        for (int v = 0; v < a.nvar; v++) a.A[w][(long)v * g.vs + c] = a.refval[v];
    }
  }
}

// ---------------------------------------------------------------------------
// Halo pack / unpack for a BCMPI face: nb layers of every variable, same face
// extents as the physical boundaries (so the x -> y -> z exchange order fills
// edges and corners).  Replaces the per-cell MPI_Pack records of
// comms/comm_mpi.cpp:285-420 with a dense [var][k][j][i] slab.
// ---------------------------------------------------------------------------
struct HaloArgs {
  GridD g;
  double* A;     // array exchanged (Ph or P)
  double* buf;   // contiguous slab
  int face;
  int nvar;
  int pack;      // 1: interior layers next to `face` -> buf ; 0: buf -> ghost layers of `face`
};
__global__ void k_halo(const __grid_constant__ HaloArgs a) {
  const GridD& g = a.g;
  int lo[3], hi[3];
  bc_face_extents(g, a.face, lo, hi);
  const int ax = a.face >> 1, pos = a.face & 1;
  if (a.pack) {  // shift the ghost slab inwards by nb cells: the cells the neighbour needs
    const int sh = pos ? -g.nb[ax] : g.nb[ax];
    lo[ax] += sh;
    hi[ax] += sh;
  }
  const int ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
  const long n = (long)ex * ey * ez;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n * a.nvar; t += (long)gridDim.x * blockDim.x) {
    const int v = (int)(t / n);
    const long r = t % n;
    const int i = (int)(r % ex) + lo[0], j = (int)((r / ex) % ey) + lo[1], k = (int)(r / ((long)ex * ey)) + lo[2];
    const long c = (long)v * g.vs + gidx(g, i, j, k);
    if (a.pack) a.buf[t] = a.A[c];
    else a.A[c] = a.buf[t];
  }
}

// both exchanged faces of one axis in one launch (buf2 / face + 1 may be absent: null)
__global__ void k_halo_axis(const __grid_constant__ HaloArgs a, double* const buf2) {
  const GridD& g = a.g;
  const int ax = a.face >> 1;
  for (int s = 0; s < 2; s++) {
    double* const buf = s ? buf2 : a.buf;
    if (!buf) continue;
    int lo[3], hi[3];
    bc_face_extents(g, 2 * ax + s, lo, hi);
    if (a.pack) {
      const int sh = s ? -g.nb[ax] : g.nb[ax];
      lo[ax] += sh;
      hi[ax] += sh;
    }
    const int ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    const long n = (long)ex * ey * ez;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n * a.nvar; t += (long)gridDim.x * blockDim.x) {
      const int v = (int)(t / n);
      const long r = t % n;
      const int i = (int)(r % ex) + lo[0], j = (int)((r / ex) % ey) + lo[1], k = (int)(r / ((long)ex * ey)) + lo[2];
      const long c = (long)v * g.vs + gidx(g, i, j, k);
      if (a.pack) buf[t] = a.A[c];
      else a.A[c] = buf[t];
    }
  }
}

// ---------------------------------------------------------------------------
// calc_dynamics_dt (calc_timestep.cpp:271-333): min over interior cells of
// CellTimeStep(P); warp-shuffle + one atomicMin per block on the ordered bits.
// ---------------------------------------------------------------------------
template <int EQ>
__global__ void k_calc_dt(GridD g, PhysParams pp, const double* __restrict__ P, const unsigned char* __restrict__ tsmask,
                          double cfl, unsigned long long* dtmin) {
  const long ncell = (long)g.NG[0] * g.NG[1] * g.NG[2];
  double my = 1.0e100;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ncell; t += (long)gridDim.x * blockDim.x) {
    int i = (int)(t % g.NG[0]), j = (int)((t / g.NG[0]) % g.NG[1]), k = (int)(t / ((long)g.NG[0] * g.NG[1]));
    long c = gidx(g, i + g.nb[0], j + g.nb[1], k + g.nb[2]);
    if (tsmask && !tsmask[c]) continue;
    Prim p = load_prim<EQ>(P, c, g.vs, 0, 1, 2);
    my = fmin(my, cell_time_step<EQ>(p, pp, g.ndim, g.dx, cfl));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my = fmin(my, __shfl_xor_sync(0xffffffffu, my, o));
  __shared__ double s[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s[w] = my;
  __syncthreads();
  if (w == 0) {
    my = (lane < (int)(blockDim.x >> 5)) ? s[lane] : 1.0e100;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my = fmin(my, __shfl_xor_sync(0xffffffffu, my, o));
    if (lane == 0 && my < 1.0e100) atomicMin(dtmin, dbl_ordered_bits(my));
  }
}

// BC_update_STWIND -> stellar_wind::set_cell_values (grid/stellar_wind_BC.cpp:642-670): the
// wind cells' reference states overwrite P AND Ph on every boundary update.
__global__ void k_wind_set(long vs, int nvar, long nw, const long* __restrict__ idx, const double* __restrict__ val,
                           double* __restrict__ P, double* __restrict__ Ph) {
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nw * nvar; t += (long)gridDim.x * blockDim.x) {
    const int v = (int)(t / nw);
    const long c = idx[t % nw];
    const double x = val[t];
    P[(long)v * vs + c] = x;
    Ph[(long)v * vs + c] = x;
  }
}

// small utility kernels
// upload / download staging: one variable between the host's compact rows (NGa[0] doubles) and the device's
// pitched rows (grid.cuh).  The PCIe copy itself is a FLAT cudaMemcpyAsync of the compact variable (55 GB/s
// measured; the strided cudaMemcpy2D form reached ~41 GB/s) and this kernel does the re-pitching in HBM.
__global__ void k_repack_var(GridD g, double* __restrict__ compact, double* __restrict__ pitched, int to_device) {
  const int nx = g.NGa[0];
  const long rows = (long)g.NGa[1] * g.NGa[2];
  for (long r = blockIdx.x; r < rows; r += gridDim.x) {  // one row per block iteration: no index division
    double* cr = compact + r * nx;
    double* pr = pitched + r * g.sy + g.xoff;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      if (to_device) pr[i] = cr[i];
      else cr[i] = pr[i];
    }
  }
}

__global__ void k_copy_interior(GridD g, const double* __restrict__ src, double* __restrict__ dst, int nvar, int zero_var) {
  const long ncell = (long)g.NG[0] * g.NG[1] * g.NG[2];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ncell; t += (long)gridDim.x * blockDim.x) {
    int i = (int)(t % g.NG[0]), j = (int)((t / g.NG[0]) % g.NG[1]), k = (int)(t / ((long)g.NG[0] * g.NG[1]));
    long c = gidx(g, i + g.nb[0], j + g.nb[1], k + g.nb[2]);
    for (int v = 0; v < nvar; v++) dst[(long)v * g.vs + c] = (v == zero_var) ? 0.0 : src[(long)v * g.vs + c];
  }
}
__global__ void k_zero_var_interior(GridD g, double* __restrict__ A, int var) {
  const long ncell = (long)g.NG[0] * g.NG[1] * g.NG[2];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ncell; t += (long)gridDim.x * blockDim.x) {
    int i = (int)(t % g.NG[0]), j = (int)((t / g.NG[0]) % g.NG[1]), k = (int)(t / ((long)g.NG[0] * g.NG[1]));
    A[(long)var * g.vs + gidx(g, i + g.nb[0], j + g.nb[1], k + g.nb[2])] = 0.0;
  }
}

}  // namespace pion
