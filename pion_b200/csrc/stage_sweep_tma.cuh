// pion_b200/csrc/stage_sweep_tma.cuh -- the 3-D production stage kernel: the flux-once
// sweep of stage_sweep.cuh with the stencil STAGED IN SHARED MEMORY BY TMA.
//
// Same reference path and the same arithmetic as k_stage_sweep (time_integrator.cpp:498-958,
// VectorOps.cpp:535-644, solver_eqn_base.cpp:152-342, solver_eqn_mhd_adi.cpp:368-443,782-844);
// what changes is where the operands come from:
//
//   * the state planes a block needs live in a ring of FOUR shared-memory plane buffers
//     [var][TY+3 rows][36 columns] (cells i0-2..i0+33 or i0-3..i0+32, j0-2..j0+TY), each filled by ONE
//     cp.async.bulk.tensor.4d (TMA, tensor map over A[v][k][j][i]) that completes on an
//     mbarrier; plane k+3 is requested while plane k is being computed, so no thread ever
//     waits on a global load for the stencil (ncu of the LDG version: long-scoreboard was the
//     second stall reason) and every stencil operand is an LDS with an IMMEDIATE offset
//     (the LDG version spent ~4 integer instructions of 64-bit address arithmetic per load);
//   * the otherwise idle extra row (warp TY-1, which only produces y fluxes) is the TMA
//     producer: it waits on the `empty` mbarrier (every consumer warp arrives after its z
//     flux, the last reader of plane k-1) and re-fills that buffer with plane k+3;
//   * the step loop is unrolled with compile-time axes, so the solver-frame rotation is
//     register renaming and the variable permutation of eqns_base::SetDirection
//     (eqns_base.cpp:94-131) is folded into the LDS offsets.
//
// Out-of-range box coordinates (first-order grids have one ghost layer) are zero-filled by
// the TMA unit; those values only reach threads whose results are discarded.
#pragma once
#include <cuda.h>
#include <cstdlib>
#include "stage_sweep.cuh"

namespace pion {

#ifndef PION_TMA_UNROLL
#define PION_TMA_UNROLL 1
#endif
#ifndef PION_TMA_PRODUCER_SLEEP_NS
#define PION_TMA_PRODUCER_SLEEP_NS 4000
#endif
#ifndef PION_TMA_ONEFLUX
#define PION_TMA_ONEFLUX 1
#endif

constexpr int TMA_CW = 36;                                            // tile columns
__host__ __device__ constexpr int tma_rh(int ty) { return ty + 3; }  // tile rows
__host__ __device__ constexpr int tma_plane_bytes(int nb, int ty) { return nb * tma_rh(ty) * TMA_CW * 8; }
__host__ __device__ constexpr int tma_plane_stride(int nb, int ty) { return (tma_plane_bytes(nb, ty) + 127) / 128 * 128; }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_spin(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// same, with a suspend-time hint (ns): the waiting thread sleeps in hardware until the phase completes
// instead of re-issuing try_wait (the producer lane waits most of a plane time for the consumers)
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long* bar, unsigned parity, unsigned hint_ns) {
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_plane(void* dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(0)
      : "memory");
}

// primitive state of the tile cell `p` points at (p = address of its density), solver frame of axis ax
template <int EQ, int VS>
__device__ __forceinline__ Prim lds_prim(const double* p, int ax, int a1, int a2) {
  Prim q;
  q.ro = p[0];
  q.pg = p[VS];
  q.vn = p[(2 + ax) * VS];
  q.vt1 = p[(2 + a1) * VS];
  q.vt2 = p[(2 + a2) * VS];
  if (EQ != EQ_EULER) {
    q.bn = p[(5 + ax) * VS];
    q.bt1 = p[(5 + a1) * VS];
    q.bt2 = p[(5 + a2) * VS];
  } else {
    q.bn = q.bt1 = q.bt2 = 0.0;
  }
  q.psi = (EQ == EQ_GLM) ? p[8 * VS] : 0.0;
  return q;
}

// edge states at the face between tile cells pL | pR (pQ0, pQ3: the next cells outwards), solver frame of
// axis ax: SetEdgeState / SetSlope (VectorOps.cpp:535-617)
template <int EQ, int VS>
__device__ __forceinline__ void edge_states_tile(const StageArgs& a, const double* pQ0, const double* pL, const double* pR,
                                                 const double* pQ3, int ax, int a1, int a2, Prim& eL, Prim& eR) {
  eL = lds_prim<EQ, VS>(pL, ax, a1, a2);
  eR = lds_prim<EQ, VS>(pR, ax, a1, a2);
  if (a.order == 2) {
    const Prim Q0 = lds_prim<EQ, VS>(pQ0, ax, a1, a2);
    const Prim Q3 = lds_prim<EQ, VS>(pQ3, ax, a1, a2);
#define PION_EDGE2(f)                                                     \
  {                                                                       \
    const double d0 = eL.f - Q0.f, d1 = eR.f - eL.f, d2 = Q3.f - eR.f;    \
    eL.f += minmod(d0, d1, a.tiny2) * 0.5;                                \
    eR.f -= minmod(d1, d2, a.tiny2) * 0.5;                                \
  }
    PION_EDGE2(ro) PION_EDGE2(pg) PION_EDGE2(vn) PION_EDGE2(vt1) PION_EDGE2(vt2)
    if (EQ != EQ_EULER) { PION_EDGE2(bn) PION_EDGE2(bt1) PION_EDGE2(bt2) }
    if (EQ == EQ_GLM) { PION_EDGE2(psi) }
#undef PION_EDGE2
  }
}

// flux through that face, as low_face_flux
template <int EQ, int SOLVER, bool FKJ, int VS>
__device__ __forceinline__ void face_flux_tile(const StageArgs& a, const double* pQ0, const double* pL, const double* pR,
                                               const double* pQ3, bool use_hll, int ax, int a1, int a2, Cons& F) {
  Prim eL, eR;
  edge_states_tile<EQ, VS>(a, pQ0, pL, pR, pQ3, ax, a1, a2, eL, eR);
  intercell_flux<EQ, SOLVER, FKJ ? AV_FKJ98 : AV_NONE>(eL, eR, a.pp, use_hll, 0.0, F);
}

// dU accumulators in the GRID frame (x, y, z): with compile-time axes nothing has to be rotated
struct NatAcc {
  double rho, erg, m0, m1, m2, b0, b1, b2, psi;
};
// component q (0,1,2) of a triple, q known at compile time
template <int Q>
__device__ __forceinline__ double& pick3(double& c0, double& c1, double& c2) { return Q == 0 ? c0 : Q == 1 ? c1 : c2; }
template <int Q>
__device__ __forceinline__ double pick3c(double c0, double c1, double c2) { return Q == 0 ? c0 : Q == 1 ? c1 : c2; }

// acc += dt (F_low - F_high) / dx for the axis AX, D in that axis' solver frame (dU_Cell + DivStateVectorComponent)
template <int EQ, int AX>
__device__ __forceinline__ void acc_flux_diff(NatAcc& A, const Cons& D, double dt, double idx, double dtdx) {
  constexpr int A1 = (AX + 1) % 3, A2 = (AX + 2) % 3;
#ifdef PION_STRICT
#define PION_ACC(dst, src) dst += dt * (src * idx);
#else
#define PION_ACC(dst, src) dst = fma(dtdx, src, dst);
#endif
  PION_ACC(A.rho, D.rho) PION_ACC(A.erg, D.erg)
  PION_ACC(pick3<AX>(A.m0, A.m1, A.m2), D.mn) PION_ACC(pick3<A1>(A.m0, A.m1, A.m2), D.mt1) PION_ACC(pick3<A2>(A.m0, A.m1, A.m2), D.mt2)
  if (EQ != EQ_EULER) {
    PION_ACC(pick3<AX>(A.b0, A.b1, A.b2), D.bbn) PION_ACC(pick3<A1>(A.b0, A.b1, A.b2), D.bbt1) PION_ACC(pick3<A2>(A.b0, A.b1, A.b2), D.bbt2)
  }
  if (EQ == EQ_GLM) { PION_ACC(A.psi, D.psi) }
#undef PION_ACC
}

// Powell + GLM sources of the two interfaces of the cell along AX, from cell-centre states
// (solver_eqn_mhd_adi.cpp:396-443,782-813): R part of interface (i-1,i), then L part of interface (i,i+1).
// C is the centre state in the grid frame (vn,vt1,vt2 = vx,vy,vz); qm / qp the tile cells at -1 / +1 along AX.
template <int EQ, int VS, int AX>
__device__ __forceinline__ void acc_sources(NatAcc& A, const Prim& C, double uB, const double* qm, const double* qp, double dt,
                                            double idx, double hdtdx) {
  if (EQ == EQ_EULER) return;
  const double bm = qm[(5 + AX) * VS], bp = qp[(5 + AX) * VS];
  const double Bn = pick3c<AX>(C.bn, C.bt1, C.bt2), Vn = pick3c<AX>(C.vn, C.vt1, C.vt2);
#ifdef PION_STRICT
  double f = dt * (0.5 * (bm + Bn));
  A.m0 += f * C.bn * idx; A.m1 += f * C.bt1 * idx; A.m2 += f * C.bt2 * idx; A.erg += f * uB * idx;
  A.b0 += f * C.vn * idx; A.b1 += f * C.vt1 * idx; A.b2 += f * C.vt2 * idx;
  double psm = 0.0, psp = 0.0;
  if (EQ == EQ_GLM) {
    psm = qm[8 * VS];
    psp = qp[8 * VS];
    double fs = dt * (0.5 * (psm + C.psi));
    A.erg += fs * (Vn * C.psi) * idx;
    A.psi += fs * Vn * idx;
  }
  f = dt * (0.5 * (Bn + bp));
  A.m0 -= f * C.bn * idx; A.m1 -= f * C.bt1 * idx; A.m2 -= f * C.bt2 * idx; A.erg -= f * uB * idx;
  A.b0 -= f * C.vn * idx; A.b1 -= f * C.vt1 * idx; A.b2 -= f * C.vt2 * idx;
  if (EQ == EQ_GLM) {
    double fs = dt * (0.5 * (C.psi + psp));
    A.erg -= fs * (Vn * C.psi) * idx;
    A.psi -= fs * Vn * idx;
  }
#else
  // the two halves regrouped: (dt/2dx)(bm + Bn) X - (dt/2dx)(Bn + bp) X = (dt/2dx)(bm - bp) X
  const double gB = hdtdx * (bm - bp);
  A.m0 = fma(gB, C.bn, A.m0); A.m1 = fma(gB, C.bt1, A.m1); A.m2 = fma(gB, C.bt2, A.m2);
  A.erg = fma(gB, uB, A.erg);
  A.b0 = fma(gB, C.vn, A.b0); A.b1 = fma(gB, C.vt1, A.b1); A.b2 = fma(gB, C.vt2, A.b2);
  if (EQ == EQ_GLM) {
    const double gS = hdtdx * (qm[8 * VS] - qp[8 * VS]);
    A.erg = fma(gS, Vn * C.psi, A.erg);
    A.psi = fma(gS, Vn, A.psi);
  }
#endif
}

__device__ __forceinline__ void cons_diff(Cons& D, const Cons& lo, const Cons& hi) {
  D.rho = lo.rho - hi.rho; D.erg = lo.erg - hi.erg; D.mn = lo.mn - hi.mn; D.mt1 = lo.mt1 - hi.mt1; D.mt2 = lo.mt2 - hi.mt2;
  D.bbn = lo.bbn - hi.bbn; D.bbt1 = lo.bbt1 - hi.bbt1; D.bbt2 = lo.bbt2 - hi.bbt2; D.psi = lo.psi - hi.psi;
}

template <int EQ, int SOLVER, bool FKJ, int TY, int MINB>
__global__ void __launch_bounds__(32 * TY, MINB)
    k_stage_sweep_tma(const __grid_constant__ StageArgs a, const __grid_constant__ CUtensorMap tmap, const int kchunk) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  constexpr int NB = nbase(EQ);
  constexpr int RH = tma_rh(TY), CW = TMA_CW;
  constexpr int VS = RH * CW;                            // doubles between variables of one tile
  constexpr int PS = tma_plane_stride(NB, TY) / 8;       // doubles between plane buffers
  constexpr unsigned PLANE_BYTES = tma_plane_bytes(NB, TY);
  constexpr int SLAB = NB * TY * 32;
  double* const s_tile = reinterpret_cast<double*>(s_raw);              // [4][NB][RH][CW]
  double* const s_flux = s_tile + 4 * PS;                                // [2][NB][TY][32]
  __shared__ unsigned long long s_bar;       // y-flux slab published (all threads arrive)
  __shared__ unsigned long long s_full[4];   // plane buffer filled (TMA transaction bytes)
  __shared__ unsigned long long s_empty;     // plane k-1 no longer read (one arrive per consumer warp)

  const GridD& g = a.g;
  const int NX = g.NG[0], NY = g.NG[1];
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int i0 = (blockIdx.x + a.tx0) * 31, j0 = (blockIdx.y + a.ty0) * (TY - 1);
  int i = i0 + lane, j = j0 + row;
  const bool row_active = (row < TY - 1) && (j < NY);               // warp-uniform
  const bool upd_xy = row_active && (lane < 31) && (i < NX);
  i = min(i, NX);  // global index only feeds the flag loads / Pb load of discarded threads
  j = min(j, NY);
  const int k0 = a.k_lo + blockIdx.z * kchunk, k1 = min(k0 + kchunk, a.k_hi);
  const int nk = k1 - k0;
  const long vs = g.vs;
  const double dt = a.dt;
  const double idx = 1.0 / g.dx;
#ifndef PION_STRICT
  const double dtdx = dt * idx, hdtdx = 0.5 * dt * idx;
#endif
  double my_dt = 1.0e100;
  int status = 0;
  const bool producer = (row == TY - 1) && (lane == 0);
  const bool pb_is_s = (a.Pb == a.S);  // predictor: the base state is the stencil centre already in registers

  // tile coordinates of the box (element units of the tensor map: x, y, z, v)
  // (the box must start on a 16-byte boundary in x -- an odd element offset faults on B200,
  // tools/micro/tma_probe.cu -- so odd starts load one column earlier and the threads shift by one)
  const int bx_cell = g.xoff + g.nb[0] - 2 + i0;
  const int xshift = bx_cell & 1;
  const int bx = bx_cell - xshift, by = g.nb[1] - 2 + j0, bz = g.nb[2] + k0 - 2;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 32 * TY);
    mbar_init(&s_empty, TY - 1);
#pragma unroll
    for (int q = 0; q < 4; q++) mbar_init(&s_full[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (producer) {
#pragma unroll
    for (int q = 0; q < 4; q++) {  // planes k0-2 .. k0+1
      mbar_expect_tx(&s_full[q], PLANE_BYTES);
      tma_load_plane(s_tile + q * PS, &tmap, &s_full[q], bx, by, bz + q);
    }
  }
  unsigned phase = 0;  // parity of the y-flux barrier

  Cons Fz;  // flux through the low z face of the current cell
  cons_zero<EQ>(Fz);
  const int coff = (row + 2) * CW + lane + 2 + xshift;  // this thread's cell inside a plane buffer

  // HLLD->HLL switch flags (solver_eqn_mhd_adi.cpp:167-177) of the cells (i,j,k), (i-1,j,k), (i,j,k+1),
  // (i,j-1,k+1): loaded ONE PLANE AHEAD so that the flux never waits on them
  unsigned f_c = 0, f_xm = 0, f_zp = 0, f_ym = 0;
  const unsigned char* hp = nullptr;  // flag of cell (i,j,k)
  if (SOLVER == SOLVE_HLLD) {
    hp = a.hll + gidx(g, i + g.nb[0], j + g.nb[1], k0 - 1 + g.nb[2]);
    f_c = hp[0];
    f_zp = hp[g.sz];
    f_ym = hp[g.sz - g.sy];
  }

  for (int kk = -1; kk < nk; kk++) {
    const int k = k0 + kk;
    const bool warm = kk < 0;  // first iteration of a chunk: only the fluxes INTO plane k0 (z face, y faces)
    const long c = gidx(g, i + g.nb[0], j + g.nb[1], k + g.nb[2]);
    // plane p lives in buffer (p - k0 + 2) & 3
    const double* const pm1 = s_tile + ((kk + 1) & 3) * PS + coff;  // plane k-1
    const double* const p0 = s_tile + ((kk + 2) & 3) * PS + coff;   // plane k
    const double* const pp1 = s_tile + ((kk + 3) & 3) * PS + coff;  // plane k+1
    const double* const pp2 = s_tile + (kk & 3) * PS + coff;        // plane k+2
    double* const sbuf = s_flux + (size_t)(kk & 1) * SLAB;
    unsigned n_xm = 0, n_zp = 0, n_ym = 0;  // next plane's flags
    if (SOLVER == SOLVE_HLLD && kk + 1 < nk) {
      n_xm = hp[g.sz - 1];
      n_zp = hp[2 * g.sz];
      n_ym = hp[2 * g.sz - g.sy];
    }

    if (warm) {  // planes k0-2, k0-1, k0 (plane k0+1 = "k+2" is waited for below like every iteration)
      mbar_wait_spin(&s_full[0], 0);
      mbar_wait_spin(&s_full[1], 0);
      mbar_wait_spin(&s_full[2], 0);
    }

#if PION_TMA_ONEFLUX
    // ---- ONE copy of the Riemann solver: a real loop over the three faces (x of plane k, z high face,
    // y of plane k+1) whose per-face parts -- stencil loads + reconstruction before the solver, flux
    // exchange + accumulation after it -- are compile-time specialised per axis, so the hot loop stays
    // inside the instruction cache and still has immediate LDS offsets and no frame rotation.
    NatAcc acc;
    acc.rho = acc.erg = acc.m0 = acc.m1 = acc.m2 = acc.b0 = acc.b1 = acc.b2 = acc.psi = 0.0;
    Prim C;
    C.ro = C.pg = C.vn = C.vt1 = C.vt2 = C.bn = C.bt1 = C.bt2 = C.psi = 0.0;
    const bool domain = upd_xy && !warm && (a.mask ? (a.mask[c] != 0) : true);
    double uB = 0.0;
    if (!warm) {
      C = lds_prim<EQ, VS>(p0, 0, 1, 2);
      if (a.mp_dE && upd_xy) acc.erg = a.mp_dE[c];  // cooling source term (energy only)
      if (EQ != EQ_EULER) uB = C.bn * C.vn + C.bt1 * C.vt1 + C.bt2 * C.vt2;
    }
#ifdef PION_STRICT
    const double dtdx = 0.0, hdtdx = 0.0;
#endif
#pragma unroll 1
    for (int f = warm ? 1 : 0; f < 3; f++) {
      if (f == 2 && kk + 1 >= nk) break;
      Cons Fnew;
      cons_zero<EQ>(Fnew);
      if (row_active || f == 2) {
        Prim eL, eR;
        bool use_hll = false;
        if (f == 0) {
          edge_states_tile<EQ, VS>(a, p0 - 2, p0 - 1, p0, p0 + 1, 0, 1, 2, eL, eR);
          if (SOLVER == SOLVE_HLLD) use_hll = (f_xm | f_c) != 0;
        } else if (f == 1) {
          // plane k+2 (first needed here): fill number (kk+4) >> 2 of buffer kk & 3
          mbar_wait_spin(&s_full[kk & 3], ((unsigned)(kk + 4) >> 2) & 1u);
          edge_states_tile<EQ, VS>(a, pm1, p0, pp1, pp2, 2, 0, 1, eL, eR);
          if (SOLVER == SOLVE_HLLD) use_hll = (f_c | f_zp) != 0;
        } else {
          edge_states_tile<EQ, VS>(a, pp1 - 2 * CW, pp1 - CW, pp1, pp1 + CW, 1, 2, 0, eL, eR);
          if (SOLVER == SOLVE_HLLD) use_hll = (f_ym | f_zp) != 0;
        }
        intercell_flux<EQ, SOLVER, FKJ ? AV_FKJ98 : AV_NONE>(eL, eR, a.pp, use_hll, 0.0, Fnew);
      } else if (f == 1) {
        mbar_wait_spin(&s_full[kk & 3], ((unsigned)(kk + 4) >> 2) & 1u);
      }
      Cons D;
      if (f == 0) {
        const Cons Fh = cons_shfl_down<EQ>(Fnew);
        cons_diff(D, Fnew, Fh);
        acc_sources<EQ, VS, 0>(acc, C, uB, p0 - 1, p0 + 1, dt, idx, hdtdx);
        acc_flux_diff<EQ, 0>(acc, D, dt, idx, dtdx);
        // y: both faces come from the slab the previous iteration published
        mbar_wait_spin(&s_bar, phase);
        phase ^= 1u;
        const int rn = min(row + 1, TY - 1);
        const Cons Fl = cons_from_smem<EQ, TY>(sbuf, row, lane);
        const Cons Fhy = cons_from_smem<EQ, TY>(sbuf, rn, lane);
        cons_diff(D, Fl, Fhy);
        acc_sources<EQ, VS, 1>(acc, C, uB, p0 - CW, p0 + CW, dt, idx, hdtdx);
        acc_flux_diff<EQ, 1>(acc, D, dt, idx, dtdx);
      } else if (f == 1) {
        cons_diff(D, Fz, Fnew);
        Fz = Fnew;
        if (!warm) {
          acc_sources<EQ, VS, 2>(acc, C, uB, pm1, pp1, dt, idx, hdtdx);
          acc_flux_diff<EQ, 2>(acc, D, dt, idx, dtdx);
        }
        // plane k-1 (z flux Q0, z sources) has been read for the last time by this warp
        __syncwarp();
        if (lane == 0 && row < TY - 1) mbar_arrive(&s_empty);
      } else {
        double* nbuf = s_flux + (size_t)((kk + 1) & 1) * SLAB;  // slab of plane k+1
        cons_to_smem<EQ, TY>(nbuf, row, lane, Fnew);
        mbar_arrive(&s_bar);
      }
    }
    Cons accx;  // grid frame == solver frame of x
    accx.rho = acc.rho; accx.erg = acc.erg; accx.mn = acc.m0; accx.mt1 = acc.m1; accx.mt2 = acc.m2;
    accx.bbn = acc.b0; accx.bbt1 = acc.b1; accx.bbt2 = acc.b2; accx.psi = acc.psi;
#else
    Cons acc;
    cons_zero<EQ>(acc);
    Prim C;
    C.ro = C.pg = C.vn = C.vt1 = C.vt2 = C.bn = C.bt1 = C.bt2 = C.psi = 0.0;
    const bool domain = upd_xy && !warm && (a.mask ? (a.mask[c] != 0) : true);
    double uB = 0.0;
    if (!warm) {
      C = lds_prim<EQ, VS>(p0, 0, 1, 2);
      if (a.mp_dE && upd_xy) acc.erg = a.mp_dE[c];  // cooling source term (energy only)
      if (EQ != EQ_EULER) uB = C.bn * C.vn + C.bt1 * C.vt1 + C.bt2 * C.vt2;
    }

    // Schedule of one plane (as k_stage_sweep):
    //   step 1: x flux (shfl), accumulate x          step 2: wait, accumulate y from shared memory
    //   step 3: z flux, accumulate z                 step 4: y flux of the NEXT plane -> shared memory, arrive
#if PION_TMA_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int step = 1; step <= 4; step++) {
      if (warm && step < 3) continue;
      if (step == 4 && kk + 1 >= nk) continue;
      const int ax = (step == 1) ? 0 : (step == 3) ? 2 : 1;
      const int a1 = (ax == 2) ? 0 : ax + 1;
      const int a2 = (a1 == 2) ? 0 : a1 + 1;
      Cons Fnew;
      cons_zero<EQ>(Fnew);
      if (step == 3) {
        // plane k+2 (first needed by the z flux): fill number (kk+4) >> 2 of buffer kk & 3
        mbar_wait_spin(&s_full[kk & 3], ((unsigned)(kk + 4) >> 2) & 1u);
      }
      if (step != 2 && (row_active || ax == 1)) {
        // ONE flux call site: the four stencil cells of the face and its HLLD->HLL switch
        const double *pQ0, *pL, *pR, *pQ3;
        bool use_hll = false;
        if (step == 1) {
          pQ0 = p0 - 2; pL = p0 - 1; pR = p0; pQ3 = p0 + 1;
          if (SOLVER == SOLVE_HLLD) use_hll = (f_xm | f_c) != 0;
        } else if (step == 3) {
          pQ0 = pm1; pL = p0; pR = pp1; pQ3 = pp2;
          if (SOLVER == SOLVE_HLLD) use_hll = (f_c | f_zp) != 0;
        } else {
          pQ0 = pp1 - 2 * CW; pL = pp1 - CW; pR = pp1; pQ3 = pp1 + CW;
          if (SOLVER == SOLVE_HLLD) use_hll = (f_ym | f_zp) != 0;
        }
        face_flux_tile<EQ, SOLVER, FKJ, VS>(a, pQ0, pL, pR, pQ3, use_hll, ax, a1, a2, Fnew);
      }
      if (step == 4) {
        double* nbuf = s_flux + (size_t)((kk + 1) & 1) * SLAB;  // slab of plane k+1
        cons_to_smem<EQ, TY>(nbuf, row, lane, Fnew);
        mbar_arrive(&s_bar);
        continue;
      }

      Cons D;
#define PION_DIFF(LOW, HIGH)                                                                     \
  D.rho = LOW.rho - HIGH.rho; D.erg = LOW.erg - HIGH.erg; D.mn = LOW.mn - HIGH.mn;               \
  D.mt1 = LOW.mt1 - HIGH.mt1; D.mt2 = LOW.mt2 - HIGH.mt2;                                        \
  if (EQ != EQ_EULER) { D.bbn = LOW.bbn - HIGH.bbn; D.bbt1 = LOW.bbt1 - HIGH.bbt1; D.bbt2 = LOW.bbt2 - HIGH.bbt2; } \
  else { D.bbn = D.bbt1 = D.bbt2 = 0.0; }                                                        \
  D.psi = (EQ == EQ_GLM) ? LOW.psi - HIGH.psi : 0.0;
      if (step == 1) {
        const Cons Fh = cons_shfl_down<EQ>(Fnew);
        PION_DIFF(Fnew, Fh)
      } else if (step == 2) {
        mbar_wait_spin(&s_bar, phase);
        phase ^= 1u;
        const int rn = min(row + 1, TY - 1);
        const Cons Fl = cons_from_smem<EQ, TY>(sbuf, row, lane);
        const Cons Fh = cons_from_smem<EQ, TY>(sbuf, rn, lane);
        PION_DIFF(Fl, Fh)
      } else {
        PION_DIFF(Fz, Fnew)
        Fz = Fnew;
      }
#undef PION_DIFF

      if (!warm) {
        // Powell + GLM sources from cell-centre states (solver_eqn_mhd_adi.cpp:396-443,782-813):
        // R part of interface (i-1,i), then L part of interface (i,i+1)
        if (EQ != EQ_EULER) {
          const double* qm = (step == 1) ? p0 - 1 : (step == 2) ? p0 - CW : pm1;
          const double* qp = (step == 1) ? p0 + 1 : (step == 2) ? p0 + CW : pp1;
          const double bm = qm[(5 + ax) * VS], bp = qp[(5 + ax) * VS];
#ifdef PION_STRICT
          double f = dt * (0.5 * (bm + C.bn));
          acc.mn += f * C.bn * idx; acc.mt1 += f * C.bt1 * idx; acc.mt2 += f * C.bt2 * idx; acc.erg += f * uB * idx;
          acc.bbn += f * C.vn * idx; acc.bbt1 += f * C.vt1 * idx; acc.bbt2 += f * C.vt2 * idx;
          double psm = 0.0, psp = 0.0;
          if (EQ == EQ_GLM) {
            psm = qm[8 * VS];
            psp = qp[8 * VS];
            double fs = dt * (0.5 * (psm + C.psi));
            acc.erg += fs * (C.vn * C.psi) * idx;
            acc.psi += fs * C.vn * idx;
          }
          f = dt * (0.5 * (C.bn + bp));
          acc.mn -= f * C.bn * idx; acc.mt1 -= f * C.bt1 * idx; acc.mt2 -= f * C.bt2 * idx; acc.erg -= f * uB * idx;
          acc.bbn -= f * C.vn * idx; acc.bbt1 -= f * C.vt1 * idx; acc.bbt2 -= f * C.vt2 * idx;
          if (EQ == EQ_GLM) {
            double fs = dt * (0.5 * (C.psi + psp));
            acc.erg -= fs * (C.vn * C.psi) * idx;
            acc.psi -= fs * C.vn * idx;
          }
#else
          // the two halves regrouped: (dt/2dx)(bm + Bn) X - (dt/2dx)(Bn + bp) X = (dt/2dx)(bm - bp) X
          const double gB = hdtdx * (bm - bp);
          acc.mn = fma(gB, C.bn, acc.mn); acc.mt1 = fma(gB, C.bt1, acc.mt1); acc.mt2 = fma(gB, C.bt2, acc.mt2);
          acc.erg = fma(gB, uB, acc.erg);
          acc.bbn = fma(gB, C.vn, acc.bbn); acc.bbt1 = fma(gB, C.vt1, acc.bbt1); acc.bbt2 = fma(gB, C.vt2, acc.bbt2);
          if (EQ == EQ_GLM) {
            const double gS = hdtdx * (qm[8 * VS] - qp[8 * VS]);
            acc.erg = fma(gS, C.vn * C.psi, acc.erg);
            acc.psi = fma(gS, C.vn, acc.psi);
          }
#endif
        }
        // flux difference (dU_Cell + DivStateVectorComponent)
#ifdef PION_STRICT
#define PION_ACC(f) acc.f += dt * (D.f * idx);
#else
#define PION_ACC(f) acc.f = fma(dtdx, D.f, acc.f);
#endif
        PION_ACC(rho) PION_ACC(erg) PION_ACC(mn) PION_ACC(mt1) PION_ACC(mt2)
        if (EQ != EQ_EULER) { PION_ACC(bbn) PION_ACC(bbt1) PION_ACC(bbt2) }
        if (EQ == EQ_GLM) { PION_ACC(psi) }
#undef PION_ACC
        // rotate the centre state and the accumulators into the next axis' frame
        rot3(C.vn, C.vt1, C.vt2);
        rot3(acc.mn, acc.mt1, acc.mt2);
        if (EQ != EQ_EULER) {
          rot3(C.bn, C.bt1, C.bt2);
          rot3(acc.bbn, acc.bbt1, acc.bbt2);
        }
      }
      if (step == 3) {
        // plane k-1 (z flux Q0, z sources) has been read for the last time by this warp
        __syncwarp();
        if (lane == 0 && row < TY - 1) mbar_arrive(&s_empty);
      }
    }
#endif
    // the producer refills the buffer of plane k-1 with plane k+3 (needed by the next iteration's z flux)
    if (producer && kk + 1 < nk) {
      mbar_wait_sleep(&s_empty, (unsigned)(kk + 1) & 1u, PION_TMA_PRODUCER_SLEEP_NS);
      unsigned long long* fb = &s_full[(kk + 1) & 3];
      mbar_expect_tx(fb, PLANE_BYTES);
      tma_load_plane(s_tile + ((kk + 1) & 3) * PS, &tmap, fb, bx, by, bz + kk + 5);
    }
    if (SOLVER == SOLVE_HLLD) {
      f_c = f_zp; f_xm = n_xm; f_zp = n_zp; f_ym = n_ym;
      hp += g.sz;
    }
    if (warm) continue;

    if (domain) {
#if PION_TMA_ONEFLUX
      if (pb_is_s) status |= cell_advance_time_pb<EQ>(a, c, C, accx, nullptr, 0, my_dt);
      else status |= cell_advance_time<EQ>(a, c, accx, nullptr, 0, my_dt);
#else
      if (pb_is_s) status |= cell_advance_time_pb<EQ>(a, c, C, acc, nullptr, 0, my_dt);
      else status |= cell_advance_time<EQ>(a, c, acc, nullptr, 0, my_dt);
#endif
    } else if (upd_xy && a.out != a.S) {
      // cell cut out of the domain (time_integrator.cpp:905-908): state untouched
      for (int v = 0; v < NB; v++) a.out[(long)v * vs + c] = a.S[(long)v * vs + c];
    }
  }

  stage_block_epilogue(a, my_dt, status);
}

template <int EQ, int SOLVER, bool FKJ>
inline void launch_sweep_tma_t(const StageArgs& a, cudaStream_t s) {
  constexpr int TY = sweep_ty(EQ), MINB = sweep_minb(EQ);
  constexpr int NB = nbase(EQ);
  const int bx = a.tx1 - a.tx0, by = a.ty1 - a.ty0, NZ = a.k_hi - a.k_lo;
  if (bx <= 0 || by <= 0 || NZ <= 0) return;
  int kchunk = 64;
  while (kchunk > 8 && (long)bx * by * ((NZ + kchunk - 1) / kchunk) < 148L * 4) kchunk >>= 1;
  const int bz = (NZ + kchunk - 1) / kchunk;
  const size_t smem = (size_t)4 * tma_plane_stride(NB, TY) + (size_t)2 * NB * TY * 32 * sizeof(double);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_stage_sweep_tma<EQ, SOLVER, FKJ, TY, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  k_stage_sweep_tma<EQ, SOLVER, FKJ, TY, MINB><<<dim3(bx, by, bz), 32 * TY, smem, s>>>(a, *reinterpret_cast<const CUtensorMap*>(a.tmap), kchunk);
}

// 3-D grids without tracers / H-correction run the TMA kernel, everything else the LDG sweep kernel
template <int EQ, int SOLVER, bool FKJ>
inline void launch_sweep_any(const StageArgs& a, cudaStream_t s) {
  if (a.tmap && a.g.ndim == 3 && a.ntr == 0 && !a.eta) launch_sweep_tma_t<EQ, SOLVER, FKJ>(a, s);
  else launch_sweep_t<EQ, SOLVER, FKJ>(a, s);
}

// box of one TMA plane load for an equation set (host side: tensor-map creation)
inline void sweep_tma_box_impl(int eq, int* cw, int* rh, int* nb) {
  *cw = TMA_CW;
  *rh = tma_rh(sweep_ty(eq));
  *nb = nbase(eq);
}

}  // namespace pion
