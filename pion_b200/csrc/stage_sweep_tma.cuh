// pion_b200/csrc/stage_sweep_tma.cuh -- the 3-D production stage kernel: the flux-once
// sweep of stage_sweep.cuh with the stencil STAGED IN SHARED MEMORY BY TMA.
//
// Same reference path and the same arithmetic as k_stage_sweep (time_integrator.cpp:498-958,
// VectorOps.cpp:535-644, solver_eqn_base.cpp:152-342, solver_eqn_mhd_adi.cpp:368-443,782-844);
// what changes is where the operands come from:
//
//   * a block is a 32 x (TY-1) tile of cells marching in z (warp = row, lane = cell) plus one LIGHT
//     warp that updates no cells: it produces the y fluxes through the tile's top edge and the x
//     fluxes through the tile's high-x edge (so all 32 lanes of the other rows update a cell);
//   * the state planes a block needs live in a ring of FOUR shared-memory plane buffers
//     [var][TY+3 rows][36 columns] (cells i0-2..i0+33, j0-2..j0+TY), each filled by ONE
//     cp.async.bulk.tensor.4d (TMA, tensor map over A[v][k][j][i]) that completes on an
//     mbarrier; plane k+3 is requested while plane k is being computed, so no thread ever
//     waits on a global load for the stencil (ncu of the LDG version: long-scoreboard was the
//     second stall reason) and every stencil operand is an LDS with an IMMEDIATE offset
//     (the LDG version spent ~4 integer instructions of 64-bit address arithmetic per load);
//   * the LAST warp to finish with plane k-1 (a running count in shared memory) issues the TMA
//     that refills its buffer with plane k+3: no producer warp, nobody spins on an "empty" barrier;
//   * the y-flux slab is single-buffered (a second split-phase mbarrier says when everybody has read it) and
//     the shared memory that frees holds the thread-private z flux carried to the next plane: 18 registers
//     fewer live across the solver (stage 19.4 -> 18.9 ms);
//   * one tracer can ride along as an extra tile variable (edge values, upwinded flux, the same exchanges);
//   * ONE copy of the Riemann solver inside a real loop over the three faces, with the per-face
//     parts (stencil loads + reconstruction before it, flux exchange + accumulation after it)
//     specialised at compile time: the hot loop fits the instruction cache (the fully unrolled
//     form stalled on instruction fetch, ncu no_instruction 0.86 warps/issue), the variable
//     permutation of eqns_base::SetDirection (eqns_base.cpp:94-131) is folded into the LDS
//     offsets, and dU is accumulated in the grid frame, so nothing is ever rotated.
//
// Out-of-range box coordinates (first-order grids have one ghost layer) are zero-filled by
// the TMA unit; those values only reach threads whose results are discarded.
#pragma once
#include <cuda.h>
#include <cstdint>
#include <cstdlib>
#include "stage_sweep.cuh"

namespace pion {


#ifndef PION_TMA_KCHUNK
#define PION_TMA_KCHUNK 64
#endif
// Tile rows and plane-ring depth per stage ORDER (MHD / GLM).  The predictor (ORDER 1, no reconstruction) only reads
// planes k-1 (Powell / GLM sources), k and k+1, so a ring of THREE plane buffers is enough for it, which leaves room
// for 16-row tiles (16 warps per SM at 128 registers instead of 12 at 168).  MEASURED AND NOT THE DEFAULT (r02o, 512^3
// GLM-HLLD predictor): 16 rows / ring 3 = 15.96 ms against 15.3 ms for 12 rows / ring 4 -- at 128 registers ptxas spills
// 16 words of loop state, the 225 KB of shared memory leave ~30 KB of L1, the spill reloads miss it (long-scoreboard
// 0.28 -> 1.63 warps per issue) and the FP64 pipe stays at 55 %; 15 rows: the same.  Parity is green for both
// (-DPION_TMA_TY1=16 -DPION_TMA_RING1=3), the knobs stay for that experiment.
#ifndef PION_TMA_TY1
#define PION_TMA_TY1 PION_SWEEP_TY
#endif
#ifndef PION_TMA_RING1
#define PION_TMA_RING1 4
#endif
__host__ __device__ constexpr int tma_ring(int eq, int order) { return (eq != EQ_EULER && order == 1) ? PION_TMA_RING1 : 4; }
constexpr int TMA_TX = 32;                                            // cells a tile updates along x
constexpr int TMA_CW = 36;                                            // tile columns: cells i0-2 .. i0+33
__host__ __device__ constexpr int tma_rh(int ty) { return ty + 3; }  // tile rows
__host__ __device__ constexpr int tma_plane_bytes(int nb, int ty) { return nb * tma_rh(ty) * TMA_CW * 8; }
__host__ __device__ constexpr int tma_plane_stride(int nb, int ty) { return (tma_plane_bytes(nb, ty) + 127) / 128 * 128; }
// dynamic shared memory of the TMA kernel: plane ring + y-flux slab + z-flux slots + x-edge slabs
__host__ __device__ constexpr size_t tma_smem_bytes(int nv, int ty, int ring) {
  return (size_t)ring * tma_plane_stride(nv, ty) + (size_t)2 * nv * ty * 32 * sizeof(double) + (size_t)3 * nv * ty * sizeof(double);
}
// shared memory one block may use so that `minb` blocks fit an SM (228 KB per SM, 1 KB reserved per block, 384 B static)
__host__ __device__ constexpr size_t tma_smem_budget(int minb) { return (size_t)(228 * 1024) / minb - 1024 - 512; }
// The corrector reads the cell's BASE state P (not the stencil state Ph) for cell_advance_time: nine / six global loads whose
// latency shows as long-scoreboard stalls (Euler corrector: 1.8 warps per issue, `profiles/r02A_wind384_*`; next to >= 203 KB of
// shared memory only 28 KB of L1 remain, so prefetch.global.L1 did not help).  Where the shared memory has room (Euler with
// at most one tracer; MHD / GLM tiles fill it), every thread copies its own cell's base state into a THREAD-PRIVATE slot with
// cp.async at the top of the iteration and reads it back three solves later: no registers held, no barrier needed.
__host__ __device__ constexpr bool tma_pb_stage(int eq, int order, int nv, int ty) {
#ifdef PION_NO_PB_STAGE
  return false;
#else
  return eq == EQ_EULER && order == 2 &&
         tma_smem_bytes(nv, ty, 4) + 128 + (size_t)nv * (ty - 1) * 32 * sizeof(double) <= tma_smem_budget(sweep_minb(eq));
#endif
}
// tracer counts the TMA kernel is instantiated for (tracers ride along as extra tile variables)
constexpr int TMA_MAXTR = 2;
// Tile rows for an equation set, stage order and tracer count: the equation set's row count (sweep_ty, or PION_TMA_TY1 for
// first-order stages of MHD / GLM), reduced until the ring + slabs of nbase + ntr variables fit the shared memory:
// Euler 8, 8, 7 rows for 0, 1, 2 tracers; ideal MHD 12, 12, 11; GLM 12, 11, 10.
__host__ __device__ constexpr int tma_ty(int eq, int order, int ntr) {
  int ty = (eq != EQ_EULER && order == 1) ? PION_TMA_TY1 : sweep_ty(eq);
  while (ty > 4 && tma_smem_bytes(nbase(eq) + ntr, ty, tma_ring(eq, order)) > tma_smem_budget(sweep_minb(eq))) ty--;
  return ty;
}

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
#ifdef PION_SPIN_SLEEP
    // a failed try costs three issue slots (YIELD, SYNCS, BRA) the other warps of the sub-partition could use:
    // ncu counted ~70 tries per wait.  Sleep a little between tries instead.
    if (!ok) __nanosleep(PION_SPIN_SLEEP);
#endif
  } while (!ok);
}
// 8-byte asynchronous copy global -> shared (LDGSTS), completion by cp.async.wait_all of the issuing thread
__device__ __forceinline__ void cp_async_f64(double* dst_smem, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void tma_load_plane(void* dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(0)
      : "memory");
}

// -DPION_RACE_STRESS: pseudo-random per-warp delays at the points where warps hand data to each other (slab publish /
// release, x-edge publish, plane refill), so that the parity suite runs under timings the normal build never produces.
// An ordering bug then shows as a parity failure (the tracer-slab read after release of r02t was found that way, by accident,
// on the slower strict build).  Compiled out by default.
#ifdef PION_RACE_STRESS
__device__ __forceinline__ void stress_delay(unsigned tag) {
  unsigned h = ((threadIdx.x >> 5) + 1u) * 2654435761u ^ (blockIdx.x * 40503u) ^ (tag * 0x9E3779B9u) ^ (unsigned)clock();
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
  h = __shfl_sync(0xffffffffu, h, 0);  // one decision per warp
  if ((h & 3u) == 0) __nanosleep(100u + (h >> 8) % 4000u);
}
#else
__device__ __forceinline__ void stress_delay(unsigned) {}
#endif

// one flag byte, issued HERE (volatile: the compiler may not sink it to its first use, a whole plane later)
__device__ __forceinline__ unsigned ldg_u8_now(const unsigned char* p) {
  unsigned v;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
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
template <int EQ, int VS, int ORDER>
__device__ __forceinline__ void edge_states_tile(const StageArgs& a, const double* pQ0, const double* pL, const double* pR,
                                                 const double* pQ3, int ax, int a1, int a2, Prim& eL, Prim& eR) {
  eL = lds_prim<EQ, VS>(pL, ax, a1, a2);
  eR = lds_prim<EQ, VS>(pR, ax, a1, a2);
  if (ORDER == 2) {
    const Prim Q0 = lds_prim<EQ, VS>(pQ0, ax, a1, a2);
    const Prim Q3 = lds_prim<EQ, VS>(pQ3, ax, a1, a2);
#ifdef PION_STRICT
#define PION_EDGE2(f)                                                     \
  {                                                                       \
    const double d0 = eL.f - Q0.f, d1 = eR.f - eL.f, d2 = Q3.f - eR.f;    \
    eL.f += minmod(d0, d1, a.tiny2) * 0.5;                                \
    eR.f -= minmod(d1, d2, a.tiny2) * 0.5;                                \
  }
#else
#define PION_EDGE2(f)                                                     \
  {                                                                       \
    const double d0 = eL.f - Q0.f, d1 = eR.f - eL.f, d2 = Q3.f - eR.f;    \
    add_limited(eL.f, d0, d1, 0.5);                                       \
    add_limited(eR.f, d1, d2, -0.5);                                      \
  }
#endif
    PION_EDGE2(ro) PION_EDGE2(pg) PION_EDGE2(vn) PION_EDGE2(vt1) PION_EDGE2(vt2)
    if (EQ != EQ_EULER) { PION_EDGE2(bn) PION_EDGE2(bt1) PION_EDGE2(bt2) }
    if (EQ == EQ_GLM) { PION_EDGE2(psi) }
#undef PION_EDGE2
  }
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

// tracers ride along as extra tile variables NB .. NB+NTR-1: edge values (SetEdgeState with the same minmod slopes)
// and the upwinded flux F[tr] = tr_upwind * F[rho] * sCMA corrector (solver_eqn_base.cpp:281-342)
template <int VS, int NB, int NTR, int ORDER>
__device__ __forceinline__ void tracer_edges_tile(const StageArgs& a, const double* pQ0, const double* pL, const double* pR,
                                                  const double* pQ3, double* trL, double* trR) {
#pragma unroll
  for (int q = 0; q < NTR; q++) {
    double L = pL[(NB + q) * VS], R = pR[(NB + q) * VS];
    if (ORDER == 2) {
      const double q0 = pQ0[(NB + q) * VS], q3 = pQ3[(NB + q) * VS];
      const double d0 = L - q0, d1 = R - L, d2 = q3 - R;
#ifdef PION_STRICT
      L += minmod(d0, d1, a.tiny2) * 0.5;
      R -= minmod(d1, d2, a.tiny2) * 0.5;
#else
      add_limited(L, d0, d1, 0.5);
      add_limited(R, d1, d2, -0.5);
#endif
    }
    trL[q] = L;
    trR[q] = R;
  }
}
__device__ __forceinline__ double tracer_upwind_flux(const StageArgs& a, double L, double R, double Frho) {
  double f = 0.0;
  if (Frho > 0.0) f = L * Frho * (a.pp.have_mp ? scma_corr(L) : 1.0);
  else if (Frho < 0.0) f = R * Frho * (a.pp.have_mp ? scma_corr(R) : 1.0);
  return f;
}

// ORDER: spatial order of the stage (1 = predictor of the second-order scheme / first-order runs, 2 = corrector);
// a template parameter so that the predictor carries no reconstruction code at all
template <int EQ, int SOLVER, bool FKJ, int TY, int MINB, int NTR, int ORDER, int RING>
__global__ void __launch_bounds__(32 * TY, MINB)
    k_stage_sweep_tma(const __grid_constant__ StageArgs a, const __grid_constant__ CUtensorMap tmap, const int kchunk) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  constexpr int NB = nbase(EQ);
  constexpr int NV = NB + NTR;                           // tile variables: the equations' + the tracers
  constexpr int NTRA = NTR > 0 ? NTR : 1;                // array extent (no zero-length arrays)
  constexpr int RH = tma_rh(TY), CW = TMA_CW;
  constexpr int VS = RH * CW;                            // doubles between variables of one tile
  constexpr int PS = tma_plane_stride(NV, TY) / 8;       // doubles between plane buffers
  constexpr unsigned PLANE_BYTES = tma_plane_bytes(NV, TY);
  constexpr int CS = TY * 32;                            // doubles between components of a flux slab
  constexpr int SLAB = NV * TY * 32;
  constexpr int XSLAB = NV * TY;
  static_assert(RING == 4 || (RING == 3 && ORDER == 1), "a three-plane ring holds planes k-1 .. k+1: first-order stages only");
  // plane k+d lives in buffer (kk + d + BASE) mod RING; the ring starts at plane k0-BASE-... (k0-2 | k0-1)
  constexpr int BASE = RING - 2;
  double* const s_tile = reinterpret_cast<double*>(s_raw);              // [RING][NV][RH][CW]
  double* const s_flux = s_tile + RING * PS;                                // [NV][TY][32] y fluxes of a plane, then [NV][TY][32] z fluxes
  double* const s_xedge = s_flux + 2 * SLAB;                             // [3][NV][TY] x flux through the tile's high x edge
  __shared__ unsigned long long s_bar;       // y-flux slab + x-edge fluxes published (all threads arrive)
  __shared__ unsigned long long s_free;      // y-flux slab read by everybody (the slab is single-buffered)
  double* const s_fz = s_flux + SLAB;        // [NV][TY][32] thread-private: the z flux carried to the next plane
  constexpr bool PBS = tma_pb_stage(EQ, ORDER, NV, TY);
  // [NV][TY-1][32] thread-private: the cell's base state, staged by cp.async (128-byte aligned behind the x-edge slabs)
  double* const s_pb = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(s_xedge + 3 * XSLAB) + 127) & ~(uintptr_t)127);
  __shared__ unsigned long long s_full[RING];   // plane buffer filled (TMA transaction bytes)
  __shared__ unsigned s_done;                // consumer warps that have finished reading plane k-1 (running count)

  const GridD& g = a.g;
  const int NX = g.NG[0], NY = g.NG[1];
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const BlockBox bb = sweep_block_box(a);
  const int i0 = bb.tx * TMA_TX, j0 = bb.ty * (TY - 1);
  // The LIGHT warp (row TY-1) updates no cells.  It produces the y fluxes through the tile's top edge (all
  // lanes), the x fluxes through the tile's high-x edge (lane r = row r, so that all 32 lanes of the
  // consumer rows update a cell), and its lane 0 issues the four TMA loads of the prologue (the refills are
  // issued by whichever consumer warp finishes with a plane last).
  const bool light = (row == TY - 1);
  int i = i0 + lane, j = j0 + row;
  const bool row_active = !light && (j < NY);               // warp-uniform
  const bool upd_xy = row_active && (i < NX);
  i = min(i, NX);  // global index only feeds the flag loads / Pb load of discarded threads
  j = min(j, NY);
  const int k0 = bb.k_lo + bb.tz * kchunk, k1 = min(k0 + kchunk, bb.k_hi);
  const int nk = k1 - k0;
  const long vs = g.vs;
  // (dt, 1/dx, dt/dx, dt/2dx are read from the kernel-parameter bank where they are used)
#define dt (a.dt)
#define idx (a.idx)
#define dtdx (a.dtdx)
#define hdtdx (a.hdtdx)
  double my_dt = 1.0e100;
  int status = 0;
  const bool producer = light && (lane == 0);
  const bool pb_is_s = (a.Pb == a.S);  // predictor: the base state is the stencil centre (re-read from the tile)

  // tile coordinates of the box (element units of the tensor map: x, y, z, v).  The box must start on a
  // 16-byte boundary in x -- an odd element offset faults on B200 (tools/micro/tma_probe.cu); with 32-cell
  // tiles and an even xoff + nb (checked by the host when it builds the tensor maps) it always does.
  const int bx = g.xoff + g.nb[0] - 2 + i0, by = g.nb[1] - 2 + j0, bz = g.nb[2] + k0 - 2;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 32 * TY);
    mbar_init(&s_free, 32 * TY);
    s_done = 0;
#pragma unroll
    for (int q = 0; q < RING; q++) mbar_init(&s_full[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (producer) {
#pragma unroll
    for (int q = 0; q < RING; q++) {  // planes k0-2 .. k0+1 (three-plane ring: k0-1 .. k0+1)
      mbar_expect_tx(&s_full[q], PLANE_BYTES);
      tma_load_plane(s_tile + q * PS, &tmap, &s_full[q], bx, by, bz + (2 - BASE) + q);
    }
  }
  unsigned phase = 0, fphase = 0;  // parities of the y-flux barriers (published / read by everybody)
  const int coff = (row + 2) * CW + lane + 2;  // this thread's cell inside a plane buffer
  // light warp, lane r: the cell just beyond the tile's high-x edge in row r (x-edge face = its low x face)
  const int erow = min(lane, TY - 2);
  const int eoff = (erow + 2) * CW + TMA_TX + 2;

  // HLLD->HLL switch (solver_eqn_mhd_adi.cpp:167-177) in face form (k_hll_face_flags: bit 0/1/2 = the low
  // x/y/z face of a cell runs HLL): this iteration needs the byte of the cells (i,j,k) and (i,j,k+1) -- the
  // light warp that of the cell beyond its x-edge face in plane k+1 -- and plane k+2's byte is requested ONE
  // PLANE AHEAD, so that the flux never waits on a global load
  unsigned w_k = 0, w_k1 = 0, we_k1 = 0;
  // isdomain byte of the cell (stellar-wind boundaries only): requested ONE PLANE AHEAD like the face bytes (the
  // synchronous load was 6 % of the Euler predictor's stall samples, profiles/r02p_*)
  const unsigned char* mp = nullptr;
  unsigned m_k = 1;
  if (a.mask) {
    mp = a.mask + gidx(g, i + g.nb[0], j + g.nb[1], k0 + g.nb[2]);
    m_k = mp[0];
  }
  const unsigned char* hp = nullptr;   // face byte of cell (i,j,k)
  const unsigned char* hpe = nullptr;  // light warp: face byte of cell (i0+32, j0+lane, k)
  if (SOLVER == SOLVE_HLLD) {
    hp = a.hllf + gidx(g, i + g.nb[0], j + g.nb[1], k0 - 1 + g.nb[2]);
    w_k = hp[0];
    w_k1 = hp[g.sz];
    if (light) {
      hpe = a.hllf + gidx(g, min(i0 + TMA_TX, NX + 1) + g.nb[0], min(j0 + erow, NY) + g.nb[1], k0 - 1 + g.nb[2]);
      we_k1 = hpe[g.sz];
    }
  }

  for (int kk = -1; kk < nk; kk++) {
    const int k = k0 + kk;
    const bool warm = kk < 0;  // first iteration of a chunk: only the fluxes INTO plane k0 (z face, y faces, x edge)
    const bool last = kk + 1 >= nk;
    const long c = gidx(g, i + g.nb[0], j + g.nb[1], k + g.nb[2]);
    // plane k+d lives in buffer (kk + d + BASE) mod RING (+RING: the warm iteration's plane k-1 of the three-plane
    // ring has no buffer; its pointer is formed but never dereferenced)
    const double* const pm1 = s_tile + ((unsigned)(kk - 1 + BASE + RING) % RING) * PS + coff;  // plane k-1
    const double* const p0 = s_tile + ((unsigned)(kk + BASE) % RING) * PS + coff;              // plane k
    const double* const pp1 = s_tile + ((unsigned)(kk + 1 + BASE) % RING) * PS + coff;         // plane k+1
    const double* const pp2 = s_tile + ((unsigned)(kk + 2 + BASE) % RING) * PS + coff;         // plane k+2 (four-plane ring only)
    // the plane this iteration reads for the first time: k+2 (ring of four: first needed by the z reconstruction),
    // k+1 (ring of three: by the light warp's x-edge flux, then by everybody's z flux)
    const unsigned m_new = (unsigned)(kk + 2 * BASE);
    unsigned long long* const bar_new = &s_full[m_new % RING];
    const unsigned par_new = (m_new / RING) & 1u;
    const double* const sbuf = s_flux;
    unsigned n_w = 0, n_we = 0;  // plane k+2's face bytes
    if (SOLVER == SOLVE_HLLD && !last) {
      n_w = ldg_u8_now(hp + 2 * g.sz);
      if (light) n_we = ldg_u8_now(hpe + 2 * g.sz);
    }

    if (warm) {  // planes k0-2, k0-1, k0 (plane k0+1 = "k+2" is waited for below like every iteration)
      mbar_wait_spin(&s_full[0], 0);
      if (RING == 4) {
        mbar_wait_spin(&s_full[1], 0);
        mbar_wait_spin(&s_full[2], 0);
      }
    }
    if (RING == 3 && light) mbar_wait_spin(bar_new, par_new);

    // ---- ONE copy of the Riemann solver: a real loop over the three faces (x of plane k, z high face,
    // y of plane k+1) whose per-face parts -- stencil loads + reconstruction before the solver, flux
    // exchange + accumulation after it -- are compile-time specialised per axis, so the hot loop stays
    // inside the instruction cache and still has immediate LDS offsets and no frame rotation.
    // Schedule of one plane: x flux, wait for the previous iteration's slab, accumulate x and y; z flux,
    // accumulate z; y flux (and, light warp, x-edge flux) of the NEXT plane -> shared memory, arrive.
    NatAcc acc;
    acc.rho = acc.erg = acc.m0 = acc.m1 = acc.m2 = acc.b0 = acc.b1 = acc.b2 = acc.psi = 0.0;
    double acctr[NTRA];
#pragma unroll
    for (int q = 0; q < NTRA; q++) acctr[q] = 0.0;
    // the centre state is RE-READ from the tile wherever it is needed (an LDS with an immediate offset is
    // cheaper than 18 registers held across the Riemann solver)
    // m_k: mask byte of plane k (the warm iteration, k = k0-1, already holds plane k0's and keeps it)
    unsigned n_m = 1;
    if (a.mask && !warm && !last) n_m = ldg_u8_now(mp + g.sz);
    const bool domain = upd_xy && !warm && (m_k != 0);
    if (!warm && a.mp_dE && upd_xy) acc.erg = a.mp_dE[c];  // cooling source term (energy only)
    const bool pb_staged = PBS && !pb_is_s;
    if (pb_staged && domain) {
#pragma unroll
      for (int v = 0; v < NV; v++) cp_async_f64(s_pb + (v * (TY - 1) + row) * 32 + lane, a.Pb + (long)v * vs + c);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
#ifdef PION_PB_PREFETCH
    // corrector: the base state P of this cell is read from HBM at the END of the iteration (cell_advance_time),
    // three Riemann solves from here; ask for its lines now so that those loads hit L1 (ncu: long-scoreboard
    // was 14 % of the corrector's stall samples, all on these loads)
    if (!pb_is_s && domain) {
#if PION_PB_PREFETCH == 2
      // into L2 only, by ONE lane per 128-byte line (lanes 0 and 16 of a row)
      if ((lane & 15) == 0) {
#pragma unroll
        for (int v = 0; v < NV; v++) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Pb + (long)v * vs + c));
      }
#else
#pragma unroll
      for (int v = 0; v < NV; v++) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.Pb + (long)v * vs + c));
#endif
    }
#endif
    // Order of the three solves of a plane.  Default x, z, y: the y fluxes of plane k+1 are published by the LAST solve of
    // an iteration and consumed by the FIRST of the next, so every warp waits there for the slowest one.  PION_TMA_YMID:
    // x, y, z -- the publish moves to the middle slot: one solve of slack between publish and consume, and one between
    // "slab read by everybody" and its overwrite (instead of none and two).  MEASURED (r02x, parity green): 18.00 vs 17.86 ms
    // per 512^3 GLM-HLLD stage, Wind3D 14.94 vs 14.74 ms per step -- the waits at the slab barriers are not what bounds the kernel.
    // Euler: three COPIES of the (small) solver with f a compile-time constant in each -- no loop-carried register moves, no
    // runtime dispatch on the face: +5 % on the 256^3 blast wave, +1.3 % on Wind3D (r02z).  For MHD / GLM the unrolled form
    // does not fit the instruction cache (52-61 KB of code with ONE copy of HLLD; the unrolled trial of round 1 stalled on
    // instruction fetch), so there it stays a real loop.
    constexpr bool UNROLL_FACES = (EQ == EQ_EULER);
#if defined(PION_TMA_YMID)
#pragma unroll 1
    for (int sl = (warm && !light) ? 1 : 0; sl < 3; sl++) {
      const int f = (sl == 0) ? 0 : (sl == 1) ? 2 : 1;
      if (f == 2 && last) continue;
#else
#pragma unroll(UNROLL_FACES ? 3 : 1)
    for (int f = UNROLL_FACES ? 0 : ((warm && !light) ? 1 : 0); f < 3; f++) {
      if (UNROLL_FACES && f == 0 && warm && !light) continue;
      if (f == 2 && last) {
        if (UNROLL_FACES) continue;
        else break;
      }
#endif
      Cons Fnew;
      cons_zero<EQ>(Fnew);
      double Ftr[NTRA];
#pragma unroll
      for (int q = 0; q < NTRA; q++) Ftr[q] = 0.0;
      // the light warp uses the x slot for the x-edge face of plane k+1
      const bool do_flux = (f == 0 && light) ? !last : (row_active || f == 2);
      if (do_flux) {
        Prim eL, eR;
        double trL[NTRA], trR[NTRA];
        bool use_hll = false;
        if (f == 0) {
          const double* const px = light ? (pp1 - coff + eoff) : p0;
          edge_states_tile<EQ, VS, ORDER>(a, px - 2, px - 1, px, px + 1, 0, 1, 2, eL, eR);
          tracer_edges_tile<VS, NB, NTR, ORDER>(a, px - 2, px - 1, px, px + 1, trL, trR);
          if (SOLVER == SOLVE_HLLD) use_hll = ((light ? we_k1 : w_k) & 1u) != 0;
        } else if (f == 1) {
          // the new plane (first needed here)
          mbar_wait_spin(bar_new, par_new);
          edge_states_tile<EQ, VS, ORDER>(a, pm1, p0, pp1, pp2, 2, 0, 1, eL, eR);
          tracer_edges_tile<VS, NB, NTR, ORDER>(a, pm1, p0, pp1, pp2, trL, trR);
          if (SOLVER == SOLVE_HLLD) use_hll = (w_k1 & 4u) != 0;
        } else {
          edge_states_tile<EQ, VS, ORDER>(a, pp1 - 2 * CW, pp1 - CW, pp1, pp1 + CW, 1, 2, 0, eL, eR);
          tracer_edges_tile<VS, NB, NTR, ORDER>(a, pp1 - 2 * CW, pp1 - CW, pp1, pp1 + CW, trL, trR);
          if (SOLVER == SOLVE_HLLD) use_hll = (w_k1 & 2u) != 0;
        }
        intercell_flux<EQ, SOLVER, FKJ ? AV_FKJ98 : AV_NONE>(eL, eR, a.pp, use_hll, 0.0, Fnew);
#pragma unroll
        for (int q = 0; q < NTR; q++) Ftr[q] = tracer_upwind_flux(a, trL[q], trR[q], Fnew.rho);
      } else if (f == 1) {
        mbar_wait_spin(bar_new, par_new);
      }
      Cons D;
      if (f == 0) {
        if (light) {  // publish the x-edge fluxes of plane k+1 (covered by this iteration's arrive on s_bar)
          stress_delay(6);
          if (lane < TY - 1 && !last) {
            // three buffers: this store runs ahead of the wait below, i.e. possibly while slower warps
            // still read the x-edge fluxes of plane k
            double* xe = s_xedge + (size_t)((kk + 3) % 3) * XSLAB + lane;
            xe[0] = Fnew.rho; xe[TY] = Fnew.erg; xe[2 * TY] = Fnew.mn; xe[3 * TY] = Fnew.mt1; xe[4 * TY] = Fnew.mt2;
            if (EQ != EQ_EULER) { xe[5 * TY] = Fnew.bbn; xe[6 * TY] = Fnew.bbt1; xe[7 * TY] = Fnew.bbt2; }
            if (EQ == EQ_GLM) xe[8 * TY] = Fnew.psi;
#pragma unroll
            for (int q = 0; q < NTR; q++) xe[(NB + q) * TY] = Ftr[q];
          }
          if (warm) continue;
        }
        // the previous iteration published this plane's y fluxes and x-edge fluxes
        stress_delay(1);
        mbar_wait_spin(&s_bar, phase);
        phase ^= 1u;
        stress_delay(2);
        Cons Fh = cons_shfl_down<EQ>(Fnew);
        double Fhtr[NTRA];
#pragma unroll
        for (int q = 0; q < NTR; q++) Fhtr[q] = __shfl_down_sync(0xffffffffu, Ftr[q], 1);
        if (lane == 31) {  // high x face of the tile's last column: from the light warp
          const double* xe = s_xedge + (size_t)((kk + 2) % 3) * XSLAB + min(row, TY - 2);
          Fh.rho = xe[0]; Fh.erg = xe[TY]; Fh.mn = xe[2 * TY]; Fh.mt1 = xe[3 * TY]; Fh.mt2 = xe[4 * TY];
          if (EQ != EQ_EULER) { Fh.bbn = xe[5 * TY]; Fh.bbt1 = xe[6 * TY]; Fh.bbt2 = xe[7 * TY]; }
          if (EQ == EQ_GLM) Fh.psi = xe[8 * TY];
#pragma unroll
          for (int q = 0; q < NTR; q++) Fhtr[q] = xe[(NB + q) * TY];
        }
        cons_diff(D, Fnew, Fh);
        const Prim C = lds_prim<EQ, VS>(p0, 0, 1, 2);
        const double uB = (EQ != EQ_EULER) ? C.bn * C.vn + C.bt1 * C.vt1 + C.bt2 * C.vt2 : 0.0;
        acc_sources<EQ, VS, 0>(acc, C, uB, p0 - 1, p0 + 1, dt, idx, hdtdx);
        acc_flux_diff<EQ, 0>(acc, D, dt, idx, dtdx);
#ifdef PION_STRICT
#define PION_ACCTR(q, lo, hi) acctr[q] += dt * (((lo) - (hi)) * idx);
#else
#define PION_ACCTR(q, lo, hi) acctr[q] = fma(dtdx, (lo) - (hi), acctr[q]);
#endif
#pragma unroll
        for (int q = 0; q < NTR; q++) PION_ACCTR(q, Ftr[q], Fhtr[q])
        // y: both faces come from the slab
        const int rn = min(row + 1, TY - 1);
        const Cons Fl = cons_from_smem<EQ, TY>(sbuf, row, lane);
        const Cons Fhy = cons_from_smem<EQ, TY>(sbuf, rn, lane);
        // the tracer fluxes are read BEFORE this thread says it is done with the slab (they used to be read after the
        // arrive: a warp already waiting to publish plane k+1 could overwrite them -- seen as a 1e-2 tracer error of
        // the slower -DPION_STRICT build, r02t)
        double Fltr[NTRA], Fhytr[NTRA];
#pragma unroll
        for (int q = 0; q < NTR; q++) {
          Fltr[q] = sbuf[(NB + q) * CS + row * 32 + lane];
          Fhytr[q] = sbuf[(NB + q) * CS + rn * 32 + lane];
        }
        mbar_arrive(&s_free);
        stress_delay(3);
        cons_diff(D, Fl, Fhy);
        acc_sources<EQ, VS, 1>(acc, C, uB, p0 - CW, p0 + CW, dt, idx, hdtdx);
        acc_flux_diff<EQ, 1>(acc, D, dt, idx, dtdx);
#pragma unroll
        for (int q = 0; q < NTR; q++) PION_ACCTR(q, Fltr[q], Fhytr[q])
      } else if (f == 1) {
        {  // the flux through this cell's low z face was computed one plane ago: thread-private slot in shared memory
          const Cons Fz = cons_from_smem<EQ, TY>(s_fz, row, lane);
          cons_diff(D, Fz, Fnew);
          cons_to_smem<EQ, TY>(s_fz, row, lane, Fnew);
#pragma unroll
          for (int q = 0; q < NTR; q++) {
            double* zt = s_fz + (NB + q) * CS + row * 32 + lane;
            if (!warm) PION_ACCTR(q, *zt, Ftr[q])
            *zt = Ftr[q];
          }
        }
        if (!warm) {
          const Prim C = lds_prim<EQ, VS>(p0, 0, 1, 2);
          const double uB = (EQ != EQ_EULER) ? C.bn * C.vn + C.bt1 * C.vt1 + C.bt2 * C.vt2 : 0.0;
          acc_sources<EQ, VS, 2>(acc, C, uB, pm1, pp1, dt, idx, hdtdx);
          acc_flux_diff<EQ, 2>(acc, D, dt, idx, dtdx);
        }
        // plane k-1 (z flux Q0, z sources) has been read for the last time by this warp; the LAST consumer
        // warp to get here refills that buffer with plane k+3 (needed by the next iteration's z flux), so
        // nobody spins on an "empty" barrier.  A warp cannot be a whole iteration ahead (s_bar), so the
        // running count identifies the iteration.
        stress_delay(5);
        __syncwarp();
        if (lane == 0 && !light) {
          __threadfence_block();
          const unsigned old = atomicAdd(&s_done, 1u);
          // (three-plane ring: the warm iteration has no plane k-1 to replace, plane k0+1 came with the prologue)
          if (!last && (RING == 4 || !warm) && old == (unsigned)(kk + 2) * (TY - 1) - 1u) {
            const int bdead = (int)((unsigned)(kk - 1 + BASE) % RING);
            unsigned long long* fb = &s_full[bdead];
            mbar_expect_tx(fb, PLANE_BYTES);
            tma_load_plane(s_tile + bdead * PS, &tmap, fb, bx, by, bz + kk + RING + 1);
          }
        }
      } else {
        double* nbuf = s_flux;  // the slab now takes plane k+1 ...
        stress_delay(4);
        if (!warm) {            // ... once everybody has read plane k out of it
          mbar_wait_spin(&s_free, fphase);
          fphase ^= 1u;
        }
        cons_to_smem<EQ, TY>(nbuf, row, lane, Fnew);
#pragma unroll
        for (int q = 0; q < NTR; q++) nbuf[(NB + q) * CS + row * 32 + lane] = Ftr[q];
        mbar_arrive(&s_bar);
      }
    }
    if (SOLVER == SOLVE_HLLD) {
      w_k = w_k1; w_k1 = n_w; we_k1 = n_we;
      hp += g.sz;
      if (light) hpe += g.sz;
    }
    if (warm) continue;
    if (a.mask) { m_k = n_m; mp += g.sz; }

    if (domain) {
      Cons accx;  // grid frame == solver frame of x
      accx.rho = acc.rho; accx.erg = acc.erg; accx.mn = acc.m0; accx.mt1 = acc.m1; accx.mt2 = acc.m2;
      accx.bbn = acc.b0; accx.bbt1 = acc.b1; accx.bbt2 = acc.b2; accx.psi = acc.psi;
      if (pb_is_s) {
        status |= cell_advance_time_pb<EQ>(a, c, lds_prim<EQ, VS>(p0, 0, 1, 2), accx, acctr, NTR, my_dt, p0 + NB * VS, VS);
      } else if (pb_staged) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        const double* q = s_pb + row * 32 + lane;  // own slot: variable v at q[v (TY-1) 32]
        status |= cell_advance_time_pb<EQ>(a, c, lds_prim<EQ, (TY - 1) * 32>(q, 0, 1, 2), accx, acctr, NTR, my_dt, q + NB * (TY - 1) * 32, (TY - 1) * 32);
      } else {
        status |= cell_advance_time<EQ>(a, c, accx, acctr, NTR, my_dt);
      }
    } else if (upd_xy && a.out != a.S) {
      // cell cut out of the domain (time_integrator.cpp:905-908): state untouched
      for (int v = 0; v < NV; v++) a.out[(long)v * vs + c] = a.S[(long)v * vs + c];
    }
  }

#undef PION_ACCTR
#undef dt
#undef idx
#undef dtdx
#undef hdtdx
  stage_block_epilogue(a, my_dt, status);
}

__host__ __device__ constexpr bool tma_fits(int eq, int ntr) {
  return ntr <= TMA_MAXTR && tma_smem_bytes(nbase(eq) + ntr, tma_ty(eq, 1, ntr), tma_ring(eq, 1)) <= tma_smem_budget(sweep_minb(eq)) &&
         tma_smem_bytes(nbase(eq) + ntr, tma_ty(eq, 2, ntr), tma_ring(eq, 2)) <= tma_smem_budget(sweep_minb(eq));
}

template <int EQ, int SOLVER, bool FKJ, int NTR, int ORDER>
inline const char* launch_sweep_tma_o(const StageArgs& a, cudaStream_t s) {
  constexpr int TY = tma_ty(EQ, ORDER, NTR), RING = tma_ring(EQ, ORDER), MINB = sweep_minb(EQ);
  constexpr int NV = nbase(EQ) + NTR;
  static char name[160], extra[96];
  static const char* nm = (snprintf(extra, sizeof extra, ",TY=%d|%d,RING=%d|%d,NTR=%d,ORDER=1|2 (TMA-staged stencil)", tma_ty(EQ, 1, NTR), tma_ty(EQ, 2, NTR),
                                    tma_ring(EQ, 1), tma_ring(EQ, 2), NTR),
                           kernel_variant_name(name, sizeof name, "k_stage_sweep_tma", EQ, SOLVER, FKJ, extra));
  const int bx = a.tx1 - a.tx0, by = a.ty1 - a.ty0, NZ = a.k_hi - a.k_lo;
  if (a.nbox == 0 && (bx <= 0 || by <= 0 || NZ <= 0)) return nm;
  int kchunk = PION_TMA_KCHUNK;
  dim3 grid;
  StageArgs ab = a;
  if (a.nbox > 0) {  // several boxes, one launch: 1-D grid
    int tot = sweep_fill_box_table(ab, kchunk);
    while (kchunk > 8 && tot < 148 * 4) { kchunk >>= 1; tot = sweep_fill_box_table(ab, kchunk); }
    if (tot <= 0) return nm;
    grid = dim3(tot, 1, 1);
  } else {
    while (kchunk > 8 && (long)bx * by * ((NZ + kchunk - 1) / kchunk) < 148L * 4) kchunk >>= 1;
    grid = dim3(bx, by, (NZ + kchunk - 1) / kchunk);
  }
  constexpr size_t smem = tma_smem_bytes(NV, TY, RING) + (tma_pb_stage(EQ, ORDER, NV, TY) ? 128 + (size_t)NV * (TY - 1) * 32 * sizeof(double) : 0);
  // the opt-in is per DEVICE: one flag per ordinal (a process may hold contexts on several GPUs)
  static bool attr_done[PION_MAX_DEVICES] = {false};
  const int dev = current_device_slot();
  if (!attr_done[dev]) {
    cudaFuncSetAttribute(k_stage_sweep_tma<EQ, SOLVER, FKJ, TY, MINB, NTR, ORDER, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done[dev] = (dev != PION_MAX_DEVICES - 1);  // the overflow slot is never cached
  }
  k_stage_sweep_tma<EQ, SOLVER, FKJ, TY, MINB, NTR, ORDER, RING><<<grid, 32 * TY, smem, s>>>(ab, *reinterpret_cast<const CUtensorMap*>(a.tmap), kchunk);
  return nm;
}
template <int EQ, int SOLVER, bool FKJ, int NTR>
inline const char* launch_sweep_tma_t(const StageArgs& a, cudaStream_t s) {
  if (a.order == 2) return launch_sweep_tma_o<EQ, SOLVER, FKJ, NTR, 2>(a, s);
  return launch_sweep_tma_o<EQ, SOLVER, FKJ, NTR, 1>(a, s);
}

// 3-D grids with at most TMA_MAXTR tracers and no H-correction run the TMA kernel when its tile fits the
// shared memory, everything else the LDG sweep kernel
template <int EQ, int SOLVER, bool FKJ>
inline const char* launch_sweep_any(const StageArgs& a, cudaStream_t s) {
  const bool tma = a.tmap && a.g.ndim == 3 && !a.eta && (SOLVER != SOLVE_HLLD || a.hllf) && tma_fits(EQ, a.ntr);
  if (tma && a.ntr == 0) return launch_sweep_tma_t<EQ, SOLVER, FKJ, 0>(a, s);
  if constexpr (tma_fits(EQ, 1)) {
    if (tma && a.ntr == 1) return launch_sweep_tma_t<EQ, SOLVER, FKJ, 1>(a, s);
  }
  if constexpr (tma_fits(EQ, 2)) {
    if (tma && a.ntr == 2) return launch_sweep_tma_t<EQ, SOLVER, FKJ, 2>(a, s);
  }
  return launch_sweep_t<EQ, SOLVER, FKJ>(a, s);
}

// box of one TMA plane load for an equation set (host side: tensor-map creation)
inline bool sweep_tma_fits_impl(int eq, int ntr) { return tma_fits(eq, ntr); }
inline void sweep_tma_box_impl(int eq, int order, int ntr, int* cw, int* rh, int* nb, int* tx, int* ty) {
  *tx = TMA_TX;
  *ty = tma_ty(eq, order, ntr) - 1;
  *cw = TMA_CW;
  *rh = tma_rh(tma_ty(eq, order, ntr));
  *nb = nbase(eq);
}

}  // namespace pion
