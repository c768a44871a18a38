// pion_b200/csrc/stage_sweep.cuh -- the production stage kernel: one predictor
// or corrector stage of the finite-volume update with every interface flux
// computed exactly ONCE.
//
// Same reference path as stage_kernel.cuh (time_integrator.cpp:498-958,
// VectorOps.cpp:535-644, solver_eqn_base.cpp:152-342, solver_eqn_mhd_adi.cpp:
// 368-443,782-844), different work decomposition:
//
//   * a thread block is a 32 x TY tile of the xy plane that MARCHES in z over a
//     chunk of planes; warp w is row j0+w, lane l is cell i0+l;
//   * every thread computes the flux through the LOW x face and the LOW y face
//     of its cell and the HIGH z face (= low face of the cell above);
//       x: the high-face flux is the next lane's low-face flux  -> __shfl_down
//       y: the high-face flux is the next warp's low-face flux  -> shared memory
//       z: the low-face flux is the one this thread computed one plane earlier
//          -> stays in registers
//     so lane 31 and row TY-1 only produce fluxes for their neighbours (the
//     extra row computes nothing but its y flux): 31 x (TY-1) cells are updated
//     per plane by 32 x TY threads, and a chunk pays one extra z-flux plane;
//   * contributions are accumulated in the reference's order (x, y, z; sources
//     before the flux difference), then CellAdvanceTime is applied in registers:
//     dU, slopes, edge states and fluxes never touch HBM.
//
// The axis loop is a real loop with ONE inlined Riemann solver (the solver frame
// is rotated in registers), which keeps the kernel inside the instruction cache.
#pragma once
#include "stage_kernel.cuh"

namespace pion {

template <int EQ>
__device__ __forceinline__ void cons_zero(Cons& f) {
  f.rho = f.erg = f.mn = f.mt1 = f.mt2 = f.bbn = f.bbt1 = f.bbt2 = f.psi = 0.0;
}

template <int EQ>
__device__ __forceinline__ Cons cons_shfl_down(const Cons& f) {
  Cons r;
  const unsigned m = 0xffffffffu;
  r.rho = __shfl_down_sync(m, f.rho, 1);
  r.erg = __shfl_down_sync(m, f.erg, 1);
  r.mn = __shfl_down_sync(m, f.mn, 1);
  r.mt1 = __shfl_down_sync(m, f.mt1, 1);
  r.mt2 = __shfl_down_sync(m, f.mt2, 1);
  if (EQ != EQ_EULER) {
    r.bbn = __shfl_down_sync(m, f.bbn, 1);
    r.bbt1 = __shfl_down_sync(m, f.bbt1, 1);
    r.bbt2 = __shfl_down_sync(m, f.bbt2, 1);
  } else {
    r.bbn = r.bbt1 = r.bbt2 = 0.0;
  }
  r.psi = (EQ == EQ_GLM) ? __shfl_down_sync(m, f.psi, 1) : 0.0;
  return r;
}

// shared-memory flux slab of one plane: [component][row][lane]
template <int EQ, int TY>
__device__ __forceinline__ void cons_to_smem(double* s, int row, int lane, const Cons& f) {
  double* p = s + row * 32 + lane;
  constexpr int CS = TY * 32;
  p[0] = f.rho; p[CS] = f.erg; p[2 * CS] = f.mn; p[3 * CS] = f.mt1; p[4 * CS] = f.mt2;
  if (EQ != EQ_EULER) { p[5 * CS] = f.bbn; p[6 * CS] = f.bbt1; p[7 * CS] = f.bbt2; }
  if (EQ == EQ_GLM) p[8 * CS] = f.psi;
}
template <int EQ, int TY>
__device__ __forceinline__ Cons cons_from_smem(const double* s, int row, int lane) {
  const double* p = s + row * 32 + lane;
  constexpr int CS = TY * 32;
  Cons f;
  f.rho = p[0]; f.erg = p[CS]; f.mn = p[2 * CS]; f.mt1 = p[3 * CS]; f.mt2 = p[4 * CS];
  if (EQ != EQ_EULER) { f.bbn = p[5 * CS]; f.bbt1 = p[6 * CS]; f.bbt2 = p[7 * CS]; } else { f.bbn = f.bbt1 = f.bbt2 = 0.0; }
  f.psi = (EQ == EQ_GLM) ? p[8 * CS] : 0.0;
  return f;
}

// Flux through the LOW face of cell X along the axis with stride `st`, in that
// axis' solver frame: slopes + edge states (VectorOps.cpp:535-617), HLLD->HLL
// switch (solver_eqn_mhd_adi.cpp:167-177), H-correction eta (solver_eqn_base.cpp:
// 608-678) and InterCellFlux.
template <int EQ, int SOLVER, bool FKJ>
__device__ __forceinline__ void low_face_flux(const StageArgs& a, long X, long st, int ax, int a1, int a2, bool has_m2, Cons& F) {
  const GridD& g = a.g;
  const long vs = g.vs;
  Prim eL = load_prim<EQ>(a.S, X - st, vs, ax, a1, a2);
  Prim eR = load_prim<EQ>(a.S, X, vs, ax, a1, a2);
  if (a.order == 2) {
    const Prim Q0 = load_prim<EQ>(a.S, X - 2 * st, vs, ax, a1, a2);
    const Prim Q3 = load_prim<EQ>(a.S, X + st, vs, ax, a1, a2);
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
  bool use_hll = false;
  if (SOLVER == SOLVE_HLLD) use_hll = (a.hll[X - st] | a.hll[X]) != 0;
  double eta = 0.0;
  if (SOLVER == SOLVE_ROE && a.eta) {
    const double* en = a.eta + (long)ax * vs;
    eta = en[X - st];
    if (g.ndim > 1) {
      const double* e1 = a.eta + (long)((ax + 1) % g.ndim) * vs;
      // the cell two below along the sweep axis does not exist next to a one-deep ghost frame (first-order
      // grids): the reference skips it (solver_eqn_base.cpp:659-676, NextPt == 0)
      eta = fmax(eta, fmax(e1[X - st], e1[X]));
      if (has_m2) eta = fmax(eta, e1[X - 2 * st]);
    }
    if (g.ndim > 2) {
      const double* e2 = a.eta + (long)((ax + 2) % g.ndim) * vs;
      eta = fmax(eta, fmax(e2[X - st], e2[X]));
      if (has_m2) eta = fmax(eta, e2[X - 2 * st]);
    }
  }
  intercell_flux<EQ, SOLVER, FKJ ? AV_FKJ98 : AV_NONE>(eL, eR, a.pp, use_hll, eta, F);
}

// Upwinded tracer flux through the low face of cell X (solver_eqn_base.cpp:281-342)
__device__ __forceinline__ double tracer_low_face_flux(const StageArgs& a, const double* __restrict__ T, long X, long st,
                                                       double Frho) {
  double L = __ldg(T + X - st), R = __ldg(T + X);
  if (a.order == 2) {
    const double q0 = __ldg(T + X - 2 * st), q3 = __ldg(T + X + st);
    const double d0 = L - q0, d1 = R - L, d2 = q3 - R;
    L += minmod(d0, d1, a.tiny2) * 0.5;
    R -= minmod(d1, d2, a.tiny2) * 0.5;
  }
  double f = 0.0;
  if (Frho > 0.0) f = L * Frho * (a.pp.have_mp ? scma_corr(L) : 1.0);
  else if (Frho < 0.0) f = R * Frho * (a.pp.have_mp ? scma_corr(R) : 1.0);
  return f;
}

#ifndef PION_SWEEP_MBAR
#define PION_SWEEP_MBAR 1
#endif
#ifndef PION_SWEEP_MINBLOCKS
#define PION_SWEEP_MINBLOCKS 1
#endif
#ifndef PION_SWEEP_TY
#define PION_SWEEP_TY 12
#endif
#ifndef PION_SWEEP_TY_EULER
#define PION_SWEEP_TY_EULER 8
#endif
#ifndef PION_SWEEP_MINBLOCKS_EULER
#define PION_SWEEP_MINBLOCKS_EULER 2
#endif
// Tile shape per equation set.  MHD (HLLD / Roe) needs ~166 registers: one 384-thread block per SM
// without spills beats two 256-thread blocks at 128 registers with spills.  Euler is lighter and
// latency-bound (ncu: long-scoreboard 39 %), so it runs more, smaller blocks per SM.
__host__ __device__ constexpr int sweep_ty(int eq) { return eq == EQ_EULER ? PION_SWEEP_TY_EULER : PION_SWEEP_TY; }
__host__ __device__ constexpr int sweep_minb(int eq) { return eq == EQ_EULER ? PION_SWEEP_MINBLOCKS_EULER : PION_SWEEP_MINBLOCKS; }

// split-phase block barrier (mbarrier): a thread ARRIVES right after publishing its y flux
// and only WAITS after it has computed its x flux, so warp skew hides behind useful work.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PION_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra PION_MBAR_WAIT;\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

template <int EQ, int SOLVER, bool FKJ, int TY, int MINB, bool TR>
__global__ void __launch_bounds__(32 * TY, MINB) k_stage_sweep(const __grid_constant__ StageArgs a, const int kchunk) {
  extern __shared__ double s_flux[];  // [2][NB + MAXTR][TY][32]
  __shared__ unsigned long long s_bar;
  constexpr int NB = nbase(EQ);
  constexpr int SLAB = (NB + (TR ? PION_MAXTR : 0)) * TY * 32;  // tracer slabs only in the tracer instantiation
  const GridD& g = a.g;
  const int NX = g.NG[0], NY = g.NG[1];
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const BlockBox bb = sweep_block_box(a);
  int i = bb.tx * 31 + lane, j = bb.ty * (TY - 1) + row;
  const bool row_active = (row < TY - 1) && (j < NY);               // warp-uniform
  const bool upd_xy = row_active && (lane < 31) && (i < NX);
  i = min(i, NX);  // clamped threads recompute a neighbour's (valid) face; their results are never used
  j = min(j, NY);
  const int k0 = bb.k_lo + bb.tz * kchunk, k1 = min(k0 + kchunk, bb.k_hi);
  const bool has_z = g.ndim > 2;
  const long vs = g.vs;
  const double idx = 1.0 / g.dx;
  const double dt = a.dt;
  const int ntr = TR ? a.ntr : 0;  // TR=false instantiation: no tracer registers at all
  double my_dt = 1.0e100;
  int status = 0;
  if (threadIdx.x == 0) mbar_init(&s_bar, 32 * TY);
  __syncthreads();
  unsigned phase = 0;

  Cons Fz;  // flux through the low z face of the current cell
  cons_zero<EQ>(Fz);
  double Fz_tr[PION_MAXTR];
#pragma unroll
  for (int q = 0; q < PION_MAXTR; q++) Fz_tr[q] = 0.0;

  for (int k = k0 - 1; k < k1; k++) {
    const bool warm = k < k0;  // first iteration of a chunk: only the fluxes INTO plane k0 (z face, y faces)
    const long c = gidx(g, i + g.nb[0], j + g.nb[1], k + g.nb[2]);
    double* sbuf = s_flux + (size_t)((k - k0) & 1) * SLAB;
    double* sbuf_tr = sbuf + NB * TY * 32;


    Cons acc;
    cons_zero<EQ>(acc);
    double acctr[PION_MAXTR];
#pragma unroll
    for (int q = 0; q < PION_MAXTR; q++) acctr[q] = 0.0;
    Prim C;
    const bool domain = upd_xy && !warm && (a.mask ? (a.mask[c] != 0) : true);
    if (!warm) {
      C = load_prim<EQ>(a.S, c, vs, 0, 1, 2);
      if (a.mp_dE && upd_xy) acc.erg = a.mp_dE[c];  // cooling source term (energy only)
    }

    // Schedule of one plane (ONE flux call site, ONE accumulate site):
    //   step 1: x flux (shfl), accumulate x          step 2: wait, accumulate y from shared memory
    //   step 3: z flux, accumulate z                 step 4: y flux of the NEXT plane -> shared memory, arrive
    // so dU is still summed in the reference's order x, y, z, and a whole plane of work separates the
    // arrive (end of the previous iteration) from the wait (step 2): warp skew never reaches the barrier.
    // The warm-up iteration does steps 3 and 4 only (the fluxes into the chunk's first plane).
#pragma unroll 1
    for (int step = warm ? 3 : 1; step <= 4; step++) {
      if ((step == 3 && !has_z) || (step == 4 && k + 1 >= k1)) continue;
      const int ax = (step == 1) ? 0 : (step == 3) ? 2 : 1;
      const int a1 = (ax == 2) ? 0 : ax + 1;
      const int a2 = (a1 == 2) ? 0 : a1 + 1;
      const long st = axis_stride(g, ax);
      // z: the HIGH face of this cell = low face of the cell above; step 4: the y face of the cell above
      const long X = (step >= 3) ? c + g.sz : c;
      Cons Fnew;
      cons_zero<EQ>(Fnew);
      double Fnew_tr[PION_MAXTR];
#pragma unroll
      for (int q = 0; q < PION_MAXTR; q++) Fnew_tr[q] = 0.0;
      if (step != 2 && (row_active || ax == 1)) {
        // padded index of X along the sweep axis (the H-correction stencil reaches two cells below it)
        const int qX = (step == 1) ? i + g.nb[0] : (step == 3) ? k + 1 + g.nb[2] : j + g.nb[1];
        low_face_flux<EQ, SOLVER, FKJ>(a, X, st, ax, a1, a2, qX >= 2, Fnew);
#pragma unroll
        for (int q = 0; q < PION_MAXTR; q++)
          if (q < ntr) Fnew_tr[q] = tracer_low_face_flux(a, a.S + (long)(NB + q) * vs, X, st, Fnew.rho);
      }
      if (step == 4) {
        double* nbuf = s_flux + (size_t)((k + 1 - k0) & 1) * SLAB;  // slab of plane k+1
        double* nbuf_tr = nbuf + NB * TY * 32;
        cons_to_smem<EQ, TY>(nbuf, row, lane, Fnew);
#pragma unroll
        for (int q = 0; q < PION_MAXTR; q++)
          if (q < ntr) nbuf_tr[(q * TY + row) * 32 + lane] = Fnew_tr[q];
#if PION_SWEEP_MBAR
        mbar_arrive(&s_bar);
#endif
        continue;
      }

      // D = F(low face) - F(high face) of this cell along the axis, formed directly from where the two
      // fluxes live (registers + shuffle / shared memory / the previous plane's registers)
      Cons D;
      double D_tr[PION_MAXTR];
#define PION_DIFF(LOW, HIGH)                                                                     \
  D.rho = LOW.rho - HIGH.rho; D.erg = LOW.erg - HIGH.erg; D.mn = LOW.mn - HIGH.mn;               \
  D.mt1 = LOW.mt1 - HIGH.mt1; D.mt2 = LOW.mt2 - HIGH.mt2;                                        \
  if (EQ != EQ_EULER) { D.bbn = LOW.bbn - HIGH.bbn; D.bbt1 = LOW.bbt1 - HIGH.bbt1; D.bbt2 = LOW.bbt2 - HIGH.bbt2; } \
  else { D.bbn = D.bbt1 = D.bbt2 = 0.0; }                                                        \
  D.psi = (EQ == EQ_GLM) ? LOW.psi - HIGH.psi : 0.0;
      if (step == 1) {
        const Cons Fh = cons_shfl_down<EQ>(Fnew);
        PION_DIFF(Fnew, Fh)
#pragma unroll
        for (int q = 0; q < PION_MAXTR; q++)
          D_tr[q] = (q < ntr) ? Fnew_tr[q] - __shfl_down_sync(0xffffffffu, Fnew_tr[q], 1) : 0.0;
      } else if (step == 2) {
#if PION_SWEEP_MBAR
        mbar_wait(&s_bar, phase);
        phase ^= 1u;
#else
        __syncthreads();
#endif
        const int rn = min(row + 1, TY - 1);
        const Cons Fl = cons_from_smem<EQ, TY>(sbuf, row, lane);
        const Cons Fh = cons_from_smem<EQ, TY>(sbuf, rn, lane);
        PION_DIFF(Fl, Fh)
#pragma unroll
        for (int q = 0; q < PION_MAXTR; q++)
          D_tr[q] = (q < ntr) ? sbuf_tr[(q * TY + row) * 32 + lane] - sbuf_tr[(q * TY + rn) * 32 + lane] : 0.0;
      } else {
        PION_DIFF(Fz, Fnew)
        Fz = Fnew;
#pragma unroll
        for (int q = 0; q < PION_MAXTR; q++) {
          D_tr[q] = Fz_tr[q] - Fnew_tr[q];
          Fz_tr[q] = Fnew_tr[q];
        }
      }
#undef PION_DIFF

      if (!warm) {
        // Powell + GLM sources from cell-centre states (solver_eqn_mhd_adi.cpp:396-443,782-813):
        // R part of interface (i-1,i) first, then L part of interface (i,i+1)
        if (EQ != EQ_EULER) {
          const double* Bn = a.S + (long)(5 + ax) * vs;
          const double bm = __ldg(Bn + c - st), bp = __ldg(Bn + c + st);
          const double uB = C.bn * C.vn + C.bt1 * C.vt1 + C.bt2 * C.vt2;
          double f = dt * (0.5 * (bm + C.bn));
          acc.mn += f * C.bn * idx; acc.mt1 += f * C.bt1 * idx; acc.mt2 += f * C.bt2 * idx; acc.erg += f * uB * idx;
          acc.bbn += f * C.vn * idx; acc.bbt1 += f * C.vt1 * idx; acc.bbt2 += f * C.vt2 * idx;
          double psm = 0.0, psp = 0.0;
          if (EQ == EQ_GLM) {
            psm = __ldg(a.S + 8 * vs + c - st);
            psp = __ldg(a.S + 8 * vs + c + st);
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
        }
        // flux difference (dU_Cell + DivStateVectorComponent)
        acc.rho += dt * (D.rho * idx);
        acc.erg += dt * (D.erg * idx);
        acc.mn += dt * (D.mn * idx);
        acc.mt1 += dt * (D.mt1 * idx);
        acc.mt2 += dt * (D.mt2 * idx);
        if (EQ != EQ_EULER) {
          acc.bbn += dt * (D.bbn * idx);
          acc.bbt1 += dt * (D.bbt1 * idx);
          acc.bbt2 += dt * (D.bbt2 * idx);
        }
        if (EQ == EQ_GLM) acc.psi += dt * (D.psi * idx);
#pragma unroll
        for (int q = 0; q < PION_MAXTR; q++)
          if (q < ntr) acctr[q] += dt * (D_tr[q] * idx);
        // rotate the centre state and the accumulators into the next axis' frame
        rot3(C.vn, C.vt1, C.vt2);
        rot3(acc.mn, acc.mt1, acc.mt2);
        if (EQ != EQ_EULER) {
          rot3(C.bn, C.bt1, C.bt2);
          rot3(acc.bbn, acc.bbt1, acc.bbt2);
        }
      }
    }
    if (warm) continue;
    if (g.ndim == 2) {  // frame is (z,x,y): one more rotation returns to x
      rot3(acc.mn, acc.mt1, acc.mt2);
      if (EQ != EQ_EULER) rot3(acc.bbn, acc.bbt1, acc.bbt2);
    }

    if (domain) {
      status |= cell_advance_time<EQ>(a, c, acc, acctr, ntr, my_dt);
    } else if (upd_xy && a.out != a.S) {
      // cell cut out of the domain (time_integrator.cpp:905-908): state untouched
      for (int v = 0; v < NB + ntr; v++) a.out[(long)v * vs + c] = a.S[(long)v * vs + c];
    }
  }

  stage_block_epilogue(a, my_dt, status);
}

// slot of the current device in the per-device "attribute set" flags of the launchers (the last slot is shared
// by every ordinal >= PION_MAX_DEVICES - 1 and never cached)
constexpr int PION_MAX_DEVICES = 65;
inline int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < PION_MAX_DEVICES - 1) ? dev : PION_MAX_DEVICES - 1;
}

template <int EQ, int SOLVER, bool FKJ>
inline const char* launch_sweep_t(const StageArgs& a, cudaStream_t s) {
  constexpr int TY = sweep_ty(EQ), MINB = sweep_minb(EQ);
  constexpr int NB = nbase(EQ);
  const int bx = a.tx1 - a.tx0, by = a.ty1 - a.ty0, NZ = a.k_hi - a.k_lo;
  static char name[2][112];
  static const char* nm[2] = {
      kernel_variant_name(name[0], sizeof name[0], "k_stage_sweep", EQ, SOLVER, FKJ, ",TR=0 (LDG stencil)"),
      kernel_variant_name(name[1], sizeof name[1], "k_stage_sweep", EQ, SOLVER, FKJ, ",TR=1 (LDG stencil)")};
  if (a.nbox == 0 && (bx <= 0 || by <= 0 || NZ <= 0)) return nm[a.ntr > 0];
  // z chunks: enough blocks to fill 148 SMs a few times over, long enough to amortise the extra flux plane
  int kchunk = NZ;
  dim3 grid;
  StageArgs ab = a;
  if (a.nbox > 0) {  // several boxes, one launch (3-D only): 1-D grid
    kchunk = 64;
    int tot = sweep_fill_box_table(ab, kchunk);
    while (kchunk > 8 && tot < 148 * 4) { kchunk >>= 1; tot = sweep_fill_box_table(ab, kchunk); }
    if (tot <= 0) return nm[a.ntr > 0];
    grid = dim3(tot, 1, 1);
  } else {
    if (a.g.ndim > 2) {
      kchunk = 64;
      while (kchunk > 8 && (long)bx * by * ((NZ + kchunk - 1) / kchunk) < 148L * 4) kchunk >>= 1;
    }
    grid = dim3(bx, by, (NZ + kchunk - 1) / kchunk);
  }
  const size_t smem = (size_t)2 * (NB + PION_MAXTR) * TY * 32 * sizeof(double);
  const size_t smem_notr = (size_t)2 * NB * TY * 32 * sizeof(double);
  // the opt-in is per DEVICE: one flag per ordinal (a process may hold contexts on several GPUs)
  static bool attr_done[PION_MAX_DEVICES] = {false};
  const int dev = current_device_slot();
  if (!attr_done[dev]) {
    cudaFuncSetAttribute(k_stage_sweep<EQ, SOLVER, FKJ, TY, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_stage_sweep<EQ, SOLVER, FKJ, TY, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_notr);
    attr_done[dev] = (dev != PION_MAX_DEVICES - 1);  // the overflow slot is never cached
  }
  if (a.ntr > 0) k_stage_sweep<EQ, SOLVER, FKJ, TY, MINB, true><<<grid, 32 * TY, smem, s>>>(ab, kchunk);
  else k_stage_sweep<EQ, SOLVER, FKJ, TY, MINB, false><<<grid, 32 * TY, smem_notr, s>>>(ab, kchunk);
  return nm[a.ntr > 0];
}

// number of cells a sweep tile updates along x and y (host side: shell / interior boxes)
inline void sweep_tile_cells_impl(int eq, int* cx, int* cy) { *cx = 31; *cy = sweep_ty(eq) - 1; }

}  // namespace pion
