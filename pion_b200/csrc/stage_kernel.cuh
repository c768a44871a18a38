// pion_b200/csrc/stage_kernel.cuh -- one predictor or corrector stage of the
// finite-volume update as ONE kernel (gather form).
//
// Reference path restated (paths relative to /root/reference/source):
//   sim_control/time_integrator.cpp:498-873  calc_dynamics_dU / set_dynamics_dU /
//                                            dynamics_dU_column (per-axis column sweeps
//                                            scattering into dU)
//   coord_sys/VectorOps.cpp:535-644          SetEdgeState, SetSlope, DivStateVectorComponent
//   spatial_solvers/solver_eqn_base.cpp:152-342  InterCellFlux, tracer flux, select_Hcorr_eta
//   spatial_solvers/solver_eqn_mhd_adi.cpp:368-443,782-844  dU_Cell, Powell + GLM sources,
//                                            CellAdvanceTime (+GLMsource)
//   sim_control/time_integrator.cpp:881-958  grid_update_state_vector
//   sim_control/calc_timestep.cpp:271-333    calc_dynamics_dt (fused into the corrector)
//
// The reference walks 1-D columns and scatters each interface's contribution into
// the two adjacent cells' dU.  Here each thread GATHERS everything that lands in
// its own cell, in the reference's accumulation order
//   dU = [microphysics] ; for axis in x,y,z: + src_R(i-1,i) ; - src_L(i,i+1) ; + dt*(F_{i-1/2}-F_{i+1/2})/dx
// and (fused mode) immediately applies CellAdvanceTime, so dU, slopes, edge states
// and fluxes never touch HBM.  The axis loop is a real loop: the solver frame is
// rotated in registers between axes, so the Riemann solver is instantiated once.
#pragma once
#include <cstdio>
#include "grid.cuh"

namespace pion {

struct StageArgs {
  GridD g;
  PhysParams pp;
  const double* S;    // stencil source: Ph (or P on the fused predictor, where Ph==P)
  const double* Pb;   // base state P of U = PtoU(P) + dU
  double* out;        // fused: destination primitive array (Ph on predictor, P on corrector)
  double* dU;         // unfused: accumulated into; unused by the fused path
  const double* mp_dE;  // fused: microphysics energy source per cell (cooling.cuh), null without microphysics
  const unsigned char* hll;  // HLLD->HLL switch flags per cell (null unless solver==HLLD)
  const unsigned char* hllf; // the same per LOW FACE, bits 0/1/2 = x/y/z (k_hll_face_flags; TMA sweep kernel only)
  const double* eta;         // H-correction eta, 3 planes of vs doubles (null unless AV 3/4)
  const unsigned char* mask; // 1 = cell is updated (isdomain); null = every interior cell
  double dt;          // stage dt == FV_dt
  double idx, dtdx, hdtdx;  // 1/dx, dt/dx, dt/(2 dx): launch constants (kernel-parameter bank operands, no registers)
  double tiny2;       // VERY_TINY_VALUE * dx^2 (minmod cut-off for undivided differences)
  double glm_damp;    // exp(-FV_dt * c_h * c_r)
  double cfl;
  unsigned long long* dtmin;  // corrector: ordered-bits min of the next CFL dt (null = skip)
  long long* counters;        // [0] negative density, [1] negative pressure fix-ups
  int order;          // spatial order of this stage (1 or 2)
  int ntr;            // tracers
  int fused;          // 1: apply CellAdvanceTime and write `out`; 0: dU += ...
  int fkj;            // FKJ98 viscosity on (AV 1 or 4)
  // sweep kernel only: the box of tiles / planes this launch covers (x tiles of 31 cells, y tiles of
  // TY-1 rows, z planes); the whole grid unless the stage is split into boundary shell + interior
  int tx0, tx1, ty0, ty1, k_lo, k_hi;
  // ... or SEVERAL boxes in one launch (the boundary shell of a split stage: six boxes, one kernel): the grid is
  // then 1-D, block b belongs to the box q with box_blk0[q] <= b < box_blk0[q+1] and decodes its tile there
  int nbox;
  int box[6][6];     // tx0, tx1, ty0, ty1, k_lo, k_hi per box
  int box_blk0[7];
  // HOST pointer to the CUtensorMap of array S (TMA sweep kernel; null = not available)
  const void* tmap;
};

// sCMA corrector of one tracer value (microphysics_base.cpp:80-126 with no element
// tracers): the second assignment wins, so only values > 1 are rescaled.
__device__ __forceinline__ double scma_corr(double tr) { return (tr > 1.0) ? fast_rcp(tr) : 1.0; }

// CellTimeStep: Euler solver_eqn_hydro_adi.cpp:460-500, MHD solver_eqn_mhd_adi.cpp:516-574
template <int EQ>
__device__ __forceinline__ double cell_time_step(const Prim& p, const PhysParams& pp, int ndim, double dx, double cfl) {
  double temp;
  if (EQ == EQ_EULER) {
    temp = p.vn * p.vn;
    if (ndim > 1) temp += p.vt1 * p.vt1;
    if (ndim > 2) temp += p.vt2 * p.vt2;
    temp = psqrt(temp) + chydro(p.ro, p.pg, pp.gamma);
  } else {
    temp = fabs(p.vn);
    if (ndim > 1) temp = pmax(fabs(p.vt1), temp);
    if (ndim > 2) temp = pmax(fabs(p.vt2), temp);
    double bx = p.bn, by = p.bt1, bz = p.bt2;
    if (ndim > 1) {
      // rotate to the axis of smallest |B| component (:541-563)
      int newdir = 0;
      if (fabs(p.bt1) < fabs(p.bn)) {
        newdir = 1;
        if (fabs(p.bt2) < fabs(p.bt1)) newdir = 2;
      } else if (fabs(p.bt2) < fabs(p.bn)) newdir = 2;
      if (newdir == 1) { bx = p.bt1; by = p.bt2; bz = p.bn; }
      if (newdir == 2) { bx = p.bt2; by = p.bn; bz = p.bt1; }
    }
    temp += cfast_components(p.ro, p.pg, bx, by, bz, pp.gamma);
  }
  return pdiv(dx, temp) * cfl;
}

__device__ __forceinline__ unsigned long long dbl_ordered_bits(double x) {
  // positive finite doubles order like their bit patterns
  return (unsigned long long)__double_as_longlong(x);
}

// CellAdvanceTime for one cell (solver_eqn_mhd_adi.cpp:452-504, :822-844; Euler
// solver_eqn_hydro_adi.cpp:372-451) + the temperature cap and P=Ph of
// grid_update_state_vector (time_integrator.cpp:881-958): U = PtoU(Pb) + dU,
// out = UtoP(U) with floors; optionally the next step's CellTimeStep.
template <int EQ>
__device__ __forceinline__ int cell_advance_time_pb(const StageArgs& a, long c, const Prim& Pb, const Cons& acc, const double* acctr, int ntr, double& my_dt,
                                                    const double* trb = nullptr, int trb_stride = 0) {
  constexpr int NB = nbase(EQ);
  const long vs = a.g.vs;
  Cons U;
  PtoU<EQ>(Pb, U, a.pp.gamma - 1.0);
  U.rho += acc.rho; U.erg += acc.erg; U.mn += acc.mn; U.mt1 += acc.mt1; U.mt2 += acc.mt2;
  if (EQ != EQ_EULER) { U.bbn += acc.bbn; U.bbt1 += acc.bbt1; U.bbt2 += acc.bbt2; }
  if (EQ == EQ_GLM) U.psi += acc.psi;
  Prim Pn;
  const int status = UtoP<EQ>(U, Pn, a.pp);
  if (EQ == EQ_GLM) Pn.psi *= a.glm_damp;
  // temperature cap of grid_update_state_vector (time_integrator.cpp:926-932)
  // T > Tmax  <=>  p mu/kB > Tmax rho (rho > 0): no division on the common path
  if (a.pp.have_mp && (Pn.pg * a.pp.mu_tot_over_kB > a.pp.max_temp * Pn.ro))
    Pn.pg = Pn.ro * a.pp.max_temp / a.pp.mu_tot_over_kB;
  store_prim<EQ>(a.out, c, vs, Pn);
#pragma unroll
  for (int q = 0; q < PION_MAXTR; q++) {
    if (q < ntr) {
      // base value of the tracer: from the caller's shared-memory tile when the base state IS the stencil state
      // (predictor), else from HBM
      double pb = trb ? trb[q * trb_stride] : __ldg(a.Pb + (long)(NB + q) * vs + c);
      if (a.pp.have_mp) pb *= scma_corr(pb);
      double u = pb * Pb.ro + acctr[q];
      double pn = pdiv(u, U.rho);
      if (a.pp.have_mp) pn *= scma_corr(pn);
      a.out[(long)(NB + q) * vs + c] = pn;
    }
  }
  if (a.dtmin) my_dt = pmin(cell_time_step<EQ>(Pn, a.pp, a.g.ndim, a.g.dx, a.cfl), my_dt);
  return status;
}
// ... with the base state loaded from a.Pb
template <int EQ>
__device__ __forceinline__ int cell_advance_time(const StageArgs& a, long c, const Cons& acc, const double* acctr, int ntr, double& my_dt) {
  const Prim Pb = load_prim<EQ>(a.Pb, c, a.g.vs, 0, 1, 2);
  return cell_advance_time_pb<EQ>(a, c, Pb, acc, acctr, ntr, my_dt);
}

// block-level reductions: min dt (warp shuffles + one atomicMin per block) and error counters
__device__ __forceinline__ void stage_block_epilogue(const StageArgs& a, double my_dt, int status) {
  if (a.dtmin) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_dt = fmin(my_dt, __shfl_xor_sync(0xffffffffu, my_dt, o));
    __shared__ double s_dt[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_dt[w] = my_dt;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = s_dt[0];
      for (int q = 1; q < (int)(blockDim.x >> 5); q++) m = fmin(m, s_dt[q]);
      if (m < 1.0e100) atomicMin(a.dtmin, dbl_ordered_bits(m));
    }
  }
  if (status && a.counters) {
    if (status & ST_NEG_RHO) atomicAdd((unsigned long long*)&a.counters[0], 1ULL);
    if (status & ST_NEG_PG) atomicAdd((unsigned long long*)&a.counters[1], 1ULL);
    if (status & ST_RS_FAIL) atomicAdd((unsigned long long*)&a.counters[3], 1ULL);  // fatal in the reference
  }
}

#ifndef PION_STAGE_MINBLOCKS
#define PION_STAGE_MINBLOCKS 2
#endif
template <int EQ, int SOLVER, bool FKJ>
__global__ void __launch_bounds__(128, PION_STAGE_MINBLOCKS) k_stage(const __grid_constant__ StageArgs a) {
  const GridD& g = a.g;
  const int NX = g.NG[0], NY = g.NG[1];
  const long ncell = (long)NX * NY * g.NG[2];
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < ncell;
  double my_dt = 1.0e100;
  int status = 0, rs_status = 0;

  if (active) {
    const int i = (int)(t % NX), j = (int)((t / NX) % NY), k = (int)(t / ((long)NX * NY));
    const long c = gidx(g, i + g.nb[0], j + g.nb[1], k + g.nb[2]);
    const long vs = g.vs;
    const bool domain = a.mask ? (a.mask[c] != 0) : true;
    constexpr int NB = nbase(EQ);
    const double idx = 1.0 / g.dx;
    const double dt = a.dt;

    Cons acc;
    acc.rho = acc.erg = acc.mn = acc.mt1 = acc.mt2 = acc.bbn = acc.bbt1 = acc.bbt2 = acc.psi = 0.0;
    double acctr[PION_MAXTR];
#pragma unroll
    for (int q = 0; q < PION_MAXTR; q++) acctr[q] = 0.0;

    if (a.fused && a.mp_dE) acc.erg = a.mp_dE[c];  // cooling source term (only the energy component is non-zero)

    Prim C = load_prim<EQ>(a.S, c, vs, 0, 1, 2);

    if (domain || !a.fused) {
#pragma unroll 1
      for (int ax = 0; ax < g.ndim; ax++) {
        const int a1 = (ax == 2) ? 0 : ax + 1;
        const int a2 = (a1 == 2) ? 0 : a1 + 1;
        const long st = axis_stride(g, ax);
        const Prim M1 = load_prim<EQ>(a.S, c - st, vs, ax, a1, a2);
        const Prim P1 = load_prim<EQ>(a.S, c + st, vs, ax, a1, a2);
        // edge states at the low (i-1/2) and high (i+1/2) faces
        Prim lowL = M1, lowR = C, highL = C, highR = P1;
        // curvilinear radial axis (VectorOps_Cyl / VectorOps_Sph): radii of the cell and its faces,
        // slope of the centre cell for the geometric source term
        const bool radial = radial_axis(g, ax);
        const int qax = (ax == 0) ? i + g.nb[0] : (ax == 1) ? j + g.nb[1] : k + g.nb[2];
        const double Rc = radial ? cell_R(g, ax, qax) : 0.0;
        Prim SC;  // slope of cell C (radial axis, second order)
        SC.ro = SC.pg = SC.vn = SC.vt1 = SC.vt2 = SC.bn = SC.bt1 = SC.bt2 = SC.psi = 0.0;
        double rdel[4] = {0, 0, 0, 0}, ri[4] = {0, 0, 0, 0};
        if (a.order == 2) {
          const Prim M2 = load_prim<EQ>(a.S, c - 2 * st, vs, ax, a1, a2);
          const Prim P2 = load_prim<EQ>(a.S, c + 2 * st, vs, ax, a1, a2);
          if (!radial) {
#define PION_EDGE(f)                                                              \
  {                                                                               \
    double d0 = M1.f - M2.f, d1 = C.f - M1.f, d2 = P1.f - C.f, d3 = P2.f - P1.f;  \
    double sm = minmod(d0, d1, a.tiny2), sc = minmod(d1, d2, a.tiny2), sp = minmod(d2, d3, a.tiny2); \
    lowL.f = M1.f + sm * 0.5;                                                     \
    lowR.f = C.f - sc * 0.5;                                                      \
    highL.f = C.f + sc * 0.5;                                                     \
    highR.f = P1.f - sp * 0.5;                                                    \
  }
            PION_EDGE(ro) PION_EDGE(pg) PION_EDGE(vn) PION_EDGE(vt1) PION_EDGE(vt2)
            if (EQ != EQ_EULER) { PION_EDGE(bn) PION_EDGE(bt1) PION_EDGE(bt2) }
            if (EQ == EQ_GLM) { PION_EDGE(psi) }
#undef PION_EDGE
          } else {
            // SetSlope: divided differences between centres of volume; SetEdgeState: distance from the
            // centre of volume to the face (VectorOps.cpp:1073-1079,1158-1182; VectorOps_spherical.cpp:312-380)
            double Rq[5], Rcm[5];
#pragma unroll
            for (int q = 0; q < 5; q++) {
              Rq[q] = cell_R(g, ax, qax - 2 + q);
              Rcm[q] = cell_Rcom(g, Rq[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) ri[q] = 1.0 / (Rcm[q + 1] - Rcm[q]);
            rdel[0] = Rq[1] + 0.5 * g.dx - Rcm[1];  // M1, + face
            rdel[1] = Rq[2] - 0.5 * g.dx - Rcm[2];  // C, - face
            rdel[2] = Rq[2] + 0.5 * g.dx - Rcm[2];  // C, + face
            rdel[3] = Rq[3] - 0.5 * g.dx - Rcm[3];  // P1, - face
#define PION_EDGE_R(f)                                                                                   \
  {                                                                                                      \
    double s0 = (M1.f - M2.f) * ri[0], s1 = (C.f - M1.f) * ri[1], s2 = (P1.f - C.f) * ri[2], s3 = (P2.f - P1.f) * ri[3]; \
    double sm = minmod(s0, s1, PION_VERY_TINY_VALUE), sc = minmod(s1, s2, PION_VERY_TINY_VALUE),         \
           sp = minmod(s2, s3, PION_VERY_TINY_VALUE);                                                    \
    lowL.f = M1.f + sm * rdel[0];                                                                        \
    lowR.f = C.f + sc * rdel[1];                                                                         \
    highL.f = C.f + sc * rdel[2];                                                                        \
    highR.f = P1.f + sp * rdel[3];                                                                       \
    SC.f = sc;                                                                                           \
  }
            PION_EDGE_R(ro) PION_EDGE_R(pg) PION_EDGE_R(vn) PION_EDGE_R(vt1) PION_EDGE_R(vt2)
            if (EQ != EQ_EULER) { PION_EDGE_R(bn) PION_EDGE_R(bt1) PION_EDGE_R(bt2) }
            if (EQ == EQ_GLM) { PION_EDGE_R(psi) }
#undef PION_EDGE_R
          }
        }
        // HLLD -> HLL switch (solver_eqn_mhd_adi.cpp:167-177)
        bool hll_low = false, hll_high = false;
        if (SOLVER == SOLVE_HLLD) {
          unsigned char fm = a.hll[c - st], fc = a.hll[c], fp = a.hll[c + st];
          hll_low = (fm | fc) != 0;
          hll_high = (fc | fp) != 0;
        }
        // H-correction eta_max (solver_eqn_base.cpp:608-678); consumed by Roe only
        double eta_low = 0.0, eta_high = 0.0;
        if (SOLVER == SOLVE_ROE && a.eta) {
          const double* en = a.eta + (long)ax * vs;
          eta_low = en[c - st];
          eta_high = en[c];
          // NB perpendicular axes are taken modulo ndim here (solver_eqn_base.cpp:639,648),
          // unlike the velocity permutation which is always modulo 3
          if (g.ndim > 1) {
            const double* e1 = a.eta + (long)((ax + 1) % g.ndim) * vs;
            // (the cell two below does not exist next to a one-deep ghost frame: skipped like the reference's NextPt == 0)
            eta_low = fmax(eta_low, fmax(e1[c - st], e1[c]));
            if (qax >= 2) eta_low = fmax(eta_low, e1[c - 2 * st]);
            eta_high = fmax(eta_high, fmax(fmax(e1[c], e1[c + st]), e1[c - st]));
          }
          if (g.ndim > 2) {
            const double* e2 = a.eta + (long)((ax + 2) % g.ndim) * vs;
            eta_low = fmax(eta_low, fmax(e2[c - st], e2[c]));
            if (qax >= 2) eta_low = fmax(eta_low, e2[c - 2 * st]);
            eta_high = fmax(eta_high, fmax(fmax(e2[c], e2[c + st]), e2[c - st]));
          }
        }
        Cons Flow, Fhigh;
        int rs_fail = 0;
        intercell_flux<EQ, SOLVER, FKJ ? AV_FKJ98 : AV_NONE>(lowL, lowR, a.pp, hll_low, eta_low, Flow, ax, &rs_fail);
        intercell_flux<EQ, SOLVER, FKJ ? AV_FKJ98 : AV_NONE>(highL, highR, a.pp, hll_high, eta_high, Fhigh, ax, &rs_fail);
        if (rs_fail) rs_status = ST_RS_FAIL;

        // geometric weights of this cell along the axis: Cartesian 1/dx everywhere; radial axis:
        // faces r- = Rc-dx/2, r+ = Rc+dx/2, cyl 2 r/(r+^2 - r-^2), sph r^2/((r+^3 - r-^3)/3)
        double wlow = idx, whigh = idx, wsrc_low = idx, wsrc_high = idx;
        if (radial) {
          const double rp = Rc + g.dx * 0.5, rn = rp - g.dx;
          if (g.coord == 2) {
            const double iv = 1.0 / (rp * rp - rn * rn);
            wlow = 2.0 * rn * iv; whigh = 2.0 * rp * iv;
            wsrc_low = wlow; wsrc_high = whigh;  // cyl MHDsource (solver_eqn_mhd_adi.cpp:1087-1098)
          } else {
            const double iv = 1.0 / ((pow(rp, 3.0) - pow(rn, 3.0)) / 3.0);
            wlow = rn * rn * iv; whigh = rp * rp * iv;
          }
        }
        // Powell + GLM sources from cell-centre states (solver_eqn_mhd_adi.cpp:396-443,782-813):
        // R part of interface (i-1,i) first, then L part of interface (i,i+1)
        if (EQ != EQ_EULER) {
          const double uB = C.bn * C.vn + C.bt1 * C.vt1 + C.bt2 * C.vt2;
          double f = dt * (0.5 * (M1.bn + C.bn));
          acc.mn += f * C.bn * wsrc_low; acc.mt1 += f * C.bt1 * wsrc_low; acc.mt2 += f * C.bt2 * wsrc_low; acc.erg += f * uB * wsrc_low;
          acc.bbn += f * C.vn * wsrc_low; acc.bbt1 += f * C.vt1 * wsrc_low; acc.bbt2 += f * C.vt2 * wsrc_low;
          if (EQ == EQ_GLM) {
            double fs = dt * (0.5 * (M1.psi + C.psi));
            acc.erg += fs * (C.vn * C.psi) * idx;
            acc.psi += fs * C.vn * idx;
          }
          f = dt * (0.5 * (C.bn + P1.bn));
          acc.mn -= f * C.bn * wsrc_high; acc.mt1 -= f * C.bt1 * wsrc_high; acc.mt2 -= f * C.bt2 * wsrc_high; acc.erg -= f * uB * wsrc_high;
          acc.bbn -= f * C.vn * wsrc_high; acc.bbt1 -= f * C.vt1 * wsrc_high; acc.bbt2 -= f * C.vt2 * wsrc_high;
          if (EQ == EQ_GLM) {
            double fs = dt * (0.5 * (C.psi + P1.psi));
            acc.erg -= fs * (C.vn * C.psi) * idx;
            acc.psi -= fs * C.vn * idx;
          }
        }
        // flux difference (dU_Cell + DivStateVectorComponent) + geometric source term
        if (!radial) {
          acc.rho += dt * ((Flow.rho - Fhigh.rho) * idx);
          acc.erg += dt * ((Flow.erg - Fhigh.erg) * idx);
          acc.mn += dt * ((Flow.mn - Fhigh.mn) * idx);
          acc.mt1 += dt * ((Flow.mt1 - Fhigh.mt1) * idx);
          acc.mt2 += dt * ((Flow.mt2 - Fhigh.mt2) * idx);
          if (EQ != EQ_EULER) {
            acc.bbn += dt * ((Flow.bbn - Fhigh.bbn) * idx);
            acc.bbt1 += dt * ((Flow.bbt1 - Fhigh.bbt1) * idx);
            acc.bbt2 += dt * ((Flow.bbt2 - Fhigh.bbt2) * idx);
          }
          if (EQ == EQ_GLM) acc.psi += dt * ((Flow.psi - Fhigh.psi) * idx);
        } else {
          // geometric_source: Euler cyl/sph solver_eqn_hydro_adi.cpp:560-590,648-668; MHD cyl
          // solver_eqn_mhd_adi.cpp:1001-1036; GLM cyl :1156-1190
          double gs_mn = 0.0, gs_bbn = 0.0;
          const double Rcom = cell_Rcom(g, Rc);
          if (g.coord == 2) {
            double ptot = C.pg, dptot = SC.pg;
            if (EQ != EQ_EULER) {
              ptot += (C.bn * C.bn + C.bt1 * C.bt1 + C.bt2 * C.bt2) / 2.;
              dptot = SC.pg + C.bn * SC.bn + C.bt1 * SC.bt1 + C.bt2 * SC.bt2;
            }
            gs_mn = (a.order == 2) ? (ptot + (Rc - Rcom) * dptot) / Rc : ptot / Rc;
            if (EQ == EQ_GLM) gs_bbn = (a.order == 2) ? a.pp.chyp * (C.psi + (Rc - Rcom) * SC.psi) / Rc : a.pp.chyp * C.psi / Rc;
          } else {
            const double R3 = Rc + g.dx * g.dx / 12.0 / Rc;
            gs_mn = (a.order == 2) ? 2.0 * ((C.pg - SC.pg * Rcom) / R3 + SC.pg) : 2.0 * C.pg / R3;
          }
          acc.rho += dt * (wlow * Flow.rho - whigh * Fhigh.rho);
          acc.erg += dt * (wlow * Flow.erg - whigh * Fhigh.erg);
          acc.mn += dt * ((wlow * Flow.mn - whigh * Fhigh.mn) + gs_mn);
          acc.mt1 += dt * (wlow * Flow.mt1 - whigh * Fhigh.mt1);
          acc.mt2 += dt * (wlow * Flow.mt2 - whigh * Fhigh.mt2);
          if (EQ != EQ_EULER) {
            acc.bbn += dt * ((wlow * Flow.bbn - whigh * Fhigh.bbn) + gs_bbn);
            acc.bbt1 += dt * (wlow * Flow.bbt1 - whigh * Fhigh.bbt1);
            acc.bbt2 += dt * (wlow * Flow.bbt2 - whigh * Fhigh.bbt2);
          }
          if (EQ == EQ_GLM) acc.psi += dt * (wlow * Flow.psi - whigh * Fhigh.psi);
        }
        // tracers: upwind on the sign of the mass flux (solver_eqn_base.cpp:281-342)
#pragma unroll
        for (int q = 0; q < PION_MAXTR; q++) {
          if (q < a.ntr) {
            const double* T = a.S + (long)(NB + q) * vs;
            double tm1 = __ldg(T + c - st), tc = __ldg(T + c), tp1 = __ldg(T + c + st);
            double lL = tm1, lR = tc, hL = tc, hR = tp1;
            if (a.order == 2) {
              double tm2 = __ldg(T + c - 2 * st), tp2 = __ldg(T + c + 2 * st);
              if (!radial) {
                double sm = minmod(tm1 - tm2, tc - tm1, a.tiny2), sc = minmod(tc - tm1, tp1 - tc, a.tiny2),
                       sp = minmod(tp1 - tc, tp2 - tp1, a.tiny2);
                lL = tm1 + sm * 0.5; lR = tc - sc * 0.5; hL = tc + sc * 0.5; hR = tp1 - sp * 0.5;
              } else {
                double s0 = (tm1 - tm2) * ri[0], s1 = (tc - tm1) * ri[1], s2 = (tp1 - tc) * ri[2], s3 = (tp2 - tp1) * ri[3];
                double sm = minmod(s0, s1, PION_VERY_TINY_VALUE), sc = minmod(s1, s2, PION_VERY_TINY_VALUE),
                       sp = minmod(s2, s3, PION_VERY_TINY_VALUE);
                lL = tm1 + sm * rdel[0]; lR = tc + sc * rdel[1]; hL = tc + sc * rdel[2]; hR = tp1 + sp * rdel[3];
              }
            }
            double fl = 0.0, fh = 0.0;
            if (Flow.rho > 0.0) fl = lL * Flow.rho * (a.pp.have_mp ? scma_corr(lL) : 1.0);
            else if (Flow.rho < 0.0) fl = lR * Flow.rho * (a.pp.have_mp ? scma_corr(lR) : 1.0);
            if (Fhigh.rho > 0.0) fh = hL * Fhigh.rho * (a.pp.have_mp ? scma_corr(hL) : 1.0);
            else if (Fhigh.rho < 0.0) fh = hR * Fhigh.rho * (a.pp.have_mp ? scma_corr(hR) : 1.0);
            acctr[q] += radial ? dt * (wlow * fl - whigh * fh) : dt * ((fl - fh) * idx);
          }
        }
        // rotate the centre state and the accumulators into the next axis' frame
        rot3(C.vn, C.vt1, C.vt2);
        rot3(acc.mn, acc.mt1, acc.mt2);
        if (EQ != EQ_EULER) {
          rot3(C.bn, C.bt1, C.bt2);
          rot3(acc.bbn, acc.bbt1, acc.bbt2);
        }
      }
      // back to the x frame after ndim rotations
      if (g.ndim == 2) {  // frame is (z,x,y)
        rot3(acc.mn, acc.mt1, acc.mt2);
        if (EQ != EQ_EULER) rot3(acc.bbn, acc.bbt1, acc.bbt2);
      } else if (g.ndim == 1) {  // frame is (y,z,x)
        rot3(acc.mn, acc.mt1, acc.mt2); rot3(acc.mn, acc.mt1, acc.mt2);
        if (EQ != EQ_EULER) { rot3(acc.bbn, acc.bbt1, acc.bbt2); rot3(acc.bbn, acc.bbt1, acc.bbt2); }
      }
    }

    if (!a.fused) {
      double* d = a.dU;
      d[c] += acc.rho; d[vs + c] += acc.erg; d[2 * vs + c] += acc.mn; d[3 * vs + c] += acc.mt1; d[4 * vs + c] += acc.mt2;
      if (EQ != EQ_EULER) { d[5 * vs + c] += acc.bbn; d[6 * vs + c] += acc.bbt1; d[7 * vs + c] += acc.bbt2; }
      if (EQ == EQ_GLM) d[8 * vs + c] += acc.psi;
#pragma unroll
      for (int q = 0; q < PION_MAXTR; q++)
        if (q < a.ntr) d[(NB + q) * vs + c] += acctr[q];
    } else if (domain) {
      status = cell_advance_time<EQ>(a, c, acc, acctr, a.ntr, my_dt);
    } else {
      // cell cut out of the domain (time_integrator.cpp:905-908): state untouched
      if (a.out != a.S) {
        for (int v = 0; v < NB + a.ntr; v++) a.out[(long)v * vs + c] = a.S[(long)v * vs + c];
      }
    }
  }

  stage_block_epilogue(a, my_dt, status | rs_status);
}

// The box of tiles / planes a block of the sweep kernels works on, and its tile coordinates inside it
struct BlockBox {
  int tx, ty, tz;      // tile coordinates (grid-absolute in x / y; chunk index in z)
  int k_lo, k_hi;      // plane range of the box
};
__device__ __forceinline__ BlockBox sweep_block_box(const StageArgs& a) {
  BlockBox b;
  if (a.nbox == 0) {
    b.tx = blockIdx.x + a.tx0; b.ty = blockIdx.y + a.ty0; b.tz = blockIdx.z; b.k_lo = a.k_lo; b.k_hi = a.k_hi;
    return b;
  }
  int q = 0;
  while (q + 1 < a.nbox && (int)blockIdx.x >= a.box_blk0[q + 1]) q++;
  const int r = blockIdx.x - a.box_blk0[q];
  const int nx = a.box[q][1] - a.box[q][0], ny = a.box[q][3] - a.box[q][2];
  b.tx = a.box[q][0] + r % nx;
  b.ty = a.box[q][2] + (r / nx) % ny;
  b.tz = r / (nx * ny);
  b.k_lo = a.box[q][4];
  b.k_hi = a.box[q][5];
  return b;
}
// host side: blocks of one box / fill box_blk0 and return the 1-D grid size
inline int sweep_box_blocks(const int* bx, int kchunk) {
  const int nx = bx[1] - bx[0], ny = bx[3] - bx[2], nz = bx[5] - bx[4];
  if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
  return nx * ny * ((nz + kchunk - 1) / kchunk);
}
inline int sweep_fill_box_table(StageArgs& a, int kchunk) {
  int tot = 0;
  for (int q = 0; q < a.nbox; q++) {
    a.box_blk0[q] = tot;
    tot += sweep_box_blocks(a.box[q], kchunk);
  }
  for (int q = a.nbox; q < 7; q++) a.box_blk0[q] = tot;
  return tot;
}

// host-side launcher implemented per equation set in stage_{euler,mhd,glm}.cu
// (every launcher returns the name of the kernel variant it launched: pion_gpu_describe reports it)
const char* launch_stage_euler(int solver, int fkj, const StageArgs& a, cudaStream_t s);
const char* launch_stage_mhd(int solver, int fkj, const StageArgs& a, cudaStream_t s);
const char* launch_stage_glm(int solver, int fkj, const StageArgs& a, cudaStream_t s);
// cells per sweep tile along x / y (stage_sweep.cuh: 32 lanes, TY rows, one of each only produces fluxes)
void sweep_tile_cells(int eq, int* cx, int* cy);  // eq: EQ_EULER / EQ_MHD / EQ_GLM
bool sweep_tma_fits(int eq, int ntr);  // the TMA sweep kernel exists for this many tracers and its tile fits shared memory
void sweep_tma_box(int eq, int order, int ntr, int* cw, int* rh, int* nb, int* tx, int* ty);  // box of one TMA plane load of a stage of that order with ntr tracers, cells per tile in x and y (stage_sweep_tma.cuh)
// flux-once sweep kernel (stage_sweep.cuh), instantiated in sweep_{euler,mhd,glm}.cu
const char* launch_sweep_euler(int solver, int fkj, const StageArgs& a, cudaStream_t s);
const char* launch_sweep_mhd(int solver, int fkj, const StageArgs& a, cudaStream_t s);
const char* launch_sweep_glm(int solver, int fkj, const StageArgs& a, cudaStream_t s);

// "k_xxx<EQ=3,SOLVER=7,FKJ=1,...>" built once per instantiation
inline const char* kernel_variant_name(char* buf, size_t n, const char* kernel, int eq, int solver, bool fkj, const char* extra) {
  snprintf(buf, n, "%s<EQ=%d,SOLVER=%d,FKJ=%d%s>", kernel, eq, solver, fkj ? 1 : 0, extra);
  return buf;
}

template <int EQ, int SOLVER, bool FKJ>
inline const char* launch_stage_t(const StageArgs& a, cudaStream_t s) {
  const long ncell = (long)a.g.NG[0] * a.g.NG[1] * a.g.NG[2];
  const int block = 128;
  const long grid = (ncell + block - 1) / block;
  k_stage<EQ, SOLVER, FKJ><<<(unsigned)grid, block, 0, s>>>(a);
  static char name[96];
  static const char* nm = kernel_variant_name(name, sizeof name, "k_stage", EQ, SOLVER, FKJ, " (gather form)");
  return nm;
}

}  // namespace pion
