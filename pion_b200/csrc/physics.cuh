// pion_b200/csrc/physics.cuh -- per-interface / per-cell device math of the
// finite-volume dynamics update, FP64, written in the *solver frame*:
// (n, t1, t2) = (sweep axis, next axis, next-next axis) in cyclic order, which
// is exactly the permutation eqns_base::SetDirection establishes in the
// reference (source/equations/eqns_base.cpp:94-131).  The kernels permute at
// load / accumulate time, so nothing in here indexes a state vector
// dynamically (everything stays in registers).
//
// Reference semantics restated here (paths relative to /root/reference/source):
//   equations/eqns_hydro_adiabatic.cpp:89-350   PtoU, UtoP(+floors), PUtoFlux, UtoFlux, chydro
//   equations/eqns_mhd_adiabatic.cpp:79-337,581-660  PtoU, UtoP, check_pressure, cfast, PUtoFlux, GLM
//   Riemann_solvers/HLL_hydro.cpp:92-170         Euler HLL
//   Riemann_solvers/HLLD_MHD.cpp:124-417         MHD HLLD / HLL / signal speeds
//   Riemann_solvers/Roe_Hydro_ConservedVar_solver.cpp:129-436   Euler Roe (conserved variables)
//   Riemann_solvers/Roe_MHD_ConservedVar_solver.cpp:218-810,1074-1131  MHD Roe (Cargo & Gallice)
//   spatial_solvers/solver_eqn_hydro_adi.cpp:94-205,283-333      Euler inviscid_flux, AVFalle
//   spatial_solvers/solver_eqn_mhd_adi.cpp:102-288,662-772       MHD / GLM inviscid_flux, AVFalle
//   coord_sys/VectorOps.cpp:40-59                minmod ("AvgFalle")
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "fastmath.cuh"

namespace pion {

enum : int { EQ_EULER = 1, EQ_MHD = 2, EQ_GLM = 3 };                 // constants.h:166-172
enum : int { SOLVE_LF = 0, SOLVE_RSLINEAR = 1, SOLVE_RSEXACT = 2, SOLVE_RSHYBRID = 3, SOLVE_ROE = 4, SOLVE_ROE_PV = 5, SOLVE_FVS = 6, SOLVE_HLLD = 7, SOLVE_HLL = 8 };  // constants.h:238-246 (5, 6: Euler only)
enum : int { AV_NONE = 0, AV_FKJ98 = 1, AV_HCORR = 3, AV_HCORR_FKJ98 = 4 };

#define PION_MACHINEACCURACY 5.e-16    // constants.h:151
#define PION_TINYVALUE 1.0e-100        // constants.h:152
#define PION_SMALLVALUE 1.0e-12        // constants.h:150
#define PION_VERY_TINY_VALUE 1.0e-200  // constants.h:153
#define PION_BASE_RHO 1.0e-5           // constants.h:339

// Primitive state in the solver frame.
struct Prim {
  double ro, pg, vn, vt1, vt2, bn, bt1, bt2, psi;
};
// Conserved state / flux in the solver frame.
struct Cons {
  double rho, erg, mn, mt1, mt2, bbn, bbt1, bbt2, psi;
};

// Per-launch physics constants (kernel argument, lives in constant bank).
struct PhysParams {
  double gamma;
  double etav;         // FKJ98 viscosity coefficient (FV_etav == FV_etaB)
  double chyp;         // GLM hyperbolic speed c_h
  double refvec_ro;    // RefVec[RO] for the (fatal) negative-density reset
  double min_temp;     // EP.MinTemperature
  double max_temp;     // EP.MaxTemperature
  double mu_tot_over_kB;  // mp_only_cooling::Mu_tot_over_kB (0 if no microphysics)
  int have_mp;
  double lf_c;         // Lax-Friedrichs only: dx / FV_dt (set per stage launch)
  double lf_ndim;      // ... and FV_gndim
  double rs_refvec[5]; // linear / exact / hybrid Riemann solvers: Euler riemann_Euler::eq_refvec = RefVec[RO, PG] and 0.1 c(RefVec) three times; MHD {RefVec[RO], RefVec[PG], 0.1 cfast(RefVec), |B(RefVec)|} (SetAvgState)
};

__device__ __forceinline__ double sq(double x) { return x * x; }

// p/(gamma-1): gamma is launch-uniform, so the reciprocal is loop-invariant and hoisted
#ifdef PION_STRICT
#define PION_OVER_GM1(x, gm1) ((x) / (gm1))
#else
#define PION_OVER_GM1(x, gm1) ((x) * fast_rcp(gm1))
#endif

// ---------------------------------------------------------------------------
// minmod: BaseVectorOps::AvgFalle, AVG_MINMOD variant (VectorOps.cpp:40-59).
// The reference computes r=a/b; min(r,1)*b.  For 0<r<1 that is (a/b)*b which
// equals `a` to within one rounding; we return `a` directly and save the FP64
// division (documented deviation, <= 1 ulp of the slope).  PION_STRICT keeps
// the reference's exact expression.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double minmod(double a, double b, double tiny) {
#ifdef PION_STRICT
  const double ab = a * b;
  if (ab <= tiny) return 0.0;
  double r = a / b;
  return (r > 0.0) ? fmin(r, 1.0) * b : 0.0;
#else
  // Opposite signs are tested on the sign bits (integer pipe) instead of a DMUL + DSETP on the FP64 pipe; a
  // zero operand still returns zero (it is the smaller magnitude).  Products below `tiny` = 1e-200 dx^2 are
  // not flushed to zero as in the reference: an absolute difference far below any variable's rounding.
  (void)tiny;
  const double m = (fabs(a) < fabs(b)) ? a : b;
  return ((__double2hiint(a) ^ __double2hiint(b)) < 0) ? 0.0 : m;
#endif
}
// e += minmod(a, b) * h for callers that only ADD the limited slope to something (h = +-0.5: the product is exact,
// so the fused form is bit-identical): the smaller-magnitude operand goes into ONE PREDICATED DFMA (predicate: the
// sign bits agree) instead of DMUL + two FSEL (zeroing) + DADD.  Inline PTX because the compiler if-converts the C
// form into an unconditional DFMA followed by two FSEL.
__device__ __forceinline__ void add_limited(double& e, double a, double b, double h) {
  const double m = (fabs(a) < fabs(b)) ? a : b;
  asm("{\n\t.reg .pred p;\n\t.reg .b32 x;\n\txor.b32 x, %3, %4;\n\tsetp.ge.s32 p, x, 0;\n\t@p fma.rn.f64 %0, %1, %2, %0;\n\t}"
      : "+d"(e)
      : "d"(m), "d"(h), "r"(__double2hiint(a)), "r"(__double2hiint(b)));
}

// ---------------------------------------------------------------------------
// Equations of state / conversions
// ---------------------------------------------------------------------------
template <int EQ>
__device__ __forceinline__ void PtoU(const Prim& p, Cons& u, double gm1) {
  u.rho = p.ro;
  u.mn = p.ro * p.vn;
  u.mt1 = p.ro * p.vt1;
  u.mt2 = p.ro * p.vt2;
  if (EQ == EQ_EULER) {
    u.erg = p.ro * (p.vn * p.vn + p.vt1 * p.vt1 + p.vt2 * p.vt2) * 0.5 + PION_OVER_GM1(p.pg, gm1);
    u.bbn = u.bbt1 = u.bbt2 = u.psi = 0.0;
  } else {
    u.bbn = p.bn;
    u.bbt1 = p.bt1;
    u.bbt2 = p.bt2;
    u.erg = (p.ro * (p.vn * p.vn + p.vt1 * p.vt1 + p.vt2 * p.vt2) * 0.5) + PION_OVER_GM1(p.pg, gm1) +
            ((u.bbn * u.bbn + u.bbt1 * u.bbt1 + u.bbt2 * u.bbt2) * 0.5);
    if (EQ == EQ_GLM) {
      u.psi = p.psi;
      u.erg += 0.5 * u.psi * u.psi;
    } else {
      u.psi = 0.0;
    }
  }
}
// ideal-MHD PtoU without the psi energy (what the Riemann solvers call:
// eqns_mhd_ideal::PtoU, HLLD_MHD.cpp:139-140)
__device__ __forceinline__ void PtoU_mhd_ideal(const Prim& p, Cons& u, double gm1) {
  PtoU<EQ_MHD>(p, u, gm1);
}

// status bits returned by UtoP
enum : int { ST_NEG_RHO = 1, ST_NEG_PG = 2, ST_RS_FAIL = 4 };  // ST_RS_FAIL: JMs_riemann_solve returned an error

// UtoP with the reference's floors (SET_NEGATIVE_PRESSURE_TO_FIXED_TEMPERATURE):
// Euler eqns_hydro_adiabatic.cpp:117-205, MHD eqns_mhd_adiabatic.cpp:110-224,
// GLM :618-641.
template <int EQ>
__device__ __forceinline__ int UtoP(const Cons& u, Prim& p, const PhysParams& pp) {
  int st = 0;
  const double gm1 = pp.gamma - 1.0;
  p.ro = u.rho;
#ifdef PION_STRICT
  p.vn = u.mn / u.rho;
  p.vt1 = u.mt1 / u.rho;
  p.vt2 = u.mt2 / u.rho;
#else
  const double ir = fast_rcp(u.rho);
  p.vn = u.mn * ir;
  p.vt1 = u.mt1 * ir;
  p.vt2 = u.mt2 * ir;
#endif
  double ke = p.ro * (p.vn * p.vn + p.vt1 * p.vt1 + p.vt2 * p.vt2);
  if (EQ == EQ_EULER) {
    p.pg = gm1 * (u.erg - ke / 2.0);
    p.bn = p.bt1 = p.bt2 = p.psi = 0.0;
  } else {
    double b2 = (u.bbn * u.bbn + u.bbt1 * u.bbt1 + u.bbt2 * u.bbt2);
    if (EQ == EQ_GLM) {
      p.psi = u.psi;
      p.pg = gm1 * (u.erg - ke * 0.5 - 0.5 * u.psi * u.psi - b2 * 0.5);
    } else {
      p.psi = 0.0;
      p.pg = gm1 * (u.erg - ke / 2. - b2 / 2.);
    }
    p.bn = u.bbn;
    p.bt1 = u.bbt1;
    p.bt2 = u.bbt2;
  }
  if (p.ro <= 0.0) {
    // fatal in the reference (rep.error); we flag it and apply the code that
    // follows the rep.error call so that the kernel stays finite.
    st |= ST_NEG_RHO;
    if (EQ == EQ_EULER) {
      p.ro = PION_BASE_RHO;
      p.vn = u.mn / p.ro;
      p.vt1 = u.mt1 / p.ro;
      p.vt2 = u.mt2 / p.ro;
      p.pg = gm1 * (u.erg - p.ro * (p.vn * p.vn + p.vt1 * p.vt1 + p.vt2 * p.vt2) / 2.0);
    } else {
      p.ro = PION_BASE_RHO * pp.refvec_ro;
      double f = u.rho / p.ro;
      p.vn *= f;
      p.vt1 *= f;
      p.vt2 *= f;
      p.pg = gm1 * (u.erg - p.ro * (p.vn * p.vn + p.vt1 * p.vt1 + p.vt2 * p.vt2) / 2. -
                    (u.bbn * u.bbn + u.bbt1 * u.bbt1 + u.bbt2 * u.bbt2) / 2.);
    }
  }
  if (p.pg <= 0.0) {
    st |= ST_NEG_PG;
    if (pp.have_mp) p.pg = p.ro * pp.min_temp / pp.mu_tot_over_kB;  // MP->Set_Temp(p,MinTemp)
    else p.pg = 0.01 * p.ro;
  } else if (pp.have_mp && (p.pg * pp.mu_tot_over_kB < pp.min_temp * p.ro)) {  // T < Tmin, rho > 0 here
    p.pg = p.ro * pp.min_temp / pp.mu_tot_over_kB;
  }
  return st;
}

// eqns_Euler::PUtoFlux (eqns_hydro_adiabatic.cpp:296-308) /
// eqns_mhd_ideal::PUtoFlux (eqns_mhd_adiabatic.cpp:307-328)
template <int EQ>
__device__ __forceinline__ void PUtoFlux(const Prim& p, const Cons& u, Cons& f) {
  f.rho = u.mn;
  if (EQ == EQ_EULER) {
    f.mn = u.mn * p.vn + p.pg;
    f.mt1 = u.mn * p.vt1;
    f.mt2 = u.mn * p.vt2;
    f.erg = p.vn * (u.erg + p.pg);
    f.bbn = f.bbt1 = f.bbt2 = f.psi = 0.0;
  } else {
    double pm = (u.bbn * u.bbn + u.bbt1 * u.bbt1 + u.bbt2 * u.bbt2) / 2.;
    f.mn = u.mn * p.vn + p.pg + pm - u.bbn * u.bbn;
    f.mt1 = u.mn * p.vt1 - u.bbn * u.bbt1;
    f.mt2 = u.mn * p.vt2 - u.bbn * u.bbt2;
    f.erg = p.vn * (u.erg + p.pg + pm) - u.bbn * (p.vn * u.bbn + p.vt1 * u.bbt1 + p.vt2 * u.bbt2);
    f.bbn = 0.;
    f.bbt1 = p.vn * p.bt1 - p.vt1 * p.bn;
    f.bbt2 = p.vn * p.bt2 - p.vt2 * p.bn;
    f.psi = 0.0;
  }
}

__device__ __forceinline__ double chydro(double ro, double pg, double g) {
#ifdef PION_STRICT
  return sqrt(g * pg / ro);
#else
  return fast_sqrt(g * pg * fast_rcp(ro));
#endif
}

// eqns_mhd_ideal::cfast_components (eqns_mhd_adiabatic.cpp:263-276).
// cfast2_ir returns the SQUARE of the fast speed given 1/ro, so that callers needing
// max(cf_l, cf_r) take one square root of the larger square (sqrt is monotonic: same value).
// irb: the reciprocal the FIELD terms are divided by (== ir unless the caller passes sums of two states, see
// intercell_flux: pg and B are then twice the mean state's, ir = 1/(2 rho), irb = ir/2)
__device__ __forceinline__ double cfast2_ir(double ir, double pg, double bx, double by, double bz, double g, double irb) {
  // one reciprocal instead of a sqrt + three divisions; ch*ch == g*pg/ro to 1 ulp
  double ch2 = g * pg * ir;
  // b_x^2/rho formed once and shared by both terms (two multiplies fewer than the literal form; 1 ulp)
  const double bxi = (bx * bx) * irb;
  double temp1 = (ch2 + bxi) + (by * by + bz * bz) * irb;
  double temp2 = (4. * ch2) * bxi;
  temp2 = pmax(temp1 * temp1 - temp2, PION_MACHINEACCURACY);
  return (temp1 + fast_sqrt_pos(temp2)) / 2.;
}
__device__ __forceinline__ double cfast2_ir(double ir, double pg, double bx, double by, double bz, double g) {
  return cfast2_ir(ir, pg, bx, by, bz, g, ir);
}
__device__ __forceinline__ double cfast_components(double ro, double pg, double bx, double by, double bz, double g) {
#ifdef PION_STRICT
  double ch = sqrt(g * pg / ro);
  double temp1 = ch * ch + (bx * bx + by * by + bz * bz) / ro;
  double temp2 = 4. * ch * ch * bx * bx / ro;
  temp2 = pmax(temp1 * temp1 - temp2, PION_MACHINEACCURACY);
  return sqrt((temp1 + sqrt(temp2)) / 2.);
#else
  return fast_sqrt_pos(cfast2_ir(fast_rcp(ro), pg, bx, by, bz, g));
#endif
}

// ---------------------------------------------------------------------------
// Riemann solvers.  All take edge states in the solver frame and return the
// flux; `ustar`/`pstar` is only produced when the FKJ98 viscosity needs it.
// ---------------------------------------------------------------------------

// HLL_hydro::hydro_HLL_flux_solver (HLL_hydro.cpp:118-170)
__device__ __forceinline__ void hydro_HLL(const Prim& L, const Prim& R, const PhysParams& pp, Cons& flux, Cons& ustar) {
  const double gm1 = pp.gamma - 1.0;
  Cons UL, UR, FL, FR;
  PtoU<EQ_EULER>(L, UL, gm1);
  PtoU<EQ_EULER>(R, UR, gm1);
  PUtoFlux<EQ_EULER>(L, UL, FL);
  PUtoFlux<EQ_EULER>(R, UR, FR);
  double cf_max = pmax(chydro(L.ro, L.pg, pp.gamma), chydro(R.ro, R.pg, pp.gamma));
  double Sl = pmin(L.vn, R.vn) - cf_max;
  double Sr = pmax(L.vn, R.vn) + cf_max;
  double idS = fast_rcp(Sr - Sl);
#define PION_HLL_COMP(c)                                                             \
  flux.c = (Sl > 0) ? FL.c : (Sr < 0) ? FR.c : (Sr * FL.c - Sl * FR.c + Sr * Sl * (UR.c - UL.c)) * idS; \
  ustar.c = (Sr * UR.c - Sl * UL.c + FL.c - FR.c) * idS;
  PION_HLL_COMP(rho) PION_HLL_COMP(erg) PION_HLL_COMP(mn) PION_HLL_COMP(mt1) PION_HLL_COMP(mt2)
#undef PION_HLL_COMP
  flux.bbn = flux.bbt1 = flux.bbt2 = flux.psi = 0.0;
  ustar.bbn = ustar.bbt1 = ustar.bbt2 = ustar.psi = 0.0;
}

// constants::equalD (constants.cpp:48-69)
__device__ __forceinline__ bool equalD(double a, double b) {
  if (a == b) return true;
  if (fabs(a) + fabs(b) < PION_TINYVALUE) return true;
#ifdef PION_STRICT
  return (fabs(a - b) / (fabs(a) + fabs(b) + PION_TINYVALUE)) < PION_SMALLVALUE;
#else
  return fabs(a - b) < PION_SMALLVALUE * (fabs(a) + fabs(b) + PION_TINYVALUE);  // same test, no division
#endif
}

// eqns_Euler::UtoFlux (eqns_hydro_adiabatic.cpp:317-333)
__device__ __forceinline__ void euler_UtoFlux(const Cons& u, Cons& f, double gm1) {
  double ir = fast_rcp(u.rho);
  double pg = gm1 * (u.erg - (u.mn * u.mn + u.mt1 * u.mt1 + u.mt2 * u.mt2) * 0.5 * ir);
  f.rho = u.mn;
  f.mn = u.mn * u.mn * ir + pg;
  f.mt1 = u.mn * u.mt1 * ir;
  f.mt2 = u.mn * u.mt2 * ir;
  f.erg = u.mn * (u.erg + pg) * ir;
}

// eqns_mhd_ideal::UtoFlux (eqns_mhd_adiabatic.cpp:337-355)
__device__ __forceinline__ void mhd_UtoFlux(const Cons& u, Cons& f, double gm1) {
  const double ir = fast_rcp(u.rho);
  const double pm = (u.bbn * u.bbn + u.bbt1 * u.bbt1 + u.bbt2 * u.bbt2) / 2.;
  const double pg = gm1 * (u.erg - (u.mn * u.mn + u.mt1 * u.mt1 + u.mt2 * u.mt2) * 0.5 * ir - pm);
  f.rho = u.mn;
  f.mn = u.mn * u.mn * ir + pg + pm - u.bbn * u.bbn;
  f.mt1 = u.mn * u.mt1 * ir - u.bbn * u.bbt1;
  f.mt2 = u.mn * u.mt2 * ir - u.bbn * u.bbt2;
  f.erg = u.mn * (u.erg + pg + pm) * ir - u.bbn * (u.mn * u.bbn + u.mt1 * u.bbt1 + u.mt2 * u.bbt2) * ir;
  f.bbn = 0.;
  f.bbt1 = (u.mn * u.bbt1 - u.mt1 * u.bbn) * ir;
  f.bbt2 = (u.mn * u.bbt2 - u.mt2 * u.bbn) * ir;
  f.psi = 0.0;
}

// FV_solver_base::get_LaxFriedrichs_flux (solver_eqn_base.cpp:109-141) + pstar = mean of the edge states
// (solver_eqn_hydro_adi.cpp:142-148, solver_eqn_mhd_adi.cpp:132-136).  The reference forces first order in space
// and time with this flux (setup_fixed_grid.cpp:188-190); pion_gpu_create does the same.
template <int EQ>
__device__ __forceinline__ void lax_friedrichs(const Prim& L, const Prim& R, const PhysParams& pp, Cons& flux, Prim& pstar) {
  const double gm1 = pp.gamma - 1.0;
  Cons u1, u2, f1, f2;
  if (EQ == EQ_EULER) {
    PtoU<EQ_EULER>(L, u1, gm1); PtoU<EQ_EULER>(R, u2, gm1);
    euler_UtoFlux(u1, f1, gm1); euler_UtoFlux(u2, f2, gm1);
  } else {
    PtoU_mhd_ideal(L, u1, gm1); PtoU_mhd_ideal(R, u2, gm1);
    mhd_UtoFlux(u1, f1, gm1); mhd_UtoFlux(u2, f2, gm1);
  }
#define PION_LF_COMP(c) flux.c = 0.5 * (f1.c + f2.c + pp.lf_c * (u1.c - u2.c) / pp.lf_ndim);
  PION_LF_COMP(rho) PION_LF_COMP(erg) PION_LF_COMP(mn) PION_LF_COMP(mt1) PION_LF_COMP(mt2)
  if (EQ != EQ_EULER) { PION_LF_COMP(bbn) PION_LF_COMP(bbt1) PION_LF_COMP(bbt2) } else { flux.bbn = flux.bbt1 = flux.bbt2 = 0.0; }
#undef PION_LF_COMP
  flux.psi = 0.0;
  pstar.ro = 0.5 * (L.ro + R.ro); pstar.pg = 0.5 * (L.pg + R.pg);
  pstar.vn = 0.5 * (L.vn + R.vn); pstar.vt1 = 0.5 * (L.vt1 + R.vt1); pstar.vt2 = 0.5 * (L.vt2 + R.vt2);
  pstar.bn = 0.5 * (L.bn + R.bn); pstar.bt1 = 0.5 * (L.bt1 + R.bt1); pstar.bt2 = 0.5 * (L.bt2 + R.bt2);
  pstar.psi = 0.5 * (L.psi + R.psi);
}

// Roe-average primitive state of two Euler states (Toro eq. 11.60): Riemann_FVS_Euler::Roe_average_state
// (Riemann_FVS_hydro.cpp:205-248) and the first part of Roe_prim_var_solver.  Returns the mean sound speed^2.
__device__ __forceinline__ double hydro_roe_average(const Prim& L, const Prim& R, double g, Prim& m) {
  const double gm1 = g - 1.0;
  const double rl = psqrt(L.ro), rr = psqrt(R.ro);
  const double lH = 0.5 * (L.vn * L.vn + L.vt1 * L.vt1 + L.vt2 * L.vt2) + pdiv(g * L.pg, gm1 * L.ro);
  const double rH = 0.5 * (R.vn * R.vn + R.vt1 * R.vt1 + R.vt2 * R.vt2) + pdiv(g * R.pg, gm1 * R.ro);
  const double denom = fast_rcp(rl + rr);
  m.ro = rl * rr;
  m.vn = (rl * L.vn + rr * R.vn) * denom;
  m.vt1 = (rl * L.vt1 + rr * R.vt1) * denom;
  m.vt2 = (rl * L.vt2 + rr * R.vt2) * denom;
  const double H = (rl * lH + rr * rH) * denom;
  const double a2 = gm1 * (H - 0.5 * (m.vn * m.vn + m.vt1 * m.vt1 + m.vt2 * m.vt2));
  m.pg = pdiv(m.ro * a2, g);
  m.bn = m.bt1 = m.bt2 = m.psi = 0.0;
  return a2;
}

// Riemann_FVS_Euler::FVS_flux (Riemann_FVS_hydro.cpp:84-198): van Leer (1982) flux-vector splitting;
// pstar = the Roe-average state (only the viscosity reads it)
template <bool NEED_PSTAR>
__device__ __forceinline__ void hydro_FVS(const Prim& L, const Prim& R, const PhysParams& pp, Cons& flux, Prim& pstar) {
  const double g = pp.gamma, gm1 = g - 1.0;
  const double ig = fast_rcp(g), ig21 = fast_rcp(g * g - 1.0);
  Cons fp, fn;
  fp.rho = fp.erg = fp.mn = fp.mt1 = fp.mt2 = 0.0;
  fn.rho = fn.erg = fn.mn = fn.mt1 = fn.mt2 = 0.0;
  const double cl = chydro(L.ro, L.pg, g), cr = chydro(R.ro, R.pg, g);
  const double Ml = pdiv(L.vn, cl), Mr = pdiv(R.vn, cr);
  if (Ml < -1.0) {
  } else if (Ml > 1.0) {
    Cons u;
    PtoU<EQ_EULER>(L, u, gm1);
    PUtoFlux<EQ_EULER>(L, u, fp);
  } else {
    const double f1 = 0.25 * L.ro * cl * (1.0 + Ml) * (1.0 + Ml);
    const double f2 = cl * (gm1 * Ml + 2);
    fp.rho = f1;
    fp.mn = f1 * f2 * ig;
    fp.mt1 = f1 * L.vt1;
    fp.mt2 = f1 * L.vt2;
    fp.erg = f1 * (f2 * f2 * 0.5 * ig21 + 0.5 * (L.vt1 * L.vt1 + L.vt2 * L.vt2));
  }
  if (Mr > 1.0) {
  } else if (Mr < -1.0) {
    Cons u;
    PtoU<EQ_EULER>(R, u, gm1);
    PUtoFlux<EQ_EULER>(R, u, fn);
  } else {
    const double f1 = -0.25 * R.ro * cr * (1.0 - Mr) * (1.0 - Mr);
    const double f2 = cr * (gm1 * Mr - 2);
    fn.rho = f1;
    fn.mn = f1 * f2 * ig;
    fn.mt1 = f1 * R.vt1;
    fn.mt2 = f1 * R.vt2;
    fn.erg = f1 * (f2 * f2 * 0.5 * ig21 + 0.5 * (R.vt1 * R.vt1 + R.vt2 * R.vt2));
  }
  flux.rho = fp.rho + fn.rho; flux.erg = fp.erg + fn.erg;
  flux.mn = fp.mn + fn.mn; flux.mt1 = fp.mt1 + fn.mt1; flux.mt2 = fp.mt2 + fn.mt2;
  flux.bbn = flux.bbt1 = flux.bbt2 = flux.psi = 0.0;
  if (NEED_PSTAR) hydro_roe_average(L, R, g, pstar);
}

// Riemann_Roe_Hydro_PV::Roe_prim_var_solver (Roe_Hydro_PrimitiveVar_solver.cpp:62-209): interface state of
// the solver linearised about the Roe average; the flux is PtoFlux(pstar) (solver_eqn_hydro_adi.cpp:178-187)
__device__ __forceinline__ void hydro_RoePV(const Prim& L, const Prim& R, const PhysParams& pp, Cons& flux, Prim& pstar) {
  const double g = pp.gamma;
  Prim m;
  const double a2 = hydro_roe_average(L, R, g, m);
  const double a_mean = psqrt(a2);
  if (m.vn - a_mean >= 0.) {
    pstar = L;
  } else if (m.vn + a_mean <= 0.) {
    pstar = R;
  } else {
    const double ia = fast_rcp(a_mean);
    pstar.pg = 0.5 * (L.pg + R.pg - m.ro * a_mean * (R.vn - L.vn));
    pstar.vn = 0.5 * (L.vn + R.vn - (R.pg - L.pg) * fast_rcp(m.ro) * ia);
    if (pstar.vn > 0.0) {
      pstar.ro = L.ro + m.ro * (L.vn - pstar.vn) * ia;
      pstar.vt1 = L.vt1;
      pstar.vt2 = L.vt2;
    } else {
      pstar.ro = R.ro + m.ro * (pstar.vn - R.vn) * ia;
      pstar.vt1 = R.vt1;
      pstar.vt2 = R.vt2;
    }
  }
  pstar.bn = pstar.bt1 = pstar.bt2 = pstar.psi = 0.0;
  Cons u;
  PtoU<EQ_EULER>(pstar, u, g - 1.0);
  PUtoFlux<EQ_EULER>(pstar, u, flux);
}

// Riemann_Roe_Hydro_CV::Roe_flux_solver_symmetric
// (Roe_Hydro_ConservedVar_solver.cpp:129-175 + helpers :215-436)
__device__ __forceinline__ void hydro_RoeCV(const Prim& L, const Prim& R, const PhysParams& pp, double hc_eta, Cons& flux,
                                            Prim& pstar) {
  const double g = pp.gamma, gm1 = g - 1.0;
  double rl = psqrt(L.ro), rr = psqrt(R.ro);
  double lH = 0.5 * (L.vn * L.vn + L.vt1 * L.vt1 + L.vt2 * L.vt2) + pdiv(g * L.pg, gm1 * L.ro);
  double rH = 0.5 * (R.vn * R.vn + R.vt1 * R.vt1 + R.vt2 * R.vt2) + pdiv(g * R.pg, gm1 * R.ro);
  double denom = fast_rcp(rl + rr);
  double m_ro = rl * rr;
  double m_vn = (rl * L.vn + rr * R.vn) * denom;
  double m_vt1 = (rl * L.vt1 + rr * R.vt1) * denom;
  double m_vt2 = (rl * L.vt2 + rr * R.vt2) * denom;
  double m_H = (rl * lH + rr * rH) * denom;
  double v2 = m_vn * m_vn + m_vt1 * m_vt1 + m_vt2 * m_vt2;
  double a = psqrt(gm1 * pmax(m_H - 0.5 * v2, 1.0e-12 * v2));
  const double ia = fast_rcp(a);
  double ev[5] = {m_vn - a, m_vn, m_vn, m_vn, m_vn + a};
#pragma unroll
  for (int v = 0; v < 5; v++) ev[v] = (ev[v] < 0.0) ? pmin(ev[v], -hc_eta) : pmax(ev[v], hc_eta);
  Cons ul, ur;
  PtoU<EQ_EULER>(L, ul, gm1);
  PtoU<EQ_EULER>(R, ur, gm1);
  double d_rho = equalD(ur.rho, ul.rho) ? 0.0 : ur.rho - ul.rho;
  double d_erg = equalD(ur.erg, ul.erg) ? 0.0 : ur.erg - ul.erg;
  double d_mn = equalD(ur.mn, ul.mn) ? 0.0 : ur.mn - ul.mn;
  double d_mt1 = equalD(ur.mt1, ul.mt1) ? 0.0 : ur.mt1 - ul.mt1;
  double d_mt2 = equalD(ur.mt2, ul.mt2) ? 0.0 : ur.mt2 - ul.mt2;
  double s2 = d_mt1 - m_vt1 * d_rho;
  double s3 = d_mt2 - m_vt2 * d_rho;
  double u5bar = d_erg - s2 * m_vt1 - s3 * m_vt2;
  double s1 = (d_rho * (m_H - m_vn * m_vn) + m_vn * d_mn - u5bar) * gm1 * ia * ia;
  double s0 = 0.5 * (d_rho * (m_vn + a) - d_mn - a * s1) * ia;
  double s4 = d_rho - s0 - s1;
  Cons fl, fr;
  euler_UtoFlux(ul, fl, gm1);
  euler_UtoFlux(ur, fr, gm1);
  flux.rho = fl.rho + fr.rho;
  flux.mn = fl.mn + fr.mn;
  flux.mt1 = fl.mt1 + fr.mt1;
  flux.mt2 = fl.mt2 + fr.mt2;
  flux.erg = fl.erg + fr.erg;
  // wave 0: (1, vn-a, vt1, vt2, H - vn a)
  double w = s0 * fabs(ev[0]);
  flux.rho -= w; flux.mn -= w * (m_vn - a); flux.mt1 -= w * m_vt1; flux.mt2 -= w * m_vt2; flux.erg -= w * (m_H - m_vn * a);
  // wave 1: (1, vn, vt1, vt2, v2/2)
  w = s1 * fabs(ev[1]);
  flux.rho -= w; flux.mn -= w * m_vn; flux.mt1 -= w * m_vt1; flux.mt2 -= w * m_vt2; flux.erg -= w * (0.5 * v2);
  // wave 2: (0,0,1,0,vt1)
  w = s2 * fabs(ev[2]);
  flux.mt1 -= w; flux.erg -= w * m_vt1;
  // wave 3: (0,0,0,1,vt2)
  w = s3 * fabs(ev[3]);
  flux.mt2 -= w; flux.erg -= w * m_vt2;
  // wave 4: (1, vn+a, vt1, vt2, H + vn a)
  w = s4 * fabs(ev[4]);
  flux.rho -= w; flux.mn -= w * (m_vn + a); flux.mt1 -= w * m_vt1; flux.mt2 -= w * m_vt2; flux.erg -= w * (m_H + m_vn * a);
  flux.rho *= 0.5; flux.mn *= 0.5; flux.mt1 *= 0.5; flux.mt2 *= 0.5; flux.erg *= 0.5;
  flux.bbn = flux.bbt1 = flux.bbt2 = flux.psi = 0.0;
  pstar.ro = m_ro; pstar.vn = m_vn; pstar.vt1 = m_vt1; pstar.vt2 = m_vt2;
  pstar.pg = pdiv(m_ro * a * a, g);
  pstar.bn = pstar.bt1 = pstar.bt2 = pstar.psi = 0.0;
}


// ---------------------------------------------------------------------------
// Euler linear / exact / hybrid Riemann solvers (solverType 1, 2, 3):
// riemann_Euler::JMs_riemann_solve (Riemann_solvers/riemann.cpp:245-463) with linear_solver (:674-747), linearOK
// (:592-600), exact_solver (:754-822: p* from findroot::solve_pos = bracket_root_pos + Brent's zbrent,
// findroot.cpp:158-183,270-310,359-452, on eqns_Euler::HydroWave, eqns_hydro_adiabatic.cpp:221-302),
// check_wave_locations (:471-585), solve_rarerare (:829-885), solve_cavitation (:892-960).  States are
// {ro, pg, vn} (+ vt1, vt2 copied from the upwind side).  Plain IEEE arithmetic (exp / log / pow / sqrt / division as
// written in the reference): these solvers run on the gather kernel only and are not performance paths.
// ---------------------------------------------------------------------------
struct RsEuler {
  double g;
  double L[3], R[3], ps[3], cl, cr;  // ro, pg, vn
};
__device__ inline double rs_hydro_wave(double gamma, int lr, double pp, const double* pre) {
  const double pratio = pp / pre[1];
  const double c0 = sqrt(gamma * pre[1] / pre[0]);
  double u;
  if (pratio < 1) {
    u = 2. * c0 / (gamma - 1.) * (1 - exp((gamma - 1.) / 2. / gamma * log(pratio)));
    u = (lr == 0) ? pre[2] + u : pre[2] - u;
  } else if (pratio > 1) {
    u = c0 * (pratio - 1.) / sqrt(gamma * (gamma - 1.) / 2. * (1. + pratio * (gamma + 1.) / (gamma - 1.)));
    u = (lr == 0) ? pre[2] - u : pre[2] + u;
  } else {
    u = pre[2];
  }
  return u;
}
__device__ inline void rs_hydro_wave_full(double gamma, int lr, double pp, const double* pre, double* u, double* rho) {
  const double pratio = pp / pre[1];
  *u = rs_hydro_wave(gamma, lr, pp, pre);
  if (pratio < 1) *rho = pre[0] * exp(log(pratio) / gamma);
  else if (pratio > 1) *rho = pre[0] * (1 + pratio * (gamma + 1) / (gamma - 1.)) / ((gamma + 1.) / (gamma - 1.) + pratio);
  else *rho = pre[0];
}
__device__ inline double rs_root_function(const RsEuler& r, double pp) {
  return rs_hydro_wave(r.g, 1, pp, r.R) - rs_hydro_wave(r.g, 0, pp, r.L);
}
__device__ inline int rs_bracket_root_pos(const RsEuler& r, double* x1, double* x2) {
  const float factor = 1.6f;  // a float in the reference
  if (*x1 == *x2) return 1;
  if (*x1 > *x2) { const double t = *x1; *x1 = *x2; *x2 = t; }
  double f1 = rs_root_function(r, *x1), f2 = rs_root_function(r, *x2);
  for (int j = 0; j < 50; j++) {
    if (f1 * f2 < 0) return 0;
    if (fabs(f1) < fabs(f2)) { *x1 *= 1. / factor; f1 = rs_root_function(r, *x1); }
    else { *x2 *= factor; f2 = rs_root_function(r, *x2); }
  }
  *x1 = 0.;
  f1 = rs_root_function(r, *x1);
  if (f1 * f2 < 0) return 0;
  *x1 = *x2 = 0.;
  return 1;
}
__device__ inline int rs_zbrent(const RsEuler& r, double x1, double x2, double tol, double* ans) {
  const double EPS = PION_MACHINEACCURACY;
  double a = x1, b = x2, c = x2, d = 0., e = 0., min1, min2;
  double fa = rs_root_function(r, a), fb = rs_root_function(r, b), fc, p, q, rr, sv, tol1, xm;
  if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) return 1;
  fc = fb;
  for (int iter = 1; iter <= 100; iter++) {
    if ((fb > 0.0 && fc > 0.0) || (fb < 0.0 && fc < 0.0)) { c = a; fc = fa; e = d = b - a; }
    if (fabs(fc) < fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    tol1 = 2.0 * EPS * fabs(b) + 0.5 * tol * fabs(b);
    xm = 0.5 * (c - b);
    if (fabs(xm) <= tol1 || fb == 0.0) { *ans = b; return 0; }
    if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
      sv = fb / fa;
      if (a == c) { p = 2.0 * xm * sv; q = 1.0 - sv; }
      else {
        q = fa / fc;
        rr = fb / fc;
        p = sv * (2.0 * xm * q * (q - rr) - (b - a) * (rr - 1.0));
        q = (q - 1.0) * (rr - 1.0) * (sv - 1.0);
      }
      if (p > 0.0) q = -q;
      p = fabs(p);
      min1 = 3.0 * xm * q - fabs(tol1 * q);
      min2 = fabs(e * q);
      if (2.0 * p < (min1 < min2 ? min1 : min2)) { e = d; d = p / q; }
      else { d = xm; e = d; }
    } else { d = xm; e = d; }
    a = b;
    fa = fb;
    if (fabs(d) > tol1) b += d;
    else b += ((xm) >= 0.0 ? fabs(tol1) : -fabs(tol1));
    fb = rs_root_function(r, b);
  }
  return 1;
}
__device__ inline void rs_check_wave_locations(RsEuler& r) {
  const double g = r.g;
  double* ps = r.ps;
  const double *L = r.L, *R = r.R;
  if (ps[1] < L[1]) {
    if (L[2] >= r.cl) { ps[1] = L[1]; ps[0] = L[0]; ps[2] = L[2]; return; }
    else if (ps[2] > 0.) {
      const double cstar = sqrt(g * ps[1] / ps[0]);
      if (ps[2] > cstar) {
        ps[2] = (2. * r.cl + L[2] * (g - 1.)) / (g + 1.);
        ps[0] = L[0] * exp(2. / (g - 1.) * log(ps[2] / r.cl));
        ps[1] = exp(g * log(ps[0] / L[0])) * L[1];
        return;
      }
    }
  }
  if (ps[1] < R[1]) {
    if (R[2] <= -r.cr) { ps[1] = R[1]; ps[0] = R[0]; ps[2] = R[2]; return; }
    else if (ps[2] < 0.) {
      const double cstar = sqrt(g * ps[1] / ps[0]);
      if (ps[2] < -cstar) {
        ps[2] = (-2. * r.cr + R[2] * (g - 1.)) / (g + 1.);
        ps[0] = R[0] * exp(2. / (g - 1.) * log(-ps[2] / r.cr));
        ps[1] = exp(g * log(ps[0] / R[0])) * R[1];
        return;
      }
    }
  }
  if (ps[1] > 1.0000001 * R[1]) {
    const double vsh = R[2] + (ps[1] / R[1] - 1.) * r.cr * r.cr / g / (ps[2] - R[2]);
    if (vsh < 0.) { ps[1] = R[1]; ps[0] = R[0]; ps[2] = R[2]; return; }
  }
  if (ps[1] > 1.0000001 * L[1]) {
    const double vsh = L[2] + (ps[1] / L[1] - 1.) * r.cl * r.cl / g / (ps[2] - L[2]);
    if (vsh > 0.) { ps[1] = L[1]; ps[0] = L[0]; ps[2] = L[2]; return; }
  }
}
__device__ inline int rs_linear_solver(RsEuler& r) {
  double m[3];
  for (int i = 0; i < 3; i++) m[i] = (r.L[i] + r.R[i]) / 2.;
  const double mcs = sqrt(r.g * m[1] / m[0]);
  double* ps = r.ps;
  const double *L = r.L, *R = r.R;
  if (m[2] - mcs >= 0.) { for (int i = 0; i < 3; i++) ps[i] = L[i]; return 0; }
  else if (m[2] + mcs <= 0.) { for (int i = 0; i < 3; i++) ps[i] = R[i]; return 0; }
  ps[1] = 0.5 * (L[1] + R[1] - m[0] * mcs * (R[2] - L[2]));
  ps[2] = 0.5 * (L[2] + R[2] - (R[1] - L[1]) / m[0] / mcs);
  if (fabs(ps[2] / mcs) <= 1.e-6) ps[0] = m[0] * (2. + (L[2] - R[2]) / mcs) / 2.;
  else if (ps[2] > 0) ps[0] = L[0] + m[0] * (L[2] - ps[2]) / mcs;
  else if (ps[2] < 0) ps[0] = R[0] + m[0] * (ps[2] - R[2]) / mcs;
  else return 1;
  return 0;
}
__device__ inline int rs_exact_solver(RsEuler& r) {
  const double g = r.g;
  double* ps = r.ps;
  int err = 0;
  {
    double x1 = (r.L[1] + r.R[1]) / 6.0, x2 = x1 * 9.0;
    if (rs_bracket_root_pos(r, &x1, &x2)) { ps[1] = -1.0; err += 1; }
    else if (rs_zbrent(r, x1, x2, 1.0e-8, &ps[1])) { ps[1] = -1.0; err += 1; }
  }
  rs_hydro_wave_full(g, 0, ps[1], r.L, &ps[2], &ps[0]);
  double rhostar, temp;
  if ((ps[2] > 0) && (fabs(ps[2] / r.cr) > 1.e-6)) rs_hydro_wave_full(g, 0, ps[1], r.L, &temp, &rhostar);
  else if ((ps[2] < 0) && (fabs(ps[2] / r.cr) > 1.e-6)) rs_hydro_wave_full(g, 1, ps[1], r.R, &temp, &rhostar);
  else if (fabs(ps[2] / r.cr) <= 1.e-6) {
    rs_hydro_wave_full(g, 0, ps[1], r.L, &temp, &rhostar);
    rs_hydro_wave_full(g, 1, ps[1], r.R, &temp, &ps[0]);
    rhostar = (rhostar + ps[0]) / 2.0;
  } else { ps[0] = -1.0; return 1; }
  ps[0] = rhostar;
  if (err != 0) { ps[1] = ps[0] = ps[2] = -1.9; return 1; }
  rs_check_wave_locations(r);
  return 0;
}
__device__ inline int rs_solve_rarerare(RsEuler& r) {
  const double g = r.g, cl = r.cl, cr = r.cr;
  double* ps = r.ps;
  const double *L = r.L, *R = r.R;
  ps[1] = pow((cl + cr - (g - 1.) / 2. * (R[2] - L[2])) /
                  ((cl * exp(-(g - 1.) / 2. / g * log(L[1]))) + (cr * exp(-(g - 1.) / 2. / g * log(R[1])))),
              2. * g / (g - 1.));
  ps[2] = L[2] + 2. * cl / (g - 1.) * (1. - exp((g - 1.) / 2. / g * log(ps[1] / L[1])));
  if ((ps[2] > 0) && (fabs(ps[2] / cr) > 1.e-6)) ps[0] = L[0] * exp(log(ps[1] / L[1]) / g);
  else if ((ps[2] < 0) && (fabs(ps[2] / cr) > 1.e-6)) ps[0] = R[0] * exp(log(ps[1] / R[1]) / g);
  else if (fabs(ps[2] / cr) <= 1.e-6) ps[0] = ((R[0] * exp(log(ps[1] / R[1]) / g)) + (L[0] * exp(log(ps[1] / L[1]) / g))) / 2.0;
  else { ps[0] = -1.0; return 1; }
  rs_check_wave_locations(r);
  return 0;
}
// copy_lr: 0 = the three scalars were set, 1 = pstar := whole left state, 2 = whole right state (their v_t too)
__device__ inline int rs_solve_cavitation(RsEuler& r, const double* refvec_ro_pg_vn) {
  const double g = r.g, cl = r.cl, cr = r.cr;
  double* ps = r.ps;
  const double *L = r.L, *R = r.R;
  if ((L[2] - cl) >= 0.) { for (int i = 0; i < 3; i++) ps[i] = L[i]; return 0; }
  const double temp = 2. / (g - 1.);
  if ((L[2] + temp * cl) >= 0.) {
    ps[2] = (2. * cl + L[2] * (g - 1.)) / (g + 1.);
    ps[0] = L[0] * exp(2. / (g - 1.) * log(ps[2] / cl));
    ps[1] = exp(g * log(ps[0] / L[0])) * L[1];
    return 0;
  }
  if ((R[2] - temp * cr) >= 0.) {
    ps[0] = refvec_ro_pg_vn[0] * 1.e-5;  // BASEPG, constants.h:336
    ps[1] = refvec_ro_pg_vn[1] * 1.e-5;
    ps[2] = refvec_ro_pg_vn[2] * 1.e-5;
    return 0;
  }
  if ((R[2] + cr) > 0.) {
    ps[2] = (-2. * cr + R[2] * (g - 1.)) / (g + 1.);
    ps[0] = R[0] * exp(2. / (g - 1.) * log(-ps[2] / cr));
    ps[1] = exp(g * log(ps[0] / R[0])) * R[1];
    return 0;
  }
  if ((R[2] + cr) <= 0.) { for (int i = 0; i < 3; i++) ps[i] = R[i]; return 0; }
  return 1;
}
// JMs_riemann_solve + PtoFlux(pstar); ax = sweep axis (maps the solver frame onto RefVec); returns 1 on a solver failure
template <int MODE>
__device__ inline int hydro_JMs(const Prim& l, const Prim& rg, const PhysParams& pp, int ax, Cons& flux, Prim& pstar) {
  const int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
  const double rv[5] = {pp.rs_refvec[0], pp.rs_refvec[1], pp.rs_refvec[2 + ax], pp.rs_refvec[2 + a1], pp.rs_refvec[2 + a2]};
  pstar.bn = pstar.bt1 = pstar.bt2 = pstar.psi = 0.0;
  int fail = 0;
  // "same state" shortcut (:300-311)
  const double diff = fabs(rg.ro - l.ro) / (fabs(rv[0]) + PION_TINYVALUE) + fabs(rg.pg - l.pg) / (fabs(rv[1]) + PION_TINYVALUE);
  // the reference sums the five components in grid order RO, PG, VX, VY, VZ
  double dv[3];
  dv[ax] = fabs(rg.vn - l.vn) / (fabs(rv[2]) + PION_TINYVALUE);
  dv[a1] = fabs(rg.vt1 - l.vt1) / (fabs(rv[3]) + PION_TINYVALUE);
  dv[a2] = fabs(rg.vt2 - l.vt2) / (fabs(rv[4]) + PION_TINYVALUE);
  if (((diff + dv[0]) + dv[1]) + dv[2] < 1.e-6) {
    pstar.ro = (l.ro + rg.ro) / 2.; pstar.pg = (l.pg + rg.pg) / 2.; pstar.vn = (l.vn + rg.vn) / 2.;
    pstar.vt1 = (l.vt1 + rg.vt1) / 2.; pstar.vt2 = (l.vt2 + rg.vt2) / 2.;
  } else {
    RsEuler r;
    r.g = pp.gamma;
    r.L[0] = l.ro; r.L[1] = l.pg; r.L[2] = l.vn;
    r.R[0] = rg.ro; r.R[1] = rg.pg; r.R[2] = rg.vn;
    r.ps[0] = r.ps[1] = r.ps[2] = 0.0;
    const double g = pp.gamma;
    r.cl = sqrt(g * l.pg / l.ro);
    r.cr = sqrt(g * rg.pg / rg.ro);
    int err = 0;
    if ((rg.vn - l.vn) <= 2. * (r.cl + sqrt((g - 1.) / 2. / g) * r.cr) / (g - 1.)) {
      if (MODE == SOLVE_RSLINEAR) {
        err = rs_linear_solver(r);
        if (err) { r.ps[1] = r.ps[0] = r.ps[2] = PION_TINYVALUE; fail = 1; }
      } else if (MODE == SOLVE_RSEXACT) {
        err = rs_exact_solver(r);
        if (err) { r.ps[1] = r.ps[0] = PION_TINYVALUE; fail = 1; }
      } else {
        err = rs_linear_solver(r);
        if (err) r.ps[1] = r.ps[0] = r.ps[2] = PION_TINYVALUE;
        if (err != 0 || !((pmax(l.pg, rg.pg) / pmin(l.pg, rg.pg) < 1.4) && (pmax(l.ro, rg.ro) / pmin(l.ro, rg.ro) < 1.4) &&
                          (fabs(rg.vn - l.vn) / pmin(r.cl, r.cr) < 0.03))) {
          err = rs_exact_solver(r);
          if (err) { r.ps[1] = r.ps[0] = PION_TINYVALUE; fail = 1; }
        }
      }
    } else if ((rg.vn - l.vn) <= 2. * (r.cl + r.cr) / (g - 1.)) {
      err = rs_solve_rarerare(r);
      if (err) { r.ps[1] = r.ps[0] = r.ps[2] = -1.9e99; fail = 1; }
    } else {
      err = rs_solve_cavitation(r, rv);
      if (err) { r.ps[1] = r.ps[0] = r.ps[2] = -1.9e100; fail = 1; }
    }
    pstar.ro = r.ps[0]; pstar.pg = r.ps[1]; pstar.vn = r.ps[2];
    if (!fail) {
      // v_t only changes across the contact (:433-441) -- also after the whole-state copies of the linear /
      // cavitation branches, which the reference overwrites here in the same way
      if (pstar.vn > 0) { pstar.vt1 = l.vt1; pstar.vt2 = l.vt2; }
      else { pstar.vt1 = rg.vt1; pstar.vt2 = rg.vt2; }
      if (pstar.pg <= PION_TINYVALUE) pstar.pg = 1.e-5 * rv[1];
      if (pstar.ro <= PION_TINYVALUE) pstar.ro = 1.e-5 * rv[0];
    } else {
      pstar.vt1 = pstar.vt2 = 0.0;
    }
  }
  Cons u;
  PtoU<EQ_EULER>(pstar, u, pp.gamma - 1.0);
  PUtoFlux<EQ_EULER>(pstar, u, flux);
  return fail;
}

// ---------------------------------------------------------------------------
// riemann_MHD: the linear MHD Riemann solver (solverType 1 with the MHD equations; Falle, Komissarov & Joarder 1998 with the
// Roe & Balsara eigenvector normalisation): JMs_riemann_solve mode 1 (Riemann_solvers/riemannMHD.cpp:165-400), get_sound_speeds
// (:555-765), get_eigenvalues (:768-777), RoeBalsara_evectors (:965-1115), calculate_wave_strengths (:813-846), get_pstar
// (:849-960), then PtoFlux(pstar).  Solver-frame variables in the reference's order RRO, RPG, RVX, RVY, RVZ, RBY, RBZ (the Prim
// fields ro, pg, vn, vt1, vt2, bt1, bt2; bn is the parameter B_x).  Plain IEEE arithmetic as written in the reference: gather
// kernel only, not a performance path.  Returns 1 where the reference calls rep.error.
// pp.rs_refvec = {RefVec[RO], RefVec[PG], 0.1 cfast(RefVec), |B(RefVec)|} (eqns_mhd_ideal::SetAvgState, built on the host).
// ---------------------------------------------------------------------------
__device__ inline int mhd_JMs_linear(const Prim& l, const Prim& r, const PhysParams& pp, Cons& flux, Prim& pstar) {
  enum { RRO = 0, RPG = 1, RVX = 2, RVY = 3, RVZ = 4, RBY = 5, RBZ = 6 };
  enum { FN = 0, AN = 1, SN = 2, CT = 3, SP = 4, AP = 5, FP = 6 };
  const double g = pp.gamma;
  const double smallB = PION_MACHINEACCURACY, tinyB = smallB * smallB * smallB;
  const double L[7] = {l.ro, l.pg, l.vn, l.vt1, l.vt2, l.bt1, l.bt2}, R[7] = {r.ro, r.pg, r.vn, r.vt1, r.vt2, r.bt1, r.bt2};
  double M[7], ps[7];
#pragma unroll
  for (int v = 0; v < 7; v++) M[v] = 0.5 * (L[v] + R[v]);
  const double ansBX = 0.5 * (l.bn + r.bn);
  const double refn[7] = {pp.rs_refvec[0], pp.rs_refvec[1], pp.rs_refvec[2], pp.rs_refvec[2], pp.rs_refvec[2], pp.rs_refvec[3], pp.rs_refvec[3]};
  double diff = 0.;
#pragma unroll
  for (int i = 0; i < 7; i++) diff += fabs(R[i] - L[i]) / (fabs(refn[i]) + PION_TINYVALUE);
  int fail = 0;
  if (diff < 1.e-6) {  // same-state shortcut (:227-268)
#pragma unroll
    for (int v = 0; v < 7; v++) ps[v] = M[v];
  } else {
    const double sro = sqrt(M[RRO]);
    const double ch = sqrt(g * M[RPG] / M[RRO]);
    const double bx = ansBX / sro;
    const double ca = fabs(bx);
    const double bt = sqrt((M[RBY] * M[RBY] + M[RBZ] * M[RBZ]) / M[RRO]);
    double betay, betaz;
    if (bt > tinyB) { betay = M[RBY] / sro / bt; betaz = M[RBZ] / sro / bt; }
    else { betay = 1. / sqrt(2.); betaz = 1. / sqrt(2.); }
    if ((ch / ((ca < bt) ? bt : ca)) < sqrt(smallB)) fail = 1;
    double temp1 = ch * ch + bx * bx + bt * bt;
    double temp2 = 4. * ch * ch * bx * bx;
    if ((temp2 = temp1 * temp1 - temp2) < PION_MACHINEACCURACY) temp2 = PION_MACHINEACCURACY;
    double cf = sqrt((temp1 + sqrt(temp2)) / 2.);
    if ((temp2 = temp1 - sqrt(temp2)) < PION_MACHINEACCURACY) temp2 = PION_MACHINEACCURACY;
    double cs = sqrt(temp2 / 2.);
    if (cs > ch) cs = ch - smallB;
    if (ch > cf) cf = ch + smallB;
    if (cs > ca) cs = ca - smallB;
    if (cs <= 0. || cs > ca) cs = ca / 2.;
    if (ca > cf) cf = ca + smallB;
    double alphaf = 0., alphas = 0., cf2diff;
    if ((cf2diff = cf * cf - cs * cs) > smallB) {
      if ((alphaf = ch * ch - cs * cs) <= smallB) alphaf = 0.;
      if ((alphas = cf * cf - ch * ch) <= smallB) alphas = 0.;
      if ((alphaf = sqrt(alphaf / cf2diff)) > 1.) alphaf = 1.;
      if ((alphas = sqrt(alphas / cf2diff)) > 1.) alphas = 1.;
    } else {
      fail = 1;  // "Near Triple degeneracy point": rep.error in the reference
    }
    if ((cf <= 0.) || (cs < 0.) || (ca < 0.) || (ch <= 0.)) fail = 1;
#pragma unroll
    for (int v = 0; v < 7; v++) ps[v] = 0.0;
    if (!fail) {
      const double ev[7] = {M[RVX] - cf, M[RVX] - ca, M[RVX] - cs, M[RVX], M[RVX] + cs, M[RVX] + ca, M[RVX] + cf};
      const double r2 = sqrt(2.);
      const double sBx = (ansBX < 0.) ? -1.0 : 1.0;
      // the three independent left eigenvectors (negative fast, Alfven, slow) and the contact; the positive ones differ
      // from them by the signs of the velocity (fast, slow) or field (Alfven) components
      double lev[7][7], rev[7][7];
#pragma unroll
      for (int w = 0; w < 7; w++)
#pragma unroll
        for (int i = 0; i < 7; i++) lev[w][i] = rev[w][i] = 0.0;
      lev[FN][RVX] = -alphaf * cf; lev[FN][RVY] = alphas * cs * sBx * betay; lev[FN][RVZ] = alphas * cs * sBx * betaz;
      lev[FN][RPG] = alphaf / M[RRO]; lev[FN][RBY] = alphas * ch * betay / sro; lev[FN][RBZ] = alphas * ch * betaz / sro;
      lev[AN][RVY] = sBx * betaz / r2; lev[AN][RVZ] = -sBx * betay / r2;
      lev[AN][RBY] = betaz / sro / r2; lev[AN][RBZ] = -betay / sro / r2;
      lev[SN][RVX] = -alphas * cs; lev[SN][RVY] = -alphaf * cf * sBx * betay; lev[SN][RVZ] = -alphaf * cf * sBx * betaz;
      lev[SN][RPG] = alphas / M[RRO]; lev[SN][RBY] = -alphaf * ch * betay / sro; lev[SN][RBZ] = -alphaf * ch * betaz / sro;
      lev[CT][RRO] = 1.; lev[CT][RPG] = -1 / ch / ch;
      lev[SP][RVX] = -lev[SN][RVX]; lev[SP][RVY] = -lev[SN][RVY]; lev[SP][RVZ] = -lev[SN][RVZ];
      lev[SP][RPG] = lev[SN][RPG]; lev[SP][RBY] = lev[SN][RBY]; lev[SP][RBZ] = lev[SN][RBZ];
      lev[AP][RVY] = lev[AN][RVY]; lev[AP][RVZ] = lev[AN][RVZ]; lev[AP][RBY] = -lev[AN][RBY]; lev[AP][RBZ] = -lev[AN][RBZ];
      lev[FP][RVX] = -lev[FN][RVX]; lev[FP][RVY] = -lev[FN][RVY]; lev[FP][RVZ] = -lev[FN][RVZ];
      lev[FP][RPG] = lev[FN][RPG]; lev[FP][RBY] = lev[FN][RBY]; lev[FP][RBZ] = lev[FN][RBZ];
      rev[FN][RRO] = alphaf * M[RRO]; rev[FN][RVX] = lev[FN][RVX]; rev[FN][RVY] = lev[FN][RVY]; rev[FN][RVZ] = lev[FN][RVZ];
      rev[FN][RPG] = alphaf * M[RRO] * ch * ch; rev[FN][RBY] = lev[FN][RBY] * M[RRO]; rev[FN][RBZ] = lev[FN][RBZ] * M[RRO];
      rev[AN][RVY] = lev[AN][RVY]; rev[AN][RVZ] = lev[AN][RVZ]; rev[AN][RBY] = lev[AN][RBY] * M[RRO]; rev[AN][RBZ] = lev[AN][RBZ] * M[RRO];
      rev[SN][RRO] = alphas * M[RRO]; rev[SN][RVX] = lev[SN][RVX]; rev[SN][RVY] = lev[SN][RVY]; rev[SN][RVZ] = lev[SN][RVZ];
      rev[SN][RPG] = alphas * M[RRO] * ch * ch; rev[SN][RBY] = lev[SN][RBY] * M[RRO]; rev[SN][RBZ] = lev[SN][RBZ] * M[RRO];
      rev[CT][RRO] = 1.0;
      rev[SP][RRO] = rev[SN][RRO]; rev[SP][RVX] = -rev[SN][RVX]; rev[SP][RVY] = -rev[SN][RVY]; rev[SP][RVZ] = -rev[SN][RVZ];
      rev[SP][RPG] = rev[SN][RPG]; rev[SP][RBY] = rev[SN][RBY]; rev[SP][RBZ] = rev[SN][RBZ];
      rev[AP][RVY] = rev[AN][RVY]; rev[AP][RVZ] = rev[AN][RVZ]; rev[AP][RBY] = -rev[AN][RBY]; rev[AP][RBZ] = -rev[AN][RBZ];
      rev[FP][RRO] = rev[FN][RRO]; rev[FP][RVX] = -rev[FN][RVX]; rev[FP][RVY] = -rev[FN][RVY]; rev[FP][RVZ] = -rev[FN][RVZ];
      rev[FP][RPG] = rev[FN][RPG]; rev[FP][RBY] = rev[FN][RBY]; rev[FP][RBZ] = rev[FN][RBZ];
      const double a22 = 1. / (2. * ch * ch);
#pragma unroll
      for (int i = 0; i < 7; i++) { lev[FN][i] *= a22; lev[SN][i] *= a22; lev[SP][i] *= a22; lev[FP][i] *= a22; }
      double pdiff[7], str[7];
#pragma unroll
      for (int i = 0; i < 7; i++) pdiff[i] = R[i] - L[i];
#pragma unroll
      for (int w = 0; w < 7; w++) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 7; i++) t += lev[w][i] * pdiff[i];
        str[w] = t;
      }
      // get_pstar: the eigenvalues are ordered, so "while (ev[i] < 0)" crosses the first n waves
#pragma unroll
      for (int j = 0; j < 7; j++) ps[j] = L[j];
      bool go = true;
#pragma unroll
      for (int w = 0; w < 7; w++) {
        go = go && (ev[w] < 0.);
        if (go) {
#pragma unroll
          for (int j = 0; j < 7; j++) ps[j] += str[w] * rev[w][j];
        }
      }
      if (fabs(M[RVX]) < (1.e-4 * ch)) {  // (nearly) stationary contact: average with the state reached from the right
#pragma unroll
        for (int j = 0; j < 7; j++) pdiff[j] = R[j];
        go = true;
#pragma unroll
        for (int w = 6; w >= 0; w--) {
          go = go && (ev[w] > 0.);
          if (go) {
#pragma unroll
            for (int j = 0; j < 7; j++) pdiff[j] -= str[w] * rev[w][j];
          }
        }
#pragma unroll
        for (int v = 0; v < 7; v++) ps[v] = 0.5 * (ps[v] + pdiff[v]);
      }
      if (ps[RPG] < 0.) ps[RPG] = refn[RPG] * 1.e-5;  // BASEPG, constants.h:336
      if (ps[RRO] < 0.) ps[RRO] = refn[RRO] * 1.e-5;
    }
  }
  pstar.ro = ps[RRO]; pstar.pg = ps[RPG]; pstar.vn = ps[RVX]; pstar.vt1 = ps[RVY]; pstar.vt2 = ps[RVZ];
  pstar.bt1 = ps[RBY]; pstar.bt2 = ps[RBZ]; pstar.bn = ansBX; pstar.psi = 0.0;
  Cons u;
  PtoU_mhd_ideal(pstar, u, g - 1.0);
  PUtoFlux<EQ_MHD>(pstar, u, flux);
  return fail;
}

// HLLD_MHD::HLLD_signal_speeds (HLLD_MHD.cpp:342-368); Bx is the same on both
// sides in the GLM case but the formula is kept general.
__device__ __forceinline__ void hlld_speeds(const Prim& L, const Prim& R, double g, double& Sl, double& Sr) {
  double Bx = 0.5 * (L.bn + R.bn);
#ifdef PION_STRICT
  double cf_l = cfast_components(L.ro, L.pg, Bx, L.bt1, L.bt2, g);
  double cf_r = cfast_components(R.ro, R.pg, Bx, R.bt1, R.bt2, g);
  double cf_max = pmax(cf_l, cf_r);
#else
  double cf_max = fast_sqrt_pos(pmax(cfast2_ir(fast_rcp(L.ro), L.pg, Bx, L.bt1, L.bt2, g),
                                     cfast2_ir(fast_rcp(R.ro), R.pg, Bx, R.bt1, R.bt2, g)));
#endif
  Sl = pmin(L.vn, R.vn) - cf_max;
  Sr = pmax(L.vn, R.vn) + cf_max;
}

// HLLD_MHD::MHD_HLL_flux_solver (HLLD_MHD.cpp:377-417)
template <bool NEED_USTAR>
__device__ __forceinline__ void mhd_HLL(const Prim& L, const Prim& R, const PhysParams& pp, Cons& flux, Cons& ustar) {
  const double gm1 = pp.gamma - 1.0;
  Cons UL, UR, FL, FR;
  PtoU_mhd_ideal(L, UL, gm1);
  PtoU_mhd_ideal(R, UR, gm1);
  PUtoFlux<EQ_MHD>(L, UL, FL);
  PUtoFlux<EQ_MHD>(R, UR, FR);
  double l0, l1;
  hlld_speeds(L, R, pp.gamma, l0, l1);
  double idl = fast_rcp(l1 - l0);
#define PION_HLLM_COMP(c)                                                                         \
  flux.c = (l0 > 0.0) ? FL.c : (l1 < 0.0) ? FR.c : (l1 * FL.c - l0 * FR.c + l1 * l0 * (UR.c - UL.c)) * idl; \
  if (NEED_USTAR) ustar.c = (l0 > 0.0) ? UL.c : (l1 < 0.0) ? UR.c : (l1 * UR.c - l0 * UL.c - FR.c + FL.c) * idl;
  PION_HLLM_COMP(rho) PION_HLLM_COMP(erg) PION_HLLM_COMP(mn) PION_HLLM_COMP(mt1) PION_HLLM_COMP(mt2)
  PION_HLLM_COMP(bbn) PION_HLLM_COMP(bbt1) PION_HLLM_COMP(bbt2)
#undef PION_HLLM_COMP
  flux.psi = 0.0;
  if (NEED_USTAR) ustar.psi = 0.0;
}

// HLLD_MHD::MHD_HLLD_flux_solver (HLLD_MHD.cpp:124-333), Miyoshi & Kusano 2005.
//
// The reference forms UL, UR, FL, FR and all four intermediate states, then picks
// one of six fan regions.  Only ONE side's states ever reach the flux, so this
// version computes the two-sided scalars the wave speeds and the double-star
// averages need (rho*, v_t*, B_t*, sqrt(rho*) of both sides), picks the side K of
// the contact (lam2 >= 0 -> left), and forms U_K, F_K, U_K*, U_K** for that side only:
//   F = F_K + c2 (U* - U_K) + c1 (U** - U*),   c2 = 0 in the outer region else the outer
//   speed (lam0 | lam4),  c1 = inner speed (lam1 | lam3) in the double-star region else 0,
// which is the reference's  F_K + lam_in U** - (lam_in - lam_out) U* - lam_out U_K
// regrouped.  The 0/0 guards (isfinite, HLLD_MHD.cpp:189-224) and the exact BX==0
// special case (:248) are kept.  `pstar` (NEED_PSTAR) is the primitive form of the
// selected region's state, i.e. UtoP(ustar) of solver_eqn_mhd_adi.cpp:183 without the
// round trip through conserved variables (v = (rho v)/rho); only ro, v and B_t are set
// (all that AVFalle reads, solver_eqn_mhd_adi.cpp:254-284).
// SAME_BN: both states carry the same normal field (GLM hands the Dedner star value to both sides,
// solver_eqn_mhd_adi.cpp:736-738).  0.5 (b + b) == b exactly, so BX is that value itself and the side's own
// B_n is BX too: identical results, a dozen FP64 instructions fewer per interface.
template <bool NEED_PSTAR, bool SAME_BN = false>
__device__ __forceinline__ void mhd_HLLD(const Prim& L, const Prim& R, const PhysParams& pp, Cons& flux, Prim& pstar) {
  const double g = pp.gamma, gm1 = g - 1.0;
  const double BX = SAME_BN ? L.bn : 0.5 * (L.bn + R.bn);
  const double BX2 = BX * BX;
  // HLLD_signal_speeds (:342-368)
  const double cf_max = fast_sqrt_pos(pmax(cfast2_ir(fast_rcp(L.ro), L.pg, BX, L.bt1, L.bt2, g),
                                           cfast2_ir(fast_rcp(R.ro), R.pg, BX, R.bt1, R.bt2, g)));
  const double lam0 = pmin(L.vn, R.vn) - cf_max;
  const double lam4 = pmax(L.vn, R.vn) + cf_max;
  const double sl_vl = lam0 - L.vn, sr_vr = lam4 - R.vn;
  // magnetic pressures with the SAME rounding sequence on both sides (explicit, never contracted: the compiler
  // would otherwise share sub-expressions with the fast speeds on one side only) -- see lam2 below
  const double pm_l = 0.5 * __dadd_rn(__dadd_rn(__dmul_rn(L.bn, L.bn), __dmul_rn(L.bt1, L.bt1)), __dmul_rn(L.bt2, L.bt2));
  const double pm_r = 0.5 * __dadd_rn(__dadd_rn(__dmul_rn(R.bn, R.bn), __dmul_rn(R.bt1, R.bt1)), __dmul_rn(R.bt2, R.bt2));
  const double tp_l = __dadd_rn(L.pg, pm_l), tp_r = __dadd_rn(R.pg, pm_r);
  const double rsl = L.ro * sl_vl, rsr = R.ro * sr_vr;
  const double itemp = fast_rcp(rsr - rsl);
  // The two momentum terms are rounded SEPARATELY (no FMA contraction): at a reflecting wall L is the mirror
  // image of R, both products are then the same number and the contact speed is an exact zero, as in the
  // reference (HLLD_MHD.cpp:159).  The left/right choice below hinges on its sign, and in ideal MHD the two
  // star fluxes differ at O(B_n B_t) there (each side's F_K carries its own B_n, which flips across the wall),
  // so a contracted product's rounding residue would pick the other side (measured: 1e-3 relative in v_t, B_t).
  const double lam2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(sr_vr, __dmul_rn(R.ro, R.vn)), -__dmul_rn(sl_vl, __dmul_rn(L.ro, L.vn))), -tp_r), tp_l) * itemp;
  const double tp_s = (rsr * tp_l - rsl * tp_r + L.ro * R.ro * sr_vr * sl_vl * (R.vn - L.vn)) * itemp;
  const double sl_sm = lam0 - lam2, sr_sm = lam4 - lam2;
  const double isl_sm = fast_rcp(sl_sm), isr_sm = fast_rcp(sr_sm);
  const double rho_ls = rsl * isl_sm, rho_rs = rsr * isr_sm;
  // transverse velocity and field of the single-star states, both sides
  const double id_l = fast_rcp(rsl * sl_sm - BX2), id_r = fast_rcp(rsr * sr_sm - BX2);
  double vys_l = L.vt1, vzs_l = L.vt2, vys_r = R.vt1, vzs_r = R.vt2;
  double bys_l = 0.0, bzs_l = 0.0, bys_r = 0.0, bzs_r = 0.0;
  {
    const double ql = (lam2 - L.vn) * id_l, qr = (lam2 - R.vn) * id_r;
    if (isfinite(ql)) { vys_l = L.vt1 - BX * L.bt1 * ql; vzs_l = L.vt2 - BX * L.bt2 * ql; }
    if (isfinite(qr)) { vys_r = R.vt1 - BX * R.bt1 * qr; vzs_r = R.vt2 - BX * R.bt2 * qr; }
    const double pl = (rsl * sl_vl - BX2) * id_l, pr = (rsr * sr_vr - BX2) * id_r;
    if (isfinite(pl)) { bys_l = L.bt1 * pl; bzs_l = L.bt2 * pl; }
    if (isfinite(pr)) { bys_r = R.bt1 * pr; bzs_r = R.bt2 * pr; }
  }
  const double isq_l = fast_rsqrt(rho_ls), isq_r = fast_rsqrt(rho_rs);
  const double sq_l = rho_ls * isq_l, sq_r = rho_rs * isq_r;
  const double aBX = fabs(BX);
  const double lam1 = lam2 - aBX * isq_l;
  const double lam3 = lam2 + aBX * isq_r;

  // side and region of the interface (the reference's if-chain, first true wins)
  const bool left = (lam0 > 0) || (lam1 >= 0) || (lam2 >= 0);
  const bool outer = left ? (lam0 > 0) : !(lam3 >= 0 || lam4 >= 0);
  const bool dstar = left ? !(lam0 > 0 || lam1 >= 0) : (lam3 >= 0);

  // double-star transverse state (shared by both sides, :248-290); equals the star state if BX == 0
  const double sgn = (double)((BX > 0) - (BX < 0));
  double vy_ss, vz_ss, by_ss, bz_ss;
  {
    const double itsum = fast_rcp(sq_l + sq_r);
    vy_ss = (sq_l * vys_l + sq_r * vys_r + (bys_r - bys_l) * sgn) * itsum;
    vz_ss = (sq_l * vzs_l + sq_r * vzs_r + (bzs_r - bzs_l) * sgn) * itsum;
    by_ss = (sq_l * bys_r + sq_r * bys_l + sq_l * sq_r * (vys_r - vys_l) * sgn) * itsum;
    bz_ss = (sq_l * bzs_r + sq_r * bzs_l + sq_l * sq_r * (vzs_r - vzs_l) * sgn) * itsum;
  }

  // everything below is for side K only
  const double K_ro = (left ? L.ro : R.ro), K_pg = (left ? L.pg : R.pg), K_vn = (left ? L.vn : R.vn);
  const double K_vt1 = (left ? L.vt1 : R.vt1), K_vt2 = (left ? L.vt2 : R.vt2);
  const double K_bn = SAME_BN ? BX : (left ? L.bn : R.bn), K_bt1 = (left ? L.bt1 : R.bt1), K_bt2 = (left ? L.bt2 : R.bt2);
  const double s_v = (left ? sl_vl : sr_vr), is_m = (left ? isl_sm : isr_sm);
  const double rho_s = (left ? rho_ls : rho_rs), sq_K = (left ? -sq_l : sq_r);  // sign of the ** energy jump folded in
  const double vys = (left ? vys_l : vys_r), vzs = (left ? vzs_l : vzs_r);
  const double bys = (left ? bys_l : bys_r), bzs = (left ? bzs_l : bzs_r);
  const double pm_K = (left ? pm_l : pm_r), tp_K = (left ? tp_l : tp_r);
  const double c2 = outer ? 0.0 : (left ? lam0 : lam4);
  const double c1 = (dstar && BX != 0) ? (left ? lam1 : lam3) : 0.0;

  // U_K (eqns_mhd_ideal::PtoU) and F_K (PUtoFlux)
  const double K_mn = K_ro * K_vn, K_mt1 = K_ro * K_vt1, K_mt2 = K_ro * K_vt2;
  const double K_erg = (K_ro * (K_vn * K_vn + K_vt1 * K_vt1 + K_vt2 * K_vt2) * 0.5) + PION_OVER_GM1(K_pg, gm1) + pm_K;
  const double vb_own = K_vn * K_bn + K_vt1 * K_bt1 + K_vt2 * K_bt2;
  // star-state energy (:226-239) -- note v.B there is taken with BX, not the side's own B_n
  const double vb_K = K_vn * BX + K_vt1 * K_bt1 + K_vt2 * K_bt2;
  const double vbs_K = lam2 * BX + vys * bys + vzs * bzs;
  const double erg_s = (s_v * K_erg - tp_K * K_vn + tp_s * lam2 + BX * (vb_K - vbs_K)) * is_m;
  const double bv = lam2 * BX + vy_ss * by_ss + vz_ss * bz_ss;
  const double derg_ss = sq_K * (vbs_K - bv) * sgn;  // U**.erg - U*.erg

  flux.rho = K_mn + c2 * (rho_s - K_ro);
  flux.mn = (K_mn * K_vn + K_pg + pm_K - K_bn * K_bn) + c2 * (lam2 * rho_s - K_mn);
  // The ** jumps are only FORMED in the ** region (c1 != 0): beyond a strong rarefaction the star density of the side that is
  // NOT used can be negative, its square root -- and with it every ** quantity -- a NaN, and 0 x NaN is not 0.  The reference
  // never evaluates the ** states outside that region (HLLD_MHD.cpp:292-333).  Found by the negative-pressure-reset probe
  // (plasma beta 1e-7, tools/floor_probe.py); same arithmetic, bit for bit, wherever every quantity is finite (checked on
  // 4e5 interfaces with a host build of this header).
  const bool in_ss = (c1 != 0.0);
  const double j_vy = in_ss ? (vy_ss - vys) : 0.0, j_vz = in_ss ? (vz_ss - vzs) : 0.0;
  const double j_by = in_ss ? (by_ss - bys) : 0.0, j_bz = in_ss ? (bz_ss - bzs) : 0.0, j_erg = in_ss ? derg_ss : 0.0;
  flux.mt1 = (K_mn * K_vt1 - K_bn * K_bt1) + c2 * (vys * rho_s - K_mt1) + c1 * (j_vy * rho_s);
  flux.mt2 = (K_mn * K_vt2 - K_bn * K_bt2) + c2 * (vzs * rho_s - K_mt2) + c1 * (j_vz * rho_s);
  flux.erg = (K_vn * (K_erg + K_pg + pm_K) - K_bn * vb_own) + c2 * (erg_s - K_erg) + c1 * j_erg;
  flux.bbn = c2 * (BX - K_bn);
  flux.bbt1 = (K_vn * K_bt1 - K_vt1 * K_bn) + c2 * (bys - K_bt1) + c1 * j_by;
  flux.bbt2 = (K_vn * K_bt2 - K_vt2 * K_bn) + c2 * (bzs - K_bt2) + c1 * j_bz;
  flux.psi = 0.0;

  if (NEED_PSTAR) {
    const bool dd = dstar && (BX != 0);  // in the ** region with BX==0 the state is the * state
    pstar.ro = outer ? K_ro : rho_s;
    pstar.vn = outer ? K_vn : lam2;
    pstar.vt1 = outer ? K_vt1 : (dd ? vy_ss : vys);
    pstar.vt2 = outer ? K_vt2 : (dd ? vz_ss : vzs);
    pstar.bt1 = outer ? K_bt1 : (dd ? by_ss : bys);
    pstar.bt2 = outer ? K_bt2 : (dd ? bz_ss : bzs);
    pstar.bn = outer ? K_bn : BX;
    pstar.pg = 0.0;
    pstar.psi = 0.0;
    if (pstar.ro <= 0.0) {  // UtoP's (fatal in the reference) negative-density reset, eqns_mhd_adiabatic.cpp:137-160
      const double r0 = PION_BASE_RHO * pp.refvec_ro;
      const double f = pstar.ro / r0;
      pstar.vn *= f; pstar.vt1 *= f; pstar.vt2 *= f;
      pstar.ro = r0;
    }
  }
}

// Riemann_Roe_MHD_CV::MHD_Roe_CV_flux_solver_symmetric
// (Roe_MHD_ConservedVar_solver.cpp:218-262; average state :300-352, difference
// states :358-393, wave speeds :399-470, eigenvalues + H-correction :476-510,
// wave strengths :516-580, Cargo & Gallice right eigenvectors :699-810,
// symmetric flux :1074-1131, pstar :283-295)
__device__ __forceinline__ void mhd_RoeCV(const Prim& L, const Prim& R, const PhysParams& pp, double hc_etamax, Cons& flux,
                                          Prim& pstar) {
  enum { FN = 0, AN = 1, SN = 2, CT = 3, SP = 4, AP = 5, FP = 6 };
  const double g = pp.gamma, gm1 = g - 1.0;
  Cons UL, UR;
  PtoU_mhd_ideal(L, UL, gm1);
  PtoU_mhd_ideal(R, UR, gm1);
  double rl = psqrt(L.ro), rr = psqrt(R.ro);
  double lH = (L.ro * (L.vn * L.vn + L.vt1 * L.vt1 + L.vt2 * L.vt2) / 2.0 + PION_OVER_GM1(g * L.pg, gm1) +
               (L.bn * L.bn + L.bt1 * L.bt1 + L.bt2 * L.bt2)) * fast_rcp(L.ro);
  double rH = (R.ro * (R.vn * R.vn + R.vt1 * R.vt1 + R.vt2 * R.vt2) / 2.0 + PION_OVER_GM1(g * R.pg, gm1) +
               (R.bn * R.bn + R.bt1 * R.bt1 + R.bt2 * R.bt2)) * fast_rcp(R.ro);
  double Roe_denom = fast_rcp(rl + rr);
  double m_ro = rl * rr;
  double m_vn = (rl * L.vn + rr * R.vn) * Roe_denom;
  double m_vt1 = (rl * L.vt1 + rr * R.vt1) * Roe_denom;
  double m_vt2 = (rl * L.vt2 + rr * R.vt2) * Roe_denom;
  double m_bt1 = (rr * L.bt1 + rl * R.bt1) * Roe_denom;
  double m_bt2 = (rr * L.bt2 + rl * R.bt2) * Roe_denom;
  double m_bn = 0.5 * (L.bn + R.bn);
  double signBX = (m_bn >= 0.0) ? 1.0 : -1.0;
  double m_H = (rl * lH + rr * rH) * Roe_denom;
  double Roe_V = psqrt(m_vn * m_vn + m_vt1 * m_vt1 + m_vt2 * m_vt2);
  double Roe_B = psqrt(m_bn * m_bn + m_bt1 * m_bt1 + m_bt2 * m_bt2);
  double Roe_Bt = psqrt(m_bt1 * m_bt1 + m_bt2 * m_bt2);
  double betay, betaz;
  if (Roe_Bt >= PION_TINYVALUE) {
    const double iBt = fast_rcp(Roe_Bt);
    betay = m_bt1 * iBt;
    betaz = m_bt2 * iBt;
  } else {
    betay = 1.0 / sqrt(2.0);
    betaz = 1.0 / sqrt(2.0);
  }
  // difference states
  double ud_mn = UR.mn - UL.mn, ud_mt1 = UR.mt1 - UL.mt1, ud_mt2 = UR.mt2 - UL.mt2, ud_erg = UR.erg - UL.erg;
  double pd_ro = R.ro - L.ro, pd_vn = R.vn - L.vn, pd_vt1 = R.vt1 - L.vt1, pd_vt2 = R.vt2 - L.vt2;
  double pd_bt1 = R.bt1 - L.bt1, pd_bt2 = R.bt2 - L.bt2;
  double CGX = (pd_bt1 * pd_bt1 + pd_bt2 * pd_bt2) * 0.5 * Roe_denom * Roe_denom;
  double pd_pg = ((0.5 * Roe_V * Roe_V - CGX) * pd_ro - (m_vn * ud_mn + m_vt1 * ud_mt1 + m_vt2 * ud_mt2) + ud_erg -
                  (m_bt1 * pd_bt1 + m_bt2 * pd_bt2)) * gm1;
  // wave speeds
  const double im_ro = fast_rcp(m_ro);
  double b2 = Roe_B * Roe_B * im_ro;
  double Roe_a = psqrt((2.0 - g) * CGX + gm1 * pmax((m_H - 0.5 * Roe_V * Roe_V - b2), 1.0e-12 * Roe_V * Roe_V));
  double astar2 = Roe_a * Roe_a + b2;
  double Roe_ca = psqrt(m_bn * m_bn * im_ro);
  double Roe_cs = astar2 * astar2 - 4.0 * Roe_a * Roe_a * Roe_ca * Roe_ca;
  Roe_cs = (Roe_cs <= 0.0) ? 0.0 : psqrt(Roe_cs);
  double Roe_cf = psqrt(0.5 * (astar2 + Roe_cs));
  Roe_cs = astar2 - Roe_cs;
  Roe_cs = (Roe_cs <= 0.0) ? 0.0 : psqrt(0.5 * Roe_cs);
  if (Roe_ca > Roe_cf) Roe_ca = Roe_cf;
  if (Roe_cs > Roe_ca) Roe_cs = Roe_ca;
  double cf2diff = Roe_cf * Roe_cf - Roe_cs * Roe_cs, alphaf, alphas;
  if (cf2diff > PION_MACHINEACCURACY) {
    alphaf = Roe_a * Roe_a - Roe_cs * Roe_cs;
    if (alphaf < 0.0) alphaf = 0.;
    alphas = Roe_cf * Roe_cf - Roe_a * Roe_a;
    if (alphas < 0.0) alphas = 0.;
    const double icf2 = fast_rcp(cf2diff);
    alphaf = psqrt(alphaf * icf2);
    if (alphaf > 1.0) alphaf = 1.0;
    alphas = psqrt(alphas * icf2);
    if (alphas > 1.0) alphas = 1.0;
  } else {
    alphaf = alphas = 1.0 / sqrt(2.0);
  }
  double ev[7];
  ev[FN] = m_vn - Roe_cf; ev[AN] = m_vn - Roe_ca; ev[SN] = m_vn - Roe_cs; ev[CT] = m_vn;
  ev[SP] = m_vn + Roe_cs; ev[AP] = m_vn + Roe_ca; ev[FP] = m_vn + Roe_cf;
#pragma unroll
  for (int v = 0; v < 7; v++) ev[v] = (ev[v] < 0.0) ? pmin(ev[v], -hc_etamax) : pmax(ev[v], hc_etamax);
  // wave strengths
  double rootrho = psqrt(m_ro);
  const double irootrho = fast_rcp(rootrho);
  double str[7];
  {
    double t_p = (CGX * pd_ro + pd_pg);
    double t_v = (betay * pd_vt1 + betaz * pd_vt2);
    double t_b = (betay * pd_bt1 + betaz * pd_bt2);
    str[FN] = 0.5 * (alphaf * t_p + m_ro * alphas * Roe_cs * signBX * t_v - m_ro * alphaf * Roe_cf * pd_vn +
                     rootrho * alphas * Roe_a * t_b);
    str[FP] = 0.5 * (alphaf * t_p - m_ro * alphas * Roe_cs * signBX * t_v + m_ro * alphaf * Roe_cf * pd_vn +
                     rootrho * alphas * Roe_a * t_b);
    str[SN] = 0.5 * (alphas * t_p - m_ro * alphaf * Roe_cf * signBX * t_v - m_ro * alphas * Roe_cs * pd_vn -
                     rootrho * alphaf * Roe_a * t_b);
    str[SP] = 0.5 * (alphas * t_p + m_ro * alphaf * Roe_cf * signBX * t_v + m_ro * alphas * Roe_cs * pd_vn -
                     rootrho * alphaf * Roe_a * t_b);
    str[AN] = 0.5 * (+betay * pd_vt2 - betaz * pd_vt1 + signBX * (betay * pd_bt2 - betaz * pd_bt1) * irootrho);
    str[AP] = 0.5 * (-betay * pd_vt2 + betaz * pd_vt1 + signBX * (betay * pd_bt2 - betaz * pd_bt1) * irootrho);
    str[CT] = (Roe_a * Roe_a - CGX) * pd_ro - pd_pg;
  }
  // right eigenvectors, component order {rho, mn, mt1, mt2, bt1, bt2, e}
  double rev[7][7];
  double ia2 = fast_rcp(Roe_a * Roe_a);
  rev[CT][0] = ia2; rev[CT][1] = m_vn * ia2; rev[CT][2] = m_vt1 * ia2; rev[CT][3] = m_vt2 * ia2;
  rev[CT][4] = 0.0; rev[CT][5] = 0.0;
  rev[CT][6] = (0.5 * Roe_V * Roe_V + PION_OVER_GM1(CGX * (g - 2), gm1)) * ia2;
  rev[AN][0] = 0.0; rev[AN][1] = 0.0;
  rev[AN][2] = -m_ro * betaz;
  rev[AN][3] = +m_ro * betay;
  rev[AN][4] = -signBX * rootrho * betaz;
  rev[AN][5] = +signBX * rootrho * betay;
  rev[AN][6] = -m_ro * (m_vt1 * betaz - m_vt2 * betay);
  rev[AP][0] = 0.0; rev[AP][1] = 0.0;
  rev[AP][2] = -rev[AN][2]; rev[AP][3] = -rev[AN][3]; rev[AP][4] = rev[AN][4]; rev[AP][5] = rev[AN][5];
  rev[AP][6] = -rev[AN][6];
  double das = m_ro * alphas, daf = m_ro * alphaf;
  double hb = m_H - Roe_B * Roe_B * im_ro;
  double vb = (m_vt1 * betay + m_vt2 * betaz);
  double inorm = fast_rcp(m_ro * Roe_a * Roe_a);
  rev[SN][0] = das;
  rev[SN][1] = das * (m_vn - Roe_cs);
  rev[SN][2] = das * m_vt1 - daf * Roe_cf * betay * signBX;
  rev[SN][3] = das * m_vt2 - daf * Roe_cf * betaz * signBX;
  rev[SN][4] = -rootrho * alphaf * Roe_a * betay;
  rev[SN][5] = -rootrho * alphaf * Roe_a * betaz;
  rev[SN][6] = das * (hb - m_vn * Roe_cs) - daf * Roe_cf * signBX * vb - rootrho * alphaf * Roe_a * Roe_Bt;
  rev[SP][0] = das;
  rev[SP][1] = das * (m_vn + Roe_cs);
  rev[SP][2] = das * m_vt1 + daf * Roe_cf * betay * signBX;
  rev[SP][3] = das * m_vt2 + daf * Roe_cf * betaz * signBX;
  rev[SP][4] = rev[SN][4];
  rev[SP][5] = rev[SN][5];
  rev[SP][6] = das * (hb + m_vn * Roe_cs) + daf * Roe_cf * signBX * vb - rootrho * alphaf * Roe_a * Roe_Bt;
  rev[FN][0] = daf;
  rev[FN][1] = daf * (m_vn - Roe_cf);
  rev[FN][2] = daf * m_vt1 + das * Roe_cs * betay * signBX;
  rev[FN][3] = daf * m_vt2 + das * Roe_cs * betaz * signBX;
  rev[FN][4] = rootrho * alphas * Roe_a * betay;
  rev[FN][5] = rootrho * alphas * Roe_a * betaz;
  rev[FN][6] = daf * (hb - m_vn * Roe_cf) + das * Roe_cs * signBX * vb + rootrho * alphas * Roe_a * Roe_Bt;
  rev[FP][0] = daf;
  rev[FP][1] = daf * (m_vn + Roe_cf);
  rev[FP][2] = daf * m_vt1 - das * Roe_cs * betay * signBX;
  rev[FP][3] = daf * m_vt2 - das * Roe_cs * betaz * signBX;
  rev[FP][4] = rev[FN][4];
  rev[FP][5] = rev[FN][5];
  rev[FP][6] = daf * (hb + m_vn * Roe_cf) - das * Roe_cs * signBX * vb + rootrho * alphas * Roe_a * Roe_Bt;
#pragma unroll
  for (int v = 0; v < 7; v++) {
    rev[SN][v] *= inorm; rev[SP][v] *= inorm; rev[FN][v] *= inorm; rev[FP][v] *= inorm;
  }
  Cons FL, FR;
  PUtoFlux<EQ_MHD>(L, UL, FL);
  PUtoFlux<EQ_MHD>(R, UR, FR);
  double f[7] = {FL.rho + FR.rho, FL.mn + FR.mn, FL.mt1 + FR.mt1, FL.mt2 + FR.mt2, FL.bbt1 + FR.bbt1, FL.bbt2 + FR.bbt2,
                 FL.erg + FR.erg};
#pragma unroll
  for (int iw = 0; iw < 7; iw++) {
    double w = str[iw] * fabs(ev[iw]);
#pragma unroll
    for (int c = 0; c < 7; c++) f[c] -= w * rev[iw][c];
  }
  flux.rho = 0.5 * f[0]; flux.mn = 0.5 * f[1]; flux.mt1 = 0.5 * f[2]; flux.mt2 = 0.5 * f[3];
  flux.bbt1 = 0.5 * f[4]; flux.bbt2 = 0.5 * f[5]; flux.erg = 0.5 * f[6];
  flux.bbn = 0.5 * (FL.bbn + FR.bbn);
  flux.psi = 0.0;
  pstar.ro = m_ro; pstar.vn = m_vn; pstar.vt1 = m_vt1; pstar.vt2 = m_vt2;
  pstar.bn = m_bn; pstar.bt1 = m_bt1; pstar.bt2 = m_bt2; pstar.psi = 0.0;
  pstar.pg = pdiv(m_ro * Roe_a * Roe_a, g);
}

// ---------------------------------------------------------------------------
// InterCellFlux: inviscid flux + FKJ98 viscosity, for one interface.
// FV_solver_base::InterCellFlux (solver_eqn_base.cpp:152-204) minus the tracer
// flux (done by the caller, which owns the tracer edge states).
//   use_hll  : HLLD->HLL switch already evaluated from divV / |grad p|/p
//   hc_etamax: H-correction eta (0 when AV != 3,4)
// ---------------------------------------------------------------------------
template <int EQ, int SOLVER, int AV>
__device__ __forceinline__ void intercell_flux(const Prim& eL, const Prim& eR, const PhysParams& pp, bool use_hll,
                                               double hc_etamax, Cons& flux, int ax = 0, int* rs_fail = nullptr) {
  constexpr bool FKJ = (AV == AV_FKJ98 || AV == AV_HCORR_FKJ98);
  Prim pstar;
  if (EQ == EQ_EULER) {
    if (SOLVER == SOLVE_LF) {
      lax_friedrichs<EQ_EULER>(eL, eR, pp, flux, pstar);
    } else if (SOLVER == SOLVE_RSLINEAR || SOLVER == SOLVE_RSEXACT || SOLVER == SOLVE_RSHYBRID) {
      const int f = hydro_JMs<SOLVER>(eL, eR, pp, ax, flux, pstar);
      if (rs_fail) *rs_fail |= f;
    } else if (SOLVER == SOLVE_ROE) {
      hydro_RoeCV(eL, eR, pp, hc_etamax, flux, pstar);
    } else if (SOLVER == SOLVE_FVS) {
      hydro_FVS<FKJ>(eL, eR, pp, flux, pstar);
    } else if (SOLVER == SOLVE_ROE_PV) {
      hydro_RoePV(eL, eR, pp, flux, pstar);
    } else {
      Cons ustar;
      hydro_HLL(eL, eR, pp, flux, ustar);
      if (FKJ) UtoP<EQ_EULER>(ustar, pstar, pp);
    }
    if (FKJ) {
      // FV_solver_Hydro_Euler::AVFalle (solver_eqn_hydro_adi.cpp:283-333)
      double prefactor = chydro(pstar.ro, pstar.pg, pp.gamma) * pp.etav * pstar.ro;
      double momvisc = prefactor * (eR.vn - eL.vn);
      double ergvisc = momvisc * pstar.vn;
      flux.mn -= momvisc;
      momvisc = prefactor * (eR.vt1 - eL.vt1);
      flux.mt1 -= momvisc;
      ergvisc += momvisc * pstar.vt1;
      momvisc = prefactor * (eR.vt2 - eL.vt2);
      flux.mt2 -= momvisc;
      ergvisc += momvisc * pstar.vt2;
      flux.erg -= ergvisc;
    }
  } else {
    // FV_solver_mhd_ideal_adi::AVFalle (solver_eqn_mhd_adi.cpp:209-288) works on the ORIGINAL edge states
    // (InterCellFlux passes lp, rp) and on pstar.  Everything it needs from the edge states is formed HERE,
    // before the Riemann solver -- the fast speed of the mean state (one scalar) and the five jumps -- so that
    // the edge states themselves are dead once the solver has consumed them (ptxas, capped at 128 registers,
    // spilled 150 bytes per thread to keep them alive across it).
    double fkj_c = 0.0, dvn = 0.0, dvt1 = 0.0, dvt2 = 0.0, dbt1 = 0.0, dbt2 = 0.0;
    if (FKJ) {
#ifdef PION_STRICT
      fkj_c = cfast_components(0.5 * (eL.ro + eR.ro), 0.5 * (eL.pg + eR.pg), 0.5 * (eL.bn + eR.bn),
                               0.5 * (eL.bt1 + eR.bt1), 0.5 * (eL.bt2 + eR.bt2), pp.gamma) * pp.etav;
#else
      // fast speed of the MEAN state from the SUMS: every term of cfast is a ratio to rho, so the halves cancel
      // against 1/(2 rho_mean) -- powers of two throughout, bit-identical to the literal form, five DMUL fewer
      const double irs = fast_rcp(eL.ro + eR.ro);  // = 1 / (2 rho_mean)
      fkj_c = fast_sqrt_pos(cfast2_ir(irs, eL.pg + eR.pg, eL.bn + eR.bn, eL.bt1 + eR.bt1, eL.bt2 + eR.bt2, pp.gamma, 0.5 * irs)) * pp.etav;
#endif
      dvn = eR.vn - eL.vn; dvt1 = eR.vt1 - eL.vt1; dvt2 = eR.vt2 - eL.vt2;
      dbt1 = eR.bt1 - eL.bt1; dbt2 = eR.bt2 - eL.bt2;
    }
    // GLM: Dedner 2x2 star state, Bx := Bx*, psi := 0 in the states handed to
    // the ideal-MHD solver (solver_eqn_mhd_adi.cpp:726-742)
    Prim l = eL, r = eR;
    double psistar = 0.0, bxstar = 0.0;
    if (EQ == EQ_GLM) {
      psistar = 0.5 * (eL.psi + eR.psi - (eR.bn - eL.bn));
      bxstar = 0.5 * (eL.bn + eR.bn - (eR.psi - eL.psi));
      l.psi = r.psi = 0.0;
      l.bn = r.bn = bxstar;
    }
    if (SOLVER == SOLVE_LF) {
      lax_friedrichs<EQ_MHD>(l, r, pp, flux, pstar);
    } else if (SOLVER == SOLVE_RSLINEAR) {
      const int f = mhd_JMs_linear(l, r, pp, flux, pstar);
      if (rs_fail) *rs_fail |= f;
    } else if (SOLVER == SOLVE_ROE) {
      mhd_RoeCV(l, r, pp, hc_etamax, flux, pstar);
    } else if (SOLVER == SOLVE_HLLD && !use_hll) {
#ifdef PION_NO_SAME_BN
      mhd_HLLD<FKJ, false>(l, r, pp, flux, pstar);
#else
      mhd_HLLD<FKJ, EQ == EQ_GLM>(l, r, pp, flux, pstar);
#endif
    } else {
      Cons ustar;
      mhd_HLL<FKJ>(l, r, pp, flux, ustar);
      // virtual UtoP: GLM version with psi==0 equals the ideal one
      if (FKJ) UtoP<EQ_MHD>(ustar, pstar, pp);
    }
    if (EQ == EQ_GLM) {
      flux.erg += pp.chyp * bxstar * psistar;
      flux.bbn = pp.chyp * psistar;
      flux.psi = pp.chyp * bxstar;
    }
    if (FKJ) {
      double prefactor = fkj_c * pstar.ro;
      double momvisc = prefactor * dvn;
      double ergvisc = momvisc * pstar.vn;
      flux.mn -= momvisc;
      momvisc = prefactor * dvt1;
      flux.mt1 -= momvisc;
      ergvisc += momvisc * pstar.vt1;
      momvisc = prefactor * dvt2;
      flux.mt2 -= momvisc;
      ergvisc += momvisc * pstar.vt2;
#ifdef PION_STRICT
      prefactor *= pdiv(pp.etav, pp.etav * pstar.ro);
#else
      // prefactor * etaB / (etav rho*) with etaB == etav (both are avcoeff, solver_eqn_base.cpp:82) is
      // cf * etav * rho* / rho*: the reciprocal cancels (2 ulp regrouping)
      prefactor = fkj_c;
#endif
      momvisc = prefactor * dbt1;
      flux.bbt1 -= momvisc;
      ergvisc += momvisc * pstar.bt1;
      momvisc = prefactor * dbt2;
      flux.bbt2 -= momvisc;
      ergvisc += momvisc * pstar.bt2;
      flux.erg -= ergvisc;
    }
  }
}

// maxspeed() used by the H-correction (solver_eqn_base.cpp:579-599): chydro for
// Euler, cfast along the sweep axis for MHD.
template <int EQ>
__device__ __forceinline__ double maxspeed(const Prim& p, double g) {
  if (EQ == EQ_EULER) return chydro(p.ro, p.pg, g);
  return cfast_components(p.ro, p.pg, p.bn, p.bt1, p.bt2, g);
}

}  // namespace pion
