// tests/host_physics/host_flux.cpp -- TEST INFRASTRUCTURE: one C entry point around the product's intercell_flux<EQ,SOLVER,AV>
// (the transformed copy of physics.cuh the test writes next to this file).
#include "physics.cuh"
using namespace pion;

template <int EQ, int SOLVER, int AV>
static int call(const Prim& l, const Prim& r, const PhysParams& pp, bool use_hll, double hc, int ax, Cons& F) {
  int fail = 0;
  intercell_flux<EQ, SOLVER, AV>(l, r, pp, use_hll, hc, F, ax, &fail);
  return fail;
}
template <int EQ, int SOLVER>
static int call_av(int av, const Prim& l, const Prim& r, const PhysParams& pp, bool use_hll, double hc, int ax, Cons& F) {
  switch (av) {
    case 0: return call<EQ, SOLVER, AV_NONE>(l, r, pp, use_hll, hc, ax, F);
    case 1: return call<EQ, SOLVER, AV_FKJ98>(l, r, pp, use_hll, hc, ax, F);
    case 3: return call<EQ, SOLVER, AV_HCORR>(l, r, pp, use_hll, hc, ax, F);
    default: return call<EQ, SOLVER, AV_HCORR_FKJ98>(l, r, pp, use_hll, hc, ax, F);
  }
}
// L, R: solver-frame primitive states {ro, pg, vn, vt1, vt2, bn, bt1, bt2, psi}; par: {gamma, etav, chyp, refvec_ro,
// rs_refvec[0..4]}; flux out in the solver frame {rho, erg, mn, mt1, mt2, bbn, bbt1, bbt2, psi}.  Returns -1 for a
// combination the product does not instantiate, else the solver's failure flag.
extern "C" int host_intercell_flux(int eq, int solver, int av, const double* L, const double* R, const double* par, int use_hll,
                                   double hc_etamax, int ax, double* flux) {
  PhysParams pp{};
  pp.gamma = par[0]; pp.etav = par[1]; pp.chyp = par[2]; pp.refvec_ro = par[3];
  for (int v = 0; v < 5; v++) pp.rs_refvec[v] = par[4 + v];
  const Prim l{L[0], L[1], L[2], L[3], L[4], L[5], L[6], L[7], L[8]}, r{R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7], R[8]};
  Cons F{};
  int rc = -1;
#define COMBO(E, S) if (eq == E && solver == S) rc = call_av<E, S>(av, l, r, pp, use_hll != 0, hc_etamax, ax, F);
  COMBO(EQ_EULER, SOLVE_RSLINEAR) COMBO(EQ_EULER, SOLVE_RSEXACT) COMBO(EQ_EULER, SOLVE_RSHYBRID) COMBO(EQ_EULER, SOLVE_ROE)
  COMBO(EQ_EULER, SOLVE_ROE_PV) COMBO(EQ_EULER, SOLVE_FVS) COMBO(EQ_EULER, SOLVE_HLL)
  COMBO(EQ_MHD, SOLVE_RSLINEAR) COMBO(EQ_MHD, SOLVE_ROE) COMBO(EQ_MHD, SOLVE_HLLD) COMBO(EQ_MHD, SOLVE_HLL)
  COMBO(EQ_GLM, SOLVE_RSLINEAR) COMBO(EQ_GLM, SOLVE_ROE) COMBO(EQ_GLM, SOLVE_HLLD) COMBO(EQ_GLM, SOLVE_HLL)
#undef COMBO
  flux[0] = F.rho; flux[1] = F.erg; flux[2] = F.mn; flux[3] = F.mt1; flux[4] = F.mt2;
  flux[5] = F.bbn; flux[6] = F.bbt1; flux[7] = F.bbt2; flux[8] = F.psi;
  return rc;
}
