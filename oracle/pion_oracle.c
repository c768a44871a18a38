/* TEST INFRASTRUCTURE ONLY -- the parity oracle, never linked into the product.
 *
 * pion_oracle.c: a plain-C, single-threaded CPU restatement of the reference's
 * finite-volume hydro/MHD dynamics update on a uniform Cartesian grid.  It keeps
 * the reference's own traversal (array-of-structures cells, 1-D column sweeps
 * that scatter into dU, per-boundary ghost lists) so that results agree
 * BIT-FOR-BIT with the reference build (oracle/_ref) when compiled without FMA
 * contraction (-ffp-contract=off; the reference's default x86-64 build has no
 * FMA either).  Every function cites the reference file:line it follows
 * (paths relative to /root/reference/source).
 *
 * Pinning: tests/test_oracle_vs_ref.py compares this file against the compiled
 * reference on every configuration family; tests/golden/ holds vectors
 * generated from the compiled reference (tests/golden/make_golden.py).
 *
 * Not restated (out of scope): FVS / Roe-PV / linear / exact
 * / hybrid solvers, Lax-Friedrichs, nested grids, ray tracing, CVODE chemistry.
 */
#include "pion_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* state-vector indices: constants.h:256-281 */
enum { RO = 0, PG = 1, VX = 2, VY = 3, VZ = 4, BX = 5, BY = 6, BZ = 7, SI = 8 };
enum { RHO = 0, ERG = 1, MMX = 2, MMY = 3, MMZ = 4, BBX = 5, BBY = 6, BBZ = 7, PSI = 8 };
enum { XN = 0, XP = 1, YN = 2, YP = 3, ZN = 4, ZP = 5 };
#define OA1 1
#define OA2 2
/* constants.h:150-153,336-339 */
#define SMALLVALUE 1.0e-12
#define MACHINEACCURACY 5.e-16
#define TINYVALUE 1.0e-100
#define VERY_TINY_VALUE 1.0e-200
#define BASE_RHO 1.0e-5

typedef struct {
  int type;   /* PO_BC_* */
  int dir;    /* off-grid direction, -1 for internal */
  long n;
  long *cell; /* ghost cell index */
  long *npt;  /* source cell index (or -1) */
  int *isedge;
  double refval[PO_MAXVAR];
} bc_list;

struct pion_oracle {
  pion_oracle_config cfg;
  int nv, ndim, nbc, ntr, ftr;
  int NGa[3], nb[3];
  long ncell, stride[3];
  double dx;
  double *P, *Ph, *dU; /* [ncell][nv] */
  double *hcorr;       /* [ncell][3] */
  double *divv, *gradp;
  unsigned char *isgd, *isdomain, *tsflag, *iswind; /* isbd = !isgd || iswind */
  /* solver "class" state: eqns_base.cpp:94-131, solver_eqn_base.h */
  int dir, eVX, eVY, eVZ, eBX, eBY, eBZ, eMX, eMY, eMZ, eBBX, eBBY, eBBZ;
  double FV_dt, chyp, cr, HC_etamax, gamma;
  long nfail_riemann; /* JMs_riemann_solve returned an error (fatal in the reference) */
  /* time: sim_params.cpp:53 */
  double simtime, dt, last_dt, next_optime;
  int timestep;
  bc_list bcs[10];
  int nbcs;
  long neg_rho, neg_pg;
  /* mp_only_cooling constants (mp_only_cooling.cpp:96-110) */
  int have_mp;
  double Mu, Mu_tot_over_kB, Mu_elec, Mu_ion, inv_Mu2, inv_Mu2_elec_H, MinT, MaxT;
  double *tT, *t_rrhp, *t_Crrh, *t_Cffhe, *t_Cfbdn, *t_Ccie;
  double *s_rrhp, *s_Crrh, *s_Cffhe, *s_Cfbdn, *s_Ccie;
  int nT;
  /* cooling_function_SD93CIE spline (EP.cooling 4..7): knots, natural-spline coefficients c = y''/2 */
  int ns;
  double *sx, *sy, *sc, s_minslope, s_maxslope;
  double mp_rho, mp_gamma; /* integrator "members" */
  /* stellar wind (grid/stellar_wind_BC.cpp): cells of every source in add order */
  long wind_n;
  long *wind_cell;
  double *wind_p; /* [wind_n][nv] */
};

/* ------------------------------------------------------------------ */
/* grid helpers                                                        */
/* ------------------------------------------------------------------ */
static inline long cidx(const pion_oracle *s, int i, int j, int k) {
  return (long)i + (long)s->NGa[0] * ((long)j + (long)s->NGa[1] * (long)k);
}
static inline void cijk(const pion_oracle *s, long c, int *ijk) {
  ijk[0] = (int)(c % s->NGa[0]);
  ijk[1] = (int)((c / s->NGa[0]) % s->NGa[1]);
  ijk[2] = (int)(c / ((long)s->NGa[0] * s->NGa[1]));
}
/* NextPt(c,dir) = c->ngb[dir] (uniform_grid.cpp:441-811): -1 at the array edge */
static inline long nextpt(const pion_oracle *s, long c, int dir) {
  int a = dir / 2;
  int ijk[3];
  cijk(s, c, ijk);
  if (dir & 1) {
    if (ijk[a] + 1 >= s->NGa[a]) return -1;
    return c + s->stride[a];
  }
  if (ijk[a] - 1 < 0) return -1;
  return c - s->stride[a];
}
/* cell-centre position: cell_interface.cpp:506-512 with pos = 2*(i-nb)+1 */
static inline double dpos(const pion_oracle *s, long c, int a) {
  int ijk[3];
  cijk(s, c, ijk);
  int ipos = 2 * (ijk[a] - s->nb[a]) + 1;
  double dxo2 = 0.5 * s->dx;
  return s->cfg.xmin[a] + ipos * dxo2;
}

/* eqns_base::SetDirection (equations/eqns_base.cpp:94-131) */
static void set_direction(pion_oracle *s, int a) {
  s->dir = a;
  s->eVX = VX + a; s->eVY = VX + (a + 1) % 3; s->eVZ = VX + (a + 2) % 3;
  s->eBX = BX + a; s->eBY = BX + (a + 1) % 3; s->eBZ = BX + (a + 2) % 3;
  s->eMX = MMX + a; s->eMY = MMX + (a + 1) % 3; s->eMZ = MMX + (a + 2) % 3;
  s->eBBX = BBX + a; s->eBBY = BBX + (a + 1) % 3; s->eBBZ = BBX + (a + 2) % 3;
}

/* ------------------------------------------------------------------ */
/* microphysics: mp_only_cooling (only what the dynamics path touches) */
/* ------------------------------------------------------------------ */
/* mp_only_cooling::Temperature / Set_Temp (mp_only_cooling.cpp:244-283) */
static inline double mp_temperature(const pion_oracle *s, const double *p) {
  return p[PG] * s->Mu_tot_over_kB / p[RO];
}
static inline void mp_set_temp(const pion_oracle *s, double *p, double T) {
  p[PG] = p[RO] * T / s->Mu_tot_over_kB;
}
/* microphysics_base::sCMA (microphysics_base.cpp:80-126).  mp_only_cooling has
 * no element tracers (n_el==0), so only the ">1 -> 1/p" clamp survives: the
 * second assignment overrides the "<0 -> 0" one (:111-112). */
static void mp_sCMA(const pion_oracle *s, double *corr, const double *p) {
  for (int v = 0; v < s->nv; v++) corr[v] = 1.0;
  for (int t = 0; t < s->ntr; t++) {
    int v = s->ftr + t;
    corr[v] = (p[v] < 0.0) ? 0.0 : 1.0;
    corr[v] = (p[v] > 1.0) ? 1.0 / p[v] : 1.0;
  }
}

/* ------------------------------------------------------------------ */
/* equations                                                           */
/* ------------------------------------------------------------------ */
/* eqns_Euler::PtoU (eqns_hydro_adiabatic.cpp:89-108) + tracers
 * (solver_eqn_hydro_adi.cpp:213-222) */
static void euler_PtoU(const pion_oracle *s, const double *p, double *u) {
  for (int t = 0; t < s->ntr; t++) u[s->ftr + t] = p[s->ftr + t] * p[RO];
  u[RHO] = p[RO];
  u[s->eMX] = p[RO] * p[s->eVX];
  u[s->eMY] = p[RO] * p[s->eVY];
  u[s->eMZ] = p[RO] * p[s->eVZ];
  u[ERG] = p[RO] * (p[s->eVX] * p[s->eVX] + p[s->eVY] * p[s->eVY] + p[s->eVZ] * p[s->eVZ]) * 0.5 +
           p[PG] / (s->gamma - 1.);
}
/* eqns_Euler::UtoP (eqns_hydro_adiabatic.cpp:117-205) + tracers (:231-241) */
static int euler_UtoP(pion_oracle *s, const double *u, double *p) {
  int err = 0;
  double g = s->gamma;
  for (int t = 0; t < s->ntr; t++) p[s->ftr + t] = u[s->ftr + t] / u[RHO];
  p[RO] = u[RHO];
  p[s->eVX] = u[s->eMX] / u[RHO];
  p[s->eVY] = u[s->eMY] / u[RHO];
  p[s->eVZ] = u[s->eMZ] / u[RHO];
  p[PG] = (g - 1.0) * (u[ERG] - p[RO] * (p[s->eVX] * p[s->eVX] + p[s->eVY] * p[s->eVY] + p[s->eVZ] * p[s->eVZ]) / 2.0);
  if (p[RO] <= 0.0) { /* fatal in the reference (:153) */
    s->neg_rho++;
    p[RO] = BASE_RHO;
    p[s->eVX] = u[s->eMX] / p[RO];
    p[s->eVY] = u[s->eMY] / p[RO];
    p[s->eVZ] = u[s->eMZ] / p[RO];
    p[PG] = (g - 1.0) * (u[ERG] - p[RO] * (p[s->eVX] * p[s->eVX] + p[s->eVY] * p[s->eVY] + p[s->eVZ] * p[s->eVZ]) / 2.0);
    err += 1;
  }
  /* SET_NEGATIVE_PRESSURE_TO_FIXED_TEMPERATURE (functionality_flags.h) */
  if (p[PG] <= 0.0) {
    s->neg_pg++;
    if (s->have_mp) mp_set_temp(s, p, s->cfg.min_temperature);
    else p[PG] = 0.01 * p[RO];
  } else if (s->have_mp && mp_temperature(s, p) < s->cfg.min_temperature) {
    mp_set_temp(s, p, s->cfg.min_temperature);
  }
  return err;
}
/* eqns_Euler::PUtoFlux (:296-308) */
static void euler_PUtoFlux(const pion_oracle *s, const double *p, const double *u, double *f) {
  f[RHO] = u[s->eMX];
  f[s->eMX] = u[s->eMX] * p[s->eVX] + p[PG];
  f[s->eMY] = u[s->eMX] * p[s->eVY];
  f[s->eMZ] = u[s->eMX] * p[s->eVZ];
  f[ERG] = p[s->eVX] * (u[ERG] + p[PG]);
}
/* eqns_Euler::UtoFlux (:317-333) */
static void euler_UtoFlux(const pion_oracle *s, const double *u, double *f) {
  double pg = (s->gamma - 1.) *
              (u[ERG] - (u[s->eMX] * u[s->eMX] + u[s->eMY] * u[s->eMY] + u[s->eMZ] * u[s->eMZ]) * 0.5 / u[RHO]);
  f[RHO] = u[s->eMX];
  f[s->eMX] = u[s->eMX] * u[s->eMX] / u[RHO] + pg;
  f[s->eMY] = u[s->eMX] * u[s->eMY] / u[RHO];
  f[s->eMZ] = u[s->eMX] * u[s->eMZ] / u[RHO];
  f[ERG] = u[s->eMX] * (u[ERG] + pg) / u[RHO];
}
static inline double chydro(const pion_oracle *s, const double *p) { return sqrt(s->gamma * p[PG] / p[RO]); }

/* eqns_mhd_ideal::PtoU (eqns_mhd_adiabatic.cpp:79-101) */
static void mhd_ideal_PtoU(const pion_oracle *s, const double *p, double *u) {
  u[RHO] = p[RO];
  u[s->eMX] = p[RO] * p[s->eVX];
  u[s->eMY] = p[RO] * p[s->eVY];
  u[s->eMZ] = p[RO] * p[s->eVZ];
  u[s->eBBX] = p[s->eBX];
  u[s->eBBY] = p[s->eBY];
  u[s->eBBZ] = p[s->eBZ];
  u[ERG] = (p[RO] * (p[s->eVX] * p[s->eVX] + p[s->eVY] * p[s->eVY] + p[s->eVZ] * p[s->eVZ]) * 0.5) +
           (p[PG] / (s->gamma - 1.)) +
           ((u[s->eBBX] * u[s->eBBX] + u[s->eBBY] * u[s->eBBY] + u[s->eBBZ] * u[s->eBBZ]) * 0.5);
}
/* eqns_mhd_mixedGLM::PtoU (:598-609) */
static void glm_PtoU(const pion_oracle *s, const double *p, double *u) {
  u[PSI] = p[SI];
  mhd_ideal_PtoU(s, p, u);
  u[ERG] += 0.5 * u[PSI] * u[PSI];
}
/* the virtual PtoU of the FV solver classes, with tracers
 * (solver_eqn_mhd_adi.cpp:296-305, :855-872) */
static void mhd_PtoU(const pion_oracle *s, const double *p, double *u) {
  if (s->cfg.eqntype == PO_EQGLM) glm_PtoU(s, p, u);
  else mhd_ideal_PtoU(s, p, u);
  for (int t = 0; t < s->ntr; t++) u[s->ftr + t] = p[s->ftr + t] * p[RO];
}
/* eqns_mhd_ideal::check_pressure (:137-224) */
static int mhd_check_pressure(pion_oracle *s, const double *u, double *p) {
  int err = 0;
  double g = s->gamma;
  if (p[RO] <= 0.0) { /* fatal in the reference (:158) */
    s->neg_rho++;
    p[RO] = BASE_RHO * s->cfg.refvec[RO];
    p[s->eVX] *= u[RHO] / p[RO];
    p[s->eVY] *= u[RHO] / p[RO];
    p[s->eVZ] *= u[RHO] / p[RO];
    p[PG] = (g - 1) * (u[ERG] - p[RO] * (p[s->eVX] * p[s->eVX] + p[s->eVY] * p[s->eVY] + p[s->eVZ] * p[s->eVZ]) / 2. -
                       (u[s->eBBX] * u[s->eBBX] + u[s->eBBY] * u[s->eBBY] + u[s->eBBZ] * u[s->eBBZ]) / 2.);
    err += 1;
  }
  if (p[PG] <= 0.0) {
    s->neg_pg++;
    if (s->have_mp) mp_set_temp(s, p, s->cfg.min_temperature);
    else p[PG] = 0.01 * p[RO];
  } else if (s->have_mp && mp_temperature(s, p) < s->cfg.min_temperature) {
    mp_set_temp(s, p, s->cfg.min_temperature);
  }
  return err;
}
/* eqns_mhd_ideal::UtoP (:110-129), eqns_mhd_mixedGLM::UtoP (:618-641), with
 * tracers (solver_eqn_mhd_adi.cpp:314-324, :881-897) */
static int mhd_UtoP(pion_oracle *s, const double *u, double *p) {
  double g = s->gamma;
  for (int t = 0; t < s->ntr; t++) p[s->ftr + t] = u[s->ftr + t] / u[RHO];
  if (s->cfg.eqntype == PO_EQGLM) {
    p[SI] = u[PSI];
    p[RO] = u[RHO];
    p[s->eVX] = u[s->eMX] / u[RHO];
    p[s->eVY] = u[s->eMY] / u[RHO];
    p[s->eVZ] = u[s->eMZ] / u[RHO];
    p[PG] = (g - 1.0) * (u[ERG] - p[RO] * (p[s->eVX] * p[s->eVX] + p[s->eVY] * p[s->eVY] + p[s->eVZ] * p[s->eVZ]) * 0.5 -
                         0.5 * u[PSI] * u[PSI] -
                         (u[s->eBBX] * u[s->eBBX] + u[s->eBBY] * u[s->eBBY] + u[s->eBBZ] * u[s->eBBZ]) * 0.5);
  } else {
    p[RO] = u[RHO];
    p[s->eVX] = u[s->eMX] / u[RHO];
    p[s->eVY] = u[s->eMY] / u[RHO];
    p[s->eVZ] = u[s->eMZ] / u[RHO];
    p[PG] = (g - 1) * (u[ERG] - p[RO] * (p[s->eVX] * p[s->eVX] + p[s->eVY] * p[s->eVY] + p[s->eVZ] * p[s->eVZ]) / 2. -
                       (u[s->eBBX] * u[s->eBBX] + u[s->eBBY] * u[s->eBBY] + u[s->eBBZ] * u[s->eBBZ]) / 2.);
  }
  p[s->eBX] = u[s->eBBX];
  p[s->eBY] = u[s->eBBY];
  p[s->eBZ] = u[s->eBBZ];
  return mhd_check_pressure(s, u, p);
}
/* eqns_mhd_ideal::PUtoFlux (:307-328) */
static void mhd_PUtoFlux(const pion_oracle *s, const double *p, const double *u, double *f) {
  double pm = (u[s->eBBX] * u[s->eBBX] + u[s->eBBY] * u[s->eBBY] + u[s->eBBZ] * u[s->eBBZ]) / 2.;
  f[RHO] = u[s->eMX];
  f[s->eMX] = u[s->eMX] * p[s->eVX] + p[PG] + pm - u[s->eBBX] * u[s->eBBX];
  f[s->eMY] = u[s->eMX] * p[s->eVY] - u[s->eBBX] * u[s->eBBY];
  f[s->eMZ] = u[s->eMX] * p[s->eVZ] - u[s->eBBX] * u[s->eBBZ];
  f[ERG] = p[s->eVX] * (u[ERG] + p[PG] + pm) -
           u[s->eBBX] * (p[s->eVX] * u[s->eBBX] + p[s->eVY] * u[s->eBBY] + p[s->eVZ] * u[s->eBBZ]);
  f[s->eBBX] = 0.;
  f[s->eBBY] = p[s->eVX] * p[s->eBY] - p[s->eVY] * p[s->eBX];
  f[s->eBBZ] = p[s->eVX] * p[s->eBZ] - p[s->eVZ] * p[s->eBX];
}
/* eqns_mhd_ideal::cfast_components (:263-276) */
static double cfast_components(double ro, double pg, double bx, double by, double bz, double g) {
  double ch = sqrt(g * pg / ro);
  double temp1 = ch * ch + (bx * bx + by * by + bz * bz) / ro;
  double temp2 = 4. * ch * ch * bx * bx / ro;
  temp2 = fmax(MACHINEACCURACY, temp1 * temp1 - temp2);
  return sqrt((temp1 + sqrt(temp2)) / 2.);
}
/* eqns_mhd_ideal::cfast (:247-258) in the current direction */
static double cfast(const pion_oracle *s, const double *p) {
  double ch = sqrt(s->gamma * p[PG] / p[RO]);
  double temp1 = ch * ch + (p[s->eBX] * p[s->eBX] + p[s->eBY] * p[s->eBY] + p[s->eBZ] * p[s->eBZ]) / p[RO];
  double temp2 = 4. * ch * ch * p[s->eBX] * p[s->eBX] / p[RO];
  temp2 = fmax(MACHINEACCURACY, temp1 * temp1 - temp2);
  return sqrt((temp1 + sqrt(temp2)) / 2.);
}
static inline double mhd_Ptot(const pion_oracle *s, const double *p) { /* :474-480 */
  return p[PG] + 0.5 * (p[s->eBX] * p[s->eBX] + p[s->eBY] * p[s->eBY] + p[s->eBZ] * p[s->eBZ]);
}
/* maxspeed(): chydro for Euler (eqns_hydro_adiabatic.h:94-97), cfast for MHD
 * (eqns_mhd_adiabatic.h:120-123) */
static inline double maxspeed(const pion_oracle *s, const double *p) {
  return (s->cfg.eqntype == PO_EQEUL) ? chydro(s, p) : cfast(s, p);
}

/* constants::equalD (constants.cpp:48-69) */
static int equalD(double a, double b) {
  if (a == b) return 1;
  if (fabs(a) + fabs(b) < TINYVALUE) return 1;
  if ((fabs(a - b) / (fabs(a) + fabs(b) + TINYVALUE)) < SMALLVALUE) return 1;
  return 0;
}

/* ------------------------------------------------------------------ */
/* Riemann solvers                                                     */
/* ------------------------------------------------------------------ */
/* HLL_hydro::hydro_HLL_flux_solver + HLL_signal_speeds (HLL_hydro.cpp:92-170).
 * Tracer entries of flux/ustar are left 0: the reference fills them from a
 * stale HD_FL[eqRHO] (solver_eqn_hydro_adi.cpp:250-259 calls the tracer line
 * BEFORE eqns_Euler::PUtoFlux) but they are overwritten by
 * set_interface_tracer_flux and unused in pstar. */
static void hydro_HLL(pion_oracle *s, const double *Pl, const double *Pr, double *flux, double *ustar) {
  double UL[PO_MAXVAR] = {0}, UR[PO_MAXVAR] = {0}, FL[PO_MAXVAR] = {0}, FR[PO_MAXVAR] = {0};
  euler_PtoU(s, Pl, UL);
  euler_PtoU(s, Pr, UR);
  euler_PUtoFlux(s, Pl, UL, FL);
  euler_PUtoFlux(s, Pr, UR, FR);
  double cf_l = chydro(s, Pl), cf_r = chydro(s, Pr);
  double cf_max = fmax(cf_l, cf_r);
  double Sl = fmin(Pl[s->eVX], Pr[s->eVX]) - cf_max;
  double Sr = fmax(Pl[s->eVX], Pr[s->eVX]) + cf_max;
  for (int v = 0; v < 5; v++) {
    if (Sl > 0) flux[v] = FL[v];
    else if (Sr < 0) flux[v] = FR[v];
    else flux[v] = (Sr * FL[v] - Sl * FR[v] + Sr * Sl * (UR[v] - UL[v])) / (Sr - Sl);
  }
  for (int v = 0; v < 5; v++) ustar[v] = (Sr * UR[v] - Sl * UL[v] + FL[v] - FR[v]) / (Sr - Sl);
  for (int t = 0; t < s->ntr; t++) { /* finite placeholder, see header comment */
    int v = s->ftr + t;
    flux[v] = 0.0;
    ustar[v] = (Sr * UR[v] - Sl * UL[v]) / (Sr - Sl);
  }
}

/* Riemann_Roe_Hydro_CV::Roe_flux_solver_symmetric
 * (Roe_Hydro_ConservedVar_solver.cpp:129-175 and the helpers :215-470) */
static void hydro_RoeCV(pion_oracle *s, const double *left, const double *right, double hc_eta, double *pstar,
                        double *flux) {
  const double g = s->gamma;
  const int eHH = PG;
  double meanp[5], eval[5], strength[5], udiff[5], ul[5], ur[5], evec[5][5], tmp[5];
  /* set_Roe_mean_state (:215-252); Enthalpy eqns_hydro_adiabatic.cpp:356-364 */
  double rl = sqrt(left[RO]), rr = sqrt(right[RO]);
  double lH = 0.5 * (left[s->eVX] * left[s->eVX] + left[s->eVY] * left[s->eVY] + left[s->eVZ] * left[s->eVZ]) +
              g * left[PG] / (g - 1.0) / left[RO];
  double rH = 0.5 * (right[s->eVX] * right[s->eVX] + right[s->eVY] * right[s->eVY] + right[s->eVZ] * right[s->eVZ]) +
              g * right[PG] / (g - 1.0) / right[RO];
  double denom = 1.0 / (rl + rr);
  meanp[RO] = rl * rr;
  meanp[s->eVX] = (rl * left[s->eVX] + rr * right[s->eVX]) * denom;
  meanp[s->eVY] = (rl * left[s->eVY] + rr * right[s->eVY]) * denom;
  meanp[s->eVZ] = (rl * left[s->eVZ] + rr * right[s->eVZ]) * denom;
  meanp[eHH] = (rl * lH + rr * rH) * denom;
  double v2 = meanp[s->eVX] * meanp[s->eVX] + meanp[s->eVY] * meanp[s->eVY] + meanp[s->eVZ] * meanp[s->eVZ];
  double a = sqrt((g - 1.0) * fmax(meanp[eHH] - 0.5 * v2, 1.0e-12 * v2));
  /* set_eigenvalues (:258-283) incl. H-correction (:369-380 in the survey's numbering) */
  eval[0] = meanp[s->eVX] - a;
  eval[1] = eval[2] = eval[3] = meanp[s->eVX];
  eval[4] = meanp[s->eVX] + a;
  for (int v = 0; v < 5; v++) {
    if (eval[v] < 0.0) eval[v] = fmin(eval[v], -hc_eta);
    else eval[v] = fmax(eval[v], hc_eta);
  }
  /* set_eigenvectors (:289-335) */
  evec[0][RHO] = 1.0; evec[0][s->eMX] = meanp[s->eVX] - a; evec[0][s->eMY] = meanp[s->eVY];
  evec[0][s->eMZ] = meanp[s->eVZ]; evec[0][ERG] = meanp[eHH] - meanp[s->eVX] * a;
  evec[1][RHO] = 1.0; evec[1][s->eMX] = meanp[s->eVX]; evec[1][s->eMY] = meanp[s->eVY];
  evec[1][s->eMZ] = meanp[s->eVZ]; evec[1][ERG] = 0.5 * v2;
  evec[2][RHO] = 0.0; evec[2][s->eMX] = 0.0; evec[2][s->eMY] = 1.0; evec[2][s->eMZ] = 0.0;
  evec[2][ERG] = meanp[s->eVY];
  evec[3][RHO] = 0.0; evec[3][s->eMX] = 0.0; evec[3][s->eMY] = 0.0; evec[3][s->eMZ] = 1.0;
  evec[3][ERG] = meanp[s->eVZ];
  evec[4][RHO] = 1.0; evec[4][s->eMX] = meanp[s->eVX] + a; evec[4][s->eMY] = meanp[s->eVY];
  evec[4][s->eMZ] = meanp[s->eVZ]; evec[4][ERG] = meanp[eHH] + meanp[s->eVX] * a;
  /* set_ul_ur_udiff (:341-360): eqns_Euler::PtoU without tracers */
  {
    double UL[PO_MAXVAR], UR[PO_MAXVAR];
    int ntr = s->ntr;
    s->ntr = 0;
    euler_PtoU(s, left, UL);
    euler_PtoU(s, right, UR);
    s->ntr = ntr;
    for (int v = 0; v < 5; v++) { ul[v] = UL[v]; ur[v] = UR[v]; }
  }
  for (int v = 0; v < 5; v++) udiff[v] = equalD(ur[v], ul[v]) ? 0.0 : ur[v] - ul[v];
  /* set_wave_strengths (:366-385) */
  strength[2] = udiff[s->eMY] - meanp[s->eVY] * udiff[RO];
  strength[3] = udiff[s->eMZ] - meanp[s->eVZ] * udiff[RO];
  double u5bar = udiff[ERG] - strength[2] * meanp[s->eVY] - strength[3] * meanp[s->eVZ];
  strength[1] = (udiff[RHO] * (meanp[eHH] - meanp[s->eVX] * meanp[s->eVX]) + meanp[s->eVX] * udiff[s->eMX] - u5bar) *
                (g - 1.0) / a / a;
  strength[0] = 0.5 * (udiff[RHO] * (meanp[s->eVX] + a) - udiff[s->eMX] - a * strength[1]) / a;
  strength[4] = udiff[RHO] - strength[0] - strength[1];
  /* calculate_symmetric_flux (:391-417) */
  euler_UtoFlux(s, ul, flux);
  euler_UtoFlux(s, ur, tmp);
  for (int v = 0; v < 5; v++) flux[v] += tmp[v];
  for (int iw = 0; iw < 5; iw++) {
    flux[RHO] -= strength[iw] * fabs(eval[iw]) * evec[iw][RHO];
    flux[s->eMX] -= strength[iw] * fabs(eval[iw]) * evec[iw][s->eMX];
    flux[s->eMY] -= strength[iw] * fabs(eval[iw]) * evec[iw][s->eMY];
    flux[s->eMZ] -= strength[iw] * fabs(eval[iw]) * evec[iw][s->eMZ];
    flux[ERG] -= strength[iw] * fabs(eval[iw]) * evec[iw][ERG];
  }
  for (int v = 0; v < 5; v++) flux[v] *= 0.5;
  /* set_pstar_from_meanp (:423-436) */
  for (int v = 0; v < 5; v++) pstar[v] = meanp[v];
  pstar[PG] = meanp[RO] * a * a / g;
}

/* HLLD_MHD::HLLD_signal_speeds (HLLD_MHD.cpp:342-368) */
static void hlld_speeds(const pion_oracle *s, const double *Pl, const double *Pr, double *Sl, double *Sr) {
  double Bx = 0.5 * (Pl[s->eBX] + Pr[s->eBX]);
  double cf_l = cfast_components(Pl[RO], Pl[PG], Bx, Pl[s->eBY], Pl[s->eBZ], s->gamma);
  double cf_r = cfast_components(Pr[RO], Pr[PG], Bx, Pr[s->eBY], Pr[s->eBZ], s->gamma);
  double cf_max = fmax(cf_l, cf_r);
  *Sl = fmin(Pl[s->eVX], Pr[s->eVX]) - cf_max;
  *Sr = fmax(Pl[s->eVX], Pr[s->eVX]) + cf_max;
}
/* HLLD_MHD::MHD_HLL_flux_solver (HLLD_MHD.cpp:377-417) */
static void mhd_HLL(pion_oracle *s, const double *Pl, const double *Pr, double *flux, double *ustar) {
  double UL[PO_MAXVAR], UR[PO_MAXVAR], FL[PO_MAXVAR], FR[PO_MAXVAR], l0, l1;
  mhd_ideal_PtoU(s, Pl, UL);
  mhd_ideal_PtoU(s, Pr, UR);
  mhd_PUtoFlux(s, Pl, UL, FL);
  mhd_PUtoFlux(s, Pr, UR, FR);
  hlld_speeds(s, Pl, Pr, &l0, &l1);
  if (l0 > 0.0) {
    for (int v = 0; v < 8; v++) { flux[v] = FL[v]; ustar[v] = UL[v]; }
  } else if (l1 < 0.0) {
    for (int v = 0; v < 8; v++) { flux[v] = FR[v]; ustar[v] = UR[v]; }
  } else {
    for (int v = 0; v < 8; v++) flux[v] = (l1 * FL[v] - l0 * FR[v] + l1 * l0 * (UR[v] - UL[v])) / (l1 - l0);
    for (int v = 0; v < 8; v++) ustar[v] = (l1 * UR[v] - l0 * UL[v] - FR[v] + FL[v]) / (l1 - l0);
  }
}
/* HLLD_MHD::MHD_HLLD_flux_solver (HLLD_MHD.cpp:124-333) */
static void mhd_HLLD(pion_oracle *s, const double *Pl, const double *Pr, double *flux, double *ustar) {
  const int eRHO = RHO, eMX = s->eMX, eMY = s->eMY, eMZ = s->eMZ, eBBX = s->eBBX, eBBY = s->eBBY, eBBZ = s->eBBZ;
  const int eVX = s->eVX, eVY = s->eVY, eVZ = s->eVZ, eBY = s->eBY, eBZ = s->eBZ;
  double UL[PO_MAXVAR], UR[PO_MAXVAR], FL[PO_MAXVAR], FR[PO_MAXVAR];
  double ULs[8], URs[8], ULss[8], URss[8], lam[5];
  double Bx = 0.5 * (Pl[s->eBX] + Pr[s->eBX]);
  mhd_ideal_PtoU(s, Pl, UL);
  mhd_ideal_PtoU(s, Pr, UR);
  mhd_PUtoFlux(s, Pl, UL, FL);
  mhd_PUtoFlux(s, Pr, UR, FR);
  hlld_speeds(s, Pl, Pr, &lam[0], &lam[4]);

  double sl_vl = lam[0] - Pl[eVX];
  double sr_vr = lam[4] - Pr[eVX];
  double tp_r = mhd_Ptot(s, Pr);
  double tp_l = mhd_Ptot(s, Pl);
  double temp = sr_vr * Pr[RO] - sl_vl * Pl[RO];
  lam[2] = (sr_vr * UR[eMX] - sl_vl * UL[eMX] - tp_r + tp_l) / temp;
  double tp_s = (sr_vr * Pr[RO] * tp_l - sl_vl * Pl[RO] * tp_r + Pl[RO] * Pr[RO] * sr_vr * sl_vl * (Pr[eVX] - Pl[eVX])) / temp;
  double sl_sm = lam[0] - lam[2];
  double sr_sm = lam[4] - lam[2];
  ULs[eRHO] = Pl[RO] * sl_vl / sl_sm;
  URs[eRHO] = Pr[RO] * sr_vr / sr_sm;
  ULs[eMX] = lam[2] * ULs[eRHO];
  URs[eMX] = lam[2] * URs[eRHO];
  double temp_l1 = lam[2] - Pl[eVX];
  double temp_l2 = Pl[RO] * sl_vl * sl_sm - Bx * Bx;
  double temp_r1 = lam[2] - Pr[eVX];
  double temp_r2 = Pr[RO] * sr_vr * sr_sm - Bx * Bx;
  double vys_l = Pl[eVY], vys_r = Pr[eVY], vzs_l = Pl[eVZ], vzs_r = Pr[eVZ];
  if (isfinite(temp_l1 / temp_l2)) {
    vys_l = Pl[eVY] - Bx * Pl[eBY] * temp_l1 / temp_l2;
    vzs_l = Pl[eVZ] - Bx * Pl[eBZ] * temp_l1 / temp_l2;
  }
  if (isfinite(temp_r1 / temp_r2)) {
    vys_r = Pr[eVY] - Bx * Pr[eBY] * temp_r1 / temp_r2;
    vzs_r = Pr[eVZ] - Bx * Pr[eBZ] * temp_r1 / temp_r2;
  }
  ULs[eMY] = vys_l * ULs[eRHO];
  URs[eMY] = vys_r * URs[eRHO];
  ULs[eMZ] = vzs_l * ULs[eRHO];
  URs[eMZ] = vzs_r * URs[eRHO];
  ULs[eBBX] = URs[eBBX] = Bx;
  temp_l1 = Pl[RO] * sl_vl * sl_vl - Bx * Bx;
  temp_r1 = Pr[RO] * sr_vr * sr_vr - Bx * Bx;
  ULs[eBBY] = 0.0; URs[eBBY] = 0.0; ULs[eBBZ] = 0.0; URs[eBBZ] = 0.0;
  if (isfinite(temp_l1 / temp_l2)) {
    ULs[eBBY] = Pl[eBY] * temp_l1 / temp_l2;
    ULs[eBBZ] = Pl[eBZ] * temp_l1 / temp_l2;
  }
  if (isfinite(temp_r1 / temp_r2)) {
    URs[eBBY] = Pr[eBY] * temp_r1 / temp_r2;
    URs[eBBZ] = Pr[eBZ] * temp_r1 / temp_r2;
  }
  temp_l1 = Pl[eVX] * Bx + Pl[eVY] * Pl[eBY] + Pl[eVZ] * Pl[eBZ];
  temp_r1 = Pr[eVX] * Bx + Pr[eVY] * Pr[eBY] + Pr[eVZ] * Pr[eBZ];
  temp_l2 = lam[2] * ULs[eBBX] + vys_l * ULs[eBBY] + vzs_l * ULs[eBBZ];
  temp_r2 = lam[2] * URs[eBBX] + vys_r * URs[eBBY] + vzs_r * URs[eBBZ];
  ULs[ERG] = (sl_vl * UL[ERG] - tp_l * Pl[eVX] + tp_s * lam[2] + Bx * (temp_l1 - temp_l2)) / sl_sm;
  URs[ERG] = (sr_vr * UR[ERG] - tp_r * Pr[eVX] + tp_s * lam[2] + Bx * (temp_r1 - temp_r2)) / sr_sm;
  lam[1] = lam[2] - fabs(Bx) / sqrt(ULs[eRHO]);
  lam[3] = lam[2] + fabs(Bx) / sqrt(URs[eRHO]);
  if (Bx == 0) {
    for (int v = 0; v < 8; v++) { ULss[v] = ULs[v]; URss[v] = URs[v]; }
  } else {
    ULss[eRHO] = ULs[eRHO];
    URss[eRHO] = URs[eRHO];
    double sgn = (Bx > 0) - (Bx < 0);
    temp_l1 = sqrt(ULs[eRHO]);
    temp_r1 = sqrt(URs[eRHO]);
    temp = temp_l1 + temp_r1;
    ULss[eMX] = lam[2] * ULss[eRHO];
    URss[eMX] = lam[2] * URss[eRHO];
    double vy_ss = (temp_l1 * vys_l + temp_r1 * vys_r + (URs[eBBY] - ULs[eBBY]) * sgn) / temp;
    ULss[eMY] = vy_ss * ULss[eRHO];
    URss[eMY] = vy_ss * URss[eRHO];
    double vz_ss = (temp_l1 * vzs_l + temp_r1 * vzs_r + (URs[eBBZ] - ULs[eBBZ]) * sgn) / temp;
    ULss[eMZ] = vz_ss * ULss[eRHO];
    URss[eMZ] = vz_ss * URss[eRHO];
    ULss[eBBX] = URss[eBBX] = Bx;
    ULss[eBBY] = URss[eBBY] = (temp_l1 * URs[eBBY] + temp_r1 * ULs[eBBY] + temp_l1 * temp_r1 * (vys_r - vys_l) * sgn) / temp;
    ULss[eBBZ] = URss[eBBZ] = (temp_l1 * URs[eBBZ] + temp_r1 * ULs[eBBZ] + temp_l1 * temp_r1 * (vzs_r - vzs_l) * sgn) / temp;
    temp = lam[2] * ULss[eBBX] + vy_ss * ULss[eBBY] + vz_ss * ULss[eBBZ];
    ULss[ERG] = ULs[ERG] - temp_l1 * (temp_l2 - temp) * sgn;
    URss[ERG] = URs[ERG] + temp_r1 * (temp_r2 - temp) * sgn;
  }
  if (lam[0] > 0) {
    for (int v = 0; v < 8; v++) { flux[v] = FL[v]; ustar[v] = UL[v]; }
  } else if (lam[1] >= 0) {
    for (int v = 0; v < 8; v++) { flux[v] = FL[v] + lam[0] * (ULs[v] - UL[v]); ustar[v] = ULs[v]; }
  } else if (lam[2] >= 0) {
    for (int v = 0; v < 8; v++) {
      flux[v] = FL[v] + lam[1] * ULss[v] - (lam[1] - lam[0]) * ULs[v] - lam[0] * UL[v];
      ustar[v] = ULss[v];
    }
  } else if (lam[3] >= 0) {
    for (int v = 0; v < 8; v++) {
      flux[v] = FR[v] + lam[3] * URss[v] - (lam[3] - lam[4]) * URs[v] - lam[4] * UR[v];
      ustar[v] = URss[v];
    }
  } else if (lam[4] >= 0) {
    for (int v = 0; v < 8; v++) { flux[v] = FR[v] + lam[4] * (URs[v] - UR[v]); ustar[v] = URs[v]; }
  } else {
    for (int v = 0; v < 8; v++) { flux[v] = FR[v]; ustar[v] = UR[v]; }
  }
}

/* Riemann_Roe_MHD_CV::MHD_Roe_CV_flux_solver_symmetric
 * (Roe_MHD_ConservedVar_solver.cpp:218-262) and helpers:
 * Roe_get_average_state :300-352, Roe_get_difference_states :358-393,
 * Roe_get_wavespeeds :399-470, Roe_get_eigenvalues :476-510 (H-correction),
 * Roe_get_wavestrengths :516-580, Roe_get_right_evectors :699-810 (Cargo &
 * Gallice 1997), calculate_symmetric_flux :1074-1131, set_pstar_from_meanp :283 */
static void mhd_RoeCV(pion_oracle *s, const double *left, const double *right, double hc_etamax, double *pstar,
                      double *flux) {
  enum { FN = 0, AN = 1, SN = 2, CT = 3, SP = 4, AP = 5, FP = 6 };
  const double g = s->gamma;
  const int nv = s->nv, eHH = PG;
  const int eVX = s->eVX, eVY = s->eVY, eVZ = s->eVZ, eBX = s->eBX, eBY = s->eBY, eBZ = s->eBZ;
  double UL[PO_MAXVAR] = {0}, UR[PO_MAXVAR] = {0}, meanp[PO_MAXVAR] = {0}, udiff[PO_MAXVAR], pdiff[PO_MAXVAR];
  double ev[7], str[7], rev[7][7];
  mhd_ideal_PtoU(s, left, UL);
  mhd_ideal_PtoU(s, right, UR);
  /* average state */
  double rl = sqrt(left[RO]), rr = sqrt(right[RO]);
  double lH = (left[RO] * (left[eVX] * left[eVX] + left[eVY] * left[eVY] + left[eVZ] * left[eVZ]) / 2.0 +
               (g * left[PG] / (g - 1.0)) + (left[eBX] * left[eBX] + left[eBY] * left[eBY] + left[eBZ] * left[eBZ])) / left[RO];
  double rH = (right[RO] * (right[eVX] * right[eVX] + right[eVY] * right[eVY] + right[eVZ] * right[eVZ]) / 2.0 +
               (g * right[PG] / (g - 1.0)) + (right[eBX] * right[eBX] + right[eBY] * right[eBY] + right[eBZ] * right[eBZ])) / right[RO];
  double Roe_denom = 1.0 / (rl + rr);
  meanp[RO] = rl * rr;
  meanp[eVX] = (rl * left[eVX] + rr * right[eVX]) * Roe_denom;
  meanp[eVY] = (rl * left[eVY] + rr * right[eVY]) * Roe_denom;
  meanp[eVZ] = (rl * left[eVZ] + rr * right[eVZ]) * Roe_denom;
  meanp[eBY] = (rr * left[eBY] + rl * right[eBY]) * Roe_denom;
  meanp[eBZ] = (rr * left[eBZ] + rl * right[eBZ]) * Roe_denom;
  meanp[eBX] = 0.5 * (left[eBX] + right[eBX]);
  int signBX = (meanp[eBX] >= 0.0) ? 1 : -1;
  meanp[eHH] = (rl * lH + rr * rH) * Roe_denom;
  double Roe_V = sqrt(meanp[eVX] * meanp[eVX] + meanp[eVY] * meanp[eVY] + meanp[eVZ] * meanp[eVZ]);
  double Roe_B = sqrt(meanp[eBX] * meanp[eBX] + meanp[eBY] * meanp[eBY] + meanp[eBZ] * meanp[eBZ]);
  double Roe_Bt = sqrt(meanp[eBY] * meanp[eBY] + meanp[eBZ] * meanp[eBZ]);
  double betay, betaz;
  if (Roe_Bt >= TINYVALUE) { betay = meanp[eBY] / Roe_Bt; betaz = meanp[eBZ] / Roe_Bt; }
  else { betay = 1.0 / sqrt(2.0); betaz = 1.0 / sqrt(2.0); }
  /* difference states */
  for (int v = 0; v < nv; v++) { udiff[v] = UR[v] - UL[v]; pdiff[v] = right[v] - left[v]; }
  udiff[s->eBBX] = pdiff[eBX] = 0.0;
  double CGX = (pdiff[eBY] * pdiff[eBY] + pdiff[eBZ] * pdiff[eBZ]) * 0.5 * Roe_denom * Roe_denom;
  pdiff[PG] = ((0.5 * Roe_V * Roe_V - CGX) * pdiff[RO] -
               (meanp[eVX] * udiff[s->eMX] + meanp[eVY] * udiff[s->eMY] + meanp[eVZ] * udiff[s->eMZ]) + udiff[ERG] -
               (meanp[eBY] * pdiff[eBY] + meanp[eBZ] * pdiff[eBZ])) * (g - 1.0);
  /* wavespeeds */
  double b2 = Roe_B * Roe_B / meanp[RO];
  double Roe_a = sqrt((2.0 - g) * CGX + (g - 1.0) * fmax((meanp[eHH] - 0.5 * Roe_V * Roe_V - b2), 1.0e-12 * Roe_V * Roe_V));
  double astar2 = Roe_a * Roe_a + b2;
  double Roe_ca = sqrt(meanp[eBX] * meanp[eBX] / meanp[RO]);
  double Roe_cs = astar2 * astar2 - 4.0 * Roe_a * Roe_a * Roe_ca * Roe_ca;
  if (Roe_cs <= 0.0) Roe_cs = 0.0; else Roe_cs = sqrt(Roe_cs);
  double Roe_cf = sqrt(0.5 * (astar2 + Roe_cs));
  Roe_cs = astar2 - Roe_cs;
  if (Roe_cs <= 0.0) Roe_cs = 0.0; else Roe_cs = sqrt(0.5 * Roe_cs);
  if (Roe_ca > Roe_cf) Roe_ca = Roe_cf;
  if (Roe_cs > Roe_ca) Roe_cs = Roe_ca;
  double cf2diff, alphaf, alphas;
  if ((cf2diff = Roe_cf * Roe_cf - Roe_cs * Roe_cs) > MACHINEACCURACY) {
    if ((alphaf = Roe_a * Roe_a - Roe_cs * Roe_cs) < 0.0) alphaf = 0.;
    if ((alphas = Roe_cf * Roe_cf - Roe_a * Roe_a) < 0.0) alphas = 0.;
    if ((alphaf = sqrt(alphaf / cf2diff)) > 1.0) alphaf = 1.0;
    if ((alphas = sqrt(alphas / cf2diff)) > 1.0) alphas = 1.0;
  } else {
    alphaf = alphas = 1.0 / sqrt(2.0);
  }
  /* eigenvalues + H-correction */
  ev[FN] = meanp[eVX] - Roe_cf; ev[AN] = meanp[eVX] - Roe_ca; ev[SN] = meanp[eVX] - Roe_cs; ev[CT] = meanp[eVX];
  ev[SP] = meanp[eVX] + Roe_cs; ev[AP] = meanp[eVX] + Roe_ca; ev[FP] = meanp[eVX] + Roe_cf;
  for (int v = 0; v < 7; v++) {
    if (ev[v] < 0.0) ev[v] = fmin(ev[v], -hc_etamax);
    else ev[v] = fmax(ev[v], hc_etamax);
  }
  /* wave strengths */
  str[FN] = 0.5 * (alphaf * (CGX * pdiff[RO] + pdiff[PG]) +
                   meanp[RO] * alphas * Roe_cs * signBX * (betay * pdiff[eVY] + betaz * pdiff[eVZ]) -
                   meanp[RO] * alphaf * Roe_cf * pdiff[eVX] +
                   sqrt(meanp[RO]) * alphas * Roe_a * (betay * pdiff[eBY] + betaz * pdiff[eBZ]));
  str[FP] = 0.5 * (alphaf * (CGX * pdiff[RO] + pdiff[PG]) -
                   meanp[RO] * alphas * Roe_cs * signBX * (betay * pdiff[eVY] + betaz * pdiff[eVZ]) +
                   meanp[RO] * alphaf * Roe_cf * pdiff[eVX] +
                   sqrt(meanp[RO]) * alphas * Roe_a * (betay * pdiff[eBY] + betaz * pdiff[eBZ]));
  str[SN] = 0.5 * (alphas * (CGX * pdiff[RO] + pdiff[PG]) -
                   meanp[RO] * alphaf * Roe_cf * signBX * (betay * pdiff[eVY] + betaz * pdiff[eVZ]) -
                   meanp[RO] * alphas * Roe_cs * pdiff[eVX] -
                   sqrt(meanp[RO]) * alphaf * Roe_a * (betay * pdiff[eBY] + betaz * pdiff[eBZ]));
  str[SP] = 0.5 * (alphas * (CGX * pdiff[RO] + pdiff[PG]) +
                   meanp[RO] * alphaf * Roe_cf * signBX * (betay * pdiff[eVY] + betaz * pdiff[eVZ]) +
                   meanp[RO] * alphas * Roe_cs * pdiff[eVX] -
                   sqrt(meanp[RO]) * alphaf * Roe_a * (betay * pdiff[eBY] + betaz * pdiff[eBZ]));
  str[AN] = 0.5 * (+betay * pdiff[eVZ] - betaz * pdiff[eVY] +
                   signBX * (betay * pdiff[eBZ] - betaz * pdiff[eBY]) / sqrt(meanp[RO]));
  str[AP] = 0.5 * (-betay * pdiff[eVZ] + betaz * pdiff[eVY] +
                   signBX * (betay * pdiff[eBZ] - betaz * pdiff[eBY]) / sqrt(meanp[RO]));
  str[CT] = (Roe_a * Roe_a - CGX) * pdiff[RO] - pdiff[PG];
  /* right eigenvectors; component order {rho, mx, my, mz, by, bz, e} */
  double rootrho = sqrt(meanp[RO]);
  rev[CT][0] = 1; rev[CT][1] = meanp[eVX]; rev[CT][2] = meanp[eVY]; rev[CT][3] = meanp[eVZ];
  rev[CT][4] = 0.0; rev[CT][5] = 0.0;
  rev[CT][6] = 0.5 * Roe_V * Roe_V + CGX * (g - 2) / (g - 1);
  for (int v = 0; v < 7; v++) rev[CT][v] /= Roe_a * Roe_a;
  rev[AN][0] = 0.0; rev[AN][1] = 0.0;
  rev[AN][2] = -meanp[RO] * betaz;
  rev[AN][3] = +meanp[RO] * betay;
  rev[AN][4] = -signBX * rootrho * betaz;
  rev[AN][5] = +signBX * rootrho * betay;
  rev[AN][6] = -meanp[RO] * (meanp[eVY] * betaz - meanp[eVZ] * betay);
  rev[AP][0] = 0.0; rev[AP][1] = 0.0;
  rev[AP][2] = -rev[AN][2]; rev[AP][3] = -rev[AN][3]; rev[AP][4] = rev[AN][4]; rev[AP][5] = rev[AN][5];
  rev[AP][6] = -rev[AN][6];
  double das = meanp[RO] * alphas, daf = meanp[RO] * alphaf;
  rev[SN][0] = das;
  rev[SN][1] = das * (meanp[eVX] - Roe_cs);
  rev[SN][2] = das * meanp[eVY] - daf * Roe_cf * betay * signBX;
  rev[SN][3] = das * meanp[eVZ] - daf * Roe_cf * betaz * signBX;
  rev[SN][4] = -rootrho * alphaf * Roe_a * betay;
  rev[SN][5] = -rootrho * alphaf * Roe_a * betaz;
  rev[SN][6] = das * (meanp[eHH] - Roe_B * Roe_B / meanp[RO] - meanp[eVX] * Roe_cs) -
               daf * Roe_cf * signBX * (meanp[eVY] * betay + meanp[eVZ] * betaz) - rootrho * alphaf * Roe_a * Roe_Bt;
  rev[SP][0] = das;
  rev[SP][1] = das * (meanp[eVX] + Roe_cs);
  rev[SP][2] = das * meanp[eVY] + daf * Roe_cf * betay * signBX;
  rev[SP][3] = das * meanp[eVZ] + daf * Roe_cf * betaz * signBX;
  rev[SP][4] = rev[SN][4];
  rev[SP][5] = rev[SN][5];
  rev[SP][6] = das * (meanp[eHH] - Roe_B * Roe_B / meanp[RO] + meanp[eVX] * Roe_cs) +
               daf * Roe_cf * signBX * (meanp[eVY] * betay + meanp[eVZ] * betaz) - rootrho * alphaf * Roe_a * Roe_Bt;
  rev[FN][0] = daf;
  rev[FN][1] = daf * (meanp[eVX] - Roe_cf);
  rev[FN][2] = daf * meanp[eVY] + das * Roe_cs * betay * signBX;
  rev[FN][3] = daf * meanp[eVZ] + das * Roe_cs * betaz * signBX;
  rev[FN][4] = rootrho * alphas * Roe_a * betay;
  rev[FN][5] = rootrho * alphas * Roe_a * betaz;
  rev[FN][6] = daf * (meanp[eHH] - Roe_B * Roe_B / meanp[RO] - meanp[eVX] * Roe_cf) +
               das * Roe_cs * signBX * (meanp[eVY] * betay + meanp[eVZ] * betaz) + rootrho * alphas * Roe_a * Roe_Bt;
  rev[FP][0] = daf;
  rev[FP][1] = daf * (meanp[eVX] + Roe_cf);
  rev[FP][2] = daf * meanp[eVY] - das * Roe_cs * betay * signBX;
  rev[FP][3] = daf * meanp[eVZ] - das * Roe_cs * betaz * signBX;
  rev[FP][4] = rev[FN][4];
  rev[FP][5] = rev[FN][5];
  rev[FP][6] = daf * (meanp[eHH] - Roe_B * Roe_B / meanp[RO] + meanp[eVX] * Roe_cf) -
               das * Roe_cs * signBX * (meanp[eVY] * betay + meanp[eVZ] * betaz) + rootrho * alphas * Roe_a * Roe_Bt;
  double norm = meanp[RO] * Roe_a * Roe_a;
  for (int v = 0; v < 7; v++) rev[SN][v] /= norm;
  for (int v = 0; v < 7; v++) rev[SP][v] /= norm;
  for (int v = 0; v < 7; v++) rev[FN][v] /= norm;
  for (int v = 0; v < 7; v++) rev[FP][v] /= norm;
  /* symmetric flux */
  double FR[PO_MAXVAR];
  mhd_PUtoFlux(s, left, UL, flux);
  mhd_PUtoFlux(s, right, UR, FR);
  for (int v = 0; v < 8; v++) flux[v] += FR[v];
  for (int iw = 0; iw < 7; iw++) {
    flux[RHO] -= str[iw] * fabs(ev[iw]) * rev[iw][0];
    flux[s->eMX] -= str[iw] * fabs(ev[iw]) * rev[iw][1];
    flux[s->eMY] -= str[iw] * fabs(ev[iw]) * rev[iw][2];
    flux[s->eMZ] -= str[iw] * fabs(ev[iw]) * rev[iw][3];
    flux[s->eBBY] -= str[iw] * fabs(ev[iw]) * rev[iw][4];
    flux[s->eBBZ] -= str[iw] * fabs(ev[iw]) * rev[iw][5];
    flux[ERG] -= str[iw] * fabs(ev[iw]) * rev[iw][6];
  }
  for (int v = 0; v < 8; v++) flux[v] *= 0.5;
  for (int v = 0; v < nv; v++) pstar[v] = meanp[v];
  pstar[PG] = pstar[RO] * Roe_a * Roe_a / g;
}

/* ------------------------------------------------------------------ */
/* spatial solver: inter-cell flux                                     */
/* ------------------------------------------------------------------ */
/* FV_solver_base::select_Hcorr_eta (solver_eqn_base.cpp:608-678).  NB the
 * "negative" neighbours are taken along the CURRENT axis (:659-667). */
static double select_Hcorr_eta(const pion_oracle *s, long cl, long cr) {
  int axis = s->dir, nd = s->ndim;
  double eta = s->hcorr[cl * 3 + axis];
  if (nd == 1) return eta;
  int perp = (axis + 1) % nd;
  eta = fmax(eta, s->hcorr[cl * 3 + perp]);
  eta = fmax(eta, s->hcorr[cr * 3 + perp]);
  if (nd > 2) {
    perp = (axis + 2) % nd;
    eta = fmax(eta, s->hcorr[cl * 3 + perp]);
    eta = fmax(eta, s->hcorr[cr * 3 + perp]);
  }
  for (int idim = 1; idim < nd; idim++) {
    perp = (axis + idim) % nd;
    int negdir = axis * 2;
    long cneg = nextpt(s, cl, negdir);
    if (cneg >= 0) eta = fmax(eta, s->hcorr[cneg * 3 + perp]);
    cneg = nextpt(s, cr, negdir);
    if (cneg >= 0) eta = fmax(eta, s->hcorr[cneg * 3 + perp]);
  }
  return eta;
}

/* eqns_Euler::Enthalpy (eqns_hydro_adiabatic.cpp:356-364) */
static inline double euler_enthalpy(const pion_oracle *s, const double *p) {
  double g = s->gamma;
  return (0.5 * (p[s->eVX] * p[s->eVX] + p[s->eVY] * p[s->eVY] + p[s->eVZ] * p[s->eVZ]) + g * p[PG] / (g - 1.0) / p[RO]);
}

/* Riemann_FVS_Euler::FVS_flux + Roe_average_state (Riemann_FVS_hydro.cpp:84-248): van Leer (1982) flux
 * vector splitting; pstar = the Roe-average state (for the viscosity) */
static void hydro_FVS(pion_oracle *s, const double *pl, const double *pr, double *flux, double *pstar) {
  double g = s->gamma;
  double fpos[PO_MAXVAR], fneg[PO_MAXVAR], utemp[PO_MAXVAR];
  for (int v = 0; v < s->nv; v++) fpos[v] = fneg[v] = 0.0;
  double cl = chydro(s, pl), cr = chydro(s, pr), Ml = pl[s->eVX] / cl, Mr = pr[s->eVX] / cr, f1 = 0.0, f2 = 0.0;
  if (Ml < -1.0) {
    /* zero */
  } else if (Ml > 1.0) {
    euler_PtoU(s, pl, utemp);
    euler_PUtoFlux(s, pl, utemp, fpos);
  } else {
    f1 = 0.25 * pl[RO] * cl * (1.0 + Ml) * (1.0 + Ml);
    f2 = cl * ((g - 1.0) * Ml + 2);
    fpos[RHO] = f1;
    fpos[s->eMX] = f1 * f2 / g;
    fpos[s->eMY] = f1 * pl[s->eVY];
    fpos[s->eMZ] = f1 * pl[s->eVZ];
    fpos[ERG] = f1 * (f2 * f2 * 0.5 / (g * g - 1.0) + 0.5 * (pl[s->eVY] * pl[s->eVY] + pl[s->eVZ] * pl[s->eVZ]));
  }
  if (Mr > 1.0) {
    /* zero */
  } else if (Mr < -1.0) {
    euler_PtoU(s, pr, utemp);
    euler_PUtoFlux(s, pr, utemp, fneg);
  } else {
    f1 = -0.25 * pr[RO] * cr * (1.0 - Mr) * (1.0 - Mr);
    f2 = cr * ((g - 1.0) * Mr - 2);
    fneg[RHO] = f1;
    fneg[s->eMX] = f1 * f2 / g;
    fneg[s->eMY] = f1 * pr[s->eVY];
    fneg[s->eMZ] = f1 * pr[s->eVZ];
    fneg[ERG] = f1 * (f2 * f2 * 0.5 / (g * g - 1) + 0.5 * (pr[s->eVY] * pr[s->eVY] + pr[s->eVZ] * pr[s->eVZ]));
  }
  /* only the five hydro components (rs_nvar = 5): tracer entries of flux stay 0 */
  flux[RHO] = fpos[RHO] + fneg[RHO];
  flux[ERG] = fpos[ERG] + fneg[ERG];
  flux[s->eMX] = fpos[s->eMX] + fneg[s->eMX];
  flux[s->eMY] = fpos[s->eMY] + fneg[s->eMY];
  flux[s->eMZ] = fpos[s->eMZ] + fneg[s->eMZ];
  /* Roe_average_state (:205-248) */
  double rl = sqrt(pl[RO]), rr = sqrt(pr[RO]), denom = 1.0 / (rl + rr);
  pstar[RO] = rl * rr;
  pstar[s->eVX] = (rl * pl[s->eVX] + rr * pr[s->eVX]) * denom;
  pstar[s->eVY] = (rl * pl[s->eVY] + rr * pr[s->eVY]) * denom;
  pstar[s->eVZ] = (rl * pl[s->eVZ] + rr * pr[s->eVZ]) * denom;
  pstar[PG] = denom * (rl * euler_enthalpy(s, pl) + rr * euler_enthalpy(s, pr));
  pstar[PG] = (g - 1.0) * (pstar[PG] - 0.5 * (pstar[s->eVX] * pstar[s->eVX] + pstar[s->eVY] * pstar[s->eVY] +
                                               pstar[s->eVZ] * pstar[s->eVZ]));
  pstar[PG] = pstar[RO] * pstar[PG] / g;
}

/* Riemann_Roe_Hydro_PV::Roe_prim_var_solver (Roe_Hydro_PrimitiveVar_solver.cpp:62-209): linearised
 * primitive-variable solver about the Roe-average state; returns the interface state */
static void hydro_RoePV(pion_oracle *s, const double *left, const double *right, double *pstar) {
  double g = s->gamma;
  double rl = sqrt(left[RO]), rr = sqrt(right[RO]), lH = euler_enthalpy(s, left), rH = euler_enthalpy(s, right),
         denom = 1.0 / (rl + rr), a_mean = 0.0, v2_mean = 0.0;
  double m_ro = rl * rr;
  double m_vx = (rl * left[s->eVX] + rr * right[s->eVX]) * denom;
  double m_vy = (rl * left[s->eVY] + rr * right[s->eVY]) * denom;
  double m_vz = (rl * left[s->eVZ] + rr * right[s->eVZ]) * denom;
  double m_H = (rl * lH + rr * rH) * denom;
  v2_mean = m_vx * m_vx + m_vy * m_vy + m_vz * m_vz;
  a_mean = sqrt((g - 1.0) * (m_H - 0.5 * v2_mean));
  if (m_vx - a_mean >= 0.) {
    for (int i = 0; i < 5; i++) pstar[i] = left[i];
  } else if (m_vx + a_mean <= 0.) {
    for (int i = 0; i < 5; i++) pstar[i] = right[i];
  } else {
    pstar[PG] = 0.5 * (left[PG] + right[PG] - m_ro * a_mean * (right[s->eVX] - left[s->eVX]));
    pstar[s->eVX] = 0.5 * (left[s->eVX] + right[s->eVX] - (right[PG] - left[PG]) / m_ro / a_mean);
    if (pstar[s->eVX] > 0.0) {
      pstar[RO] = left[RO] + m_ro * (left[s->eVX] - pstar[s->eVX]) / a_mean;
      pstar[s->eVY] = left[s->eVY];
      pstar[s->eVZ] = left[s->eVZ];
    } else {
      pstar[RO] = right[RO] + m_ro * (pstar[s->eVX] - right[s->eVX]) / a_mean;
      pstar[s->eVY] = right[s->eVY];
      pstar[s->eVZ] = right[s->eVZ];
    }
  }
}

/* eqns_mhd_ideal::UtoFlux (eqns_mhd_adiabatic.cpp:337-355) */
static void mhd_UtoFlux(const pion_oracle *s, const double *u, double *f) {
  double pm = (u[s->eBBX] * u[s->eBBX] + u[s->eBBY] * u[s->eBBY] + u[s->eBBZ] * u[s->eBBZ]) / 2.;
  double pg = (s->gamma - 1.) *
              (u[ERG] - (u[s->eMX] * u[s->eMX] + u[s->eMY] * u[s->eMY] + u[s->eMZ] * u[s->eMZ]) / (2. * u[RHO]) - pm);
  f[RHO] = u[s->eMX];
  f[s->eMX] = u[s->eMX] * u[s->eMX] / u[RHO] + pg + pm - u[s->eBBX] * u[s->eBBX];
  f[s->eMY] = u[s->eMX] * u[s->eMY] / u[RHO] - u[s->eBBX] * u[s->eBBY];
  f[s->eMZ] = u[s->eMX] * u[s->eMZ] / u[RHO] - u[s->eBBX] * u[s->eBBZ];
  f[ERG] = u[s->eMX] * (u[ERG] + pg + pm) / u[RHO] -
           u[s->eBBX] * (u[s->eMX] * u[s->eBBX] + u[s->eMY] * u[s->eBBY] + u[s->eMZ] * u[s->eBBZ]) / u[RHO];
  f[s->eBBX] = 0.;
  f[s->eBBY] = (u[s->eMX] * u[s->eBBY] - u[s->eMY] * u[s->eBBX]) / u[RHO];
  f[s->eBBZ] = (u[s->eMX] * u[s->eBBZ] - u[s->eMZ] * u[s->eBBX]) / u[RHO];
}
/* FV_solver_base::get_LaxFriedrichs_flux (solver_eqn_base.cpp:109-141): the tracer part is overwritten by
 * set_interface_tracer_flux; for GLM the psi component is overwritten by the Dedner flux (glm_inviscid_flux) */
static void lax_friedrichs_flux(pion_oracle *s, const double *l, const double *r, double *f) {
  double u1[PO_MAXVAR], u2[PO_MAXVAR], f1[PO_MAXVAR], f2[PO_MAXVAR];
  for (int v = 0; v < PO_MAXVAR; v++) u1[v] = u2[v] = f1[v] = f2[v] = 0.0;
  const int nq = (s->cfg.eqntype == PO_EQEUL) ? 5 : 8;
  if (s->cfg.eqntype == PO_EQEUL) {
    euler_PtoU(s, l, u1); euler_PtoU(s, r, u2);
    euler_UtoFlux(s, u1, f1); euler_UtoFlux(s, u2, f2);
  } else {
    mhd_PtoU(s, l, u1); mhd_PtoU(s, r, u2);
    mhd_UtoFlux(s, u1, f1); mhd_UtoFlux(s, u2, f2);
  }
  for (int v = 0; v < nq; v++) f[v] = 0.5 * (f1[v] + f2[v] + s->dx / s->FV_dt * (u1[v] - u2[v]) / s->ndim);
}

/* ------------------------------------------------------------------ */
/* Euler linear / exact / hybrid Riemann solvers (solverType 1, 2, 3)  */
/* Riemann_solvers/riemann.cpp + findroot.cpp + eqns_Euler::HydroWave  */
/* ------------------------------------------------------------------ */
#define BASEPG 1.e-5 /* constants.h:336 */
typedef struct {
  const pion_oracle *s;
  double left[5], right[5], pstar[5], cl, cr; /* natural order RO, PG, VX, VY, VZ of the SOLVER frame (eVX = sweep axis) */
} rs_euler;
/* eqns_Euler::HydroWave (eqns_hydro_adiabatic.cpp:221-262); lr: 0 = XN (left wave), 1 = XP; pre[] = {ro, pg, vx} */
static double hydro_wave(double gamma, int lr, double pp, const double *pre) {
  double pratio = pp / pre[1], u;
  double c0 = sqrt(gamma * pre[1] / pre[0]);
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
/* eqns_Euler::HydroWaveFull (:269-302) */
static void hydro_wave_full(double gamma, int lr, double pp, const double *pre, double *u, double *rho) {
  double pratio = pp / pre[1];
  *u = hydro_wave(gamma, lr, pp, pre);
  if (pratio < 1) *rho = pre[0] * exp(log(pratio) / gamma);
  else if (pratio > 1) *rho = pre[0] * (1 + pratio * (gamma + 1) / (gamma - 1.)) / ((gamma + 1.) / (gamma - 1.) + pratio);
  else *rho = pre[0];
}
/* riemann_Euler::FR_root_function (riemann.cpp:99-121) */
static double rs_root_function(const rs_euler *r, double pp) {
  double g = r->s->gamma;
  double L[3] = {r->left[0], r->left[1], r->left[2]}, R[3] = {r->right[0], r->right[1], r->right[2]};
  return hydro_wave(g, 1, pp, R) - hydro_wave(g, 0, pp, L);
}
/* findroot::bracket_root_pos (findroot.cpp:270-310): `factor` is a float in the reference */
static int rs_bracket_root_pos(const rs_euler *r, double *x1, double *x2) {
  float factor = 1.6;
  if (*x1 == *x2) return 1;
  if (*x1 > *x2) { double t = *x1; *x1 = *x2; *x2 = t; }
  double f1 = rs_root_function(r, *x1), f2 = rs_root_function(r, *x2);
  for (int j = 0; j < 50; j++) {
    if (f1 * f2 < 0) return 0;
    if (fabs(f1) < fabs(f2)) f1 = rs_root_function(r, *x1 *= 1. / factor);
    else f2 = rs_root_function(r, *x2 *= factor);
  }
  f1 = rs_root_function(r, *x1 = 0.);
  if (f1 * f2 < 0) return 0;
  *x1 = *x2 = 0.;
  return 1;
}
/* findroot::find_root_zbrent (findroot.cpp:359-452): NR zbrent with a purely relative tolerance */
static int rs_zbrent(const rs_euler *r, double x1, double x2, double tol, double *ans) {
  int ITMAX = 100;
  double EPS = MACHINEACCURACY;
  double a = x1, b = x2, c = x2, d, e, min1, min2;
  double fa = rs_root_function(r, a), fb = rs_root_function(r, b), fc, p, q, rr, sv, tol1, xm;
  d = 0.; e = 0.;
  if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) return 1;
  fc = fb;
  for (int iter = 1; iter <= ITMAX; iter++) {
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
/* riemann_Euler::check_wave_locations (riemann.cpp:471-585) */
static void rs_check_wave_locations(rs_euler *r) {
  const double g = r->s->gamma;
  double *ps = r->pstar;
  const double *L = r->left, *R = r->right;
  if (ps[1] < L[1]) {
    if (L[2] >= r->cl) { ps[1] = L[1]; ps[0] = L[0]; ps[2] = L[2]; return; }
    else if (ps[2] > 0.) {
      double cstar = sqrt(g * ps[1] / ps[0]);
      if (ps[2] > cstar) {
        ps[2] = (2. * r->cl + L[2] * (g - 1.)) / (g + 1.);
        ps[0] = L[0] * exp(2. / (g - 1.) * log(ps[2] / r->cl));
        ps[1] = exp(g * log(ps[0] / L[0])) * L[1];
        return;
      }
    }
  }
  if (ps[1] < R[1]) {
    if (R[2] <= -r->cr) { ps[1] = R[1]; ps[0] = R[0]; ps[2] = R[2]; return; }
    else if (ps[2] < 0.) {
      double cstar = sqrt(g * ps[1] / ps[0]);
      if (ps[2] < -cstar) {
        ps[2] = (-2. * r->cr + R[2] * (g - 1.)) / (g + 1.);
        ps[0] = R[0] * exp(2. / (g - 1.) * log(-ps[2] / r->cr));
        ps[1] = exp(g * log(ps[0] / R[0])) * R[1];
        return;
      }
    }
  }
  if (ps[1] > 1.0000001 * R[1]) {
    double vsh = R[2] + (ps[1] / R[1] - 1.) * r->cr * r->cr / g / (ps[2] - R[2]);
    if (vsh < 0.) { ps[1] = R[1]; ps[0] = R[0]; ps[2] = R[2]; return; }
  }
  if (ps[1] > 1.0000001 * L[1]) {
    double vsh = L[2] + (ps[1] / L[1] - 1.) * r->cl * r->cl / g / (ps[2] - L[2]);
    if (vsh > 0.) { ps[1] = L[1]; ps[0] = L[0]; ps[2] = L[2]; return; }
  }
}
/* riemann_Euler::linearOK (riemann.cpp:592-600) */
static int rs_linearOK(const rs_euler *r) {
  const double *L = r->left, *R = r->right;
  if ((fmax(L[1], R[1]) / fmin(L[1], R[1]) < 1.4) && (fmax(L[0], R[0]) / fmin(L[0], R[0]) < 1.4) &&
      (fabs(R[2] - L[2]) / fmin(r->cl, r->cr) < 0.03)) return 0;
  return 1;
}
/* riemann_Euler::linear_solver (riemann.cpp:674-747) */
static int rs_linear_solver(rs_euler *r) {
  const double g = r->s->gamma;
  double meanp[5];
  for (int i = 0; i < 5; i++) meanp[i] = (r->left[i] + r->right[i]) / 2.;
  double mcs = sqrt(g * meanp[1] / meanp[0]);
  double *ps = r->pstar;
  const double *L = r->left, *R = r->right;
  if (meanp[2] - mcs >= 0.) { for (int i = 0; i < 5; i++) ps[i] = L[i]; return 0; }
  else if (meanp[2] + mcs <= 0.) { for (int i = 0; i < 5; i++) ps[i] = R[i]; return 0; }
  ps[1] = 0.5 * (L[1] + R[1] - meanp[0] * mcs * (R[2] - L[2]));
  ps[2] = 0.5 * (L[2] + R[2] - (R[1] - L[1]) / meanp[0] / mcs);
  if (fabs(ps[2] / mcs) <= 1.e-6) ps[0] = meanp[0] * (2. + (L[2] - R[2]) / mcs) / 2.;
  else if (ps[2] > 0) ps[0] = L[0] + meanp[0] * (L[2] - ps[2]) / mcs;
  else if (ps[2] < 0) ps[0] = R[0] + meanp[0] * (ps[2] - R[2]) / mcs;
  else return 1;
  return 0;
}
/* riemann_Euler::exact_solver (riemann.cpp:754-822) with FR_find_root (:56-92) and findroot::solve_pos */
static int rs_exact_solver(rs_euler *r) {
  const double g = r->s->gamma;
  double *ps = r->pstar;
  int err = 0;
  {
    double x1 = (r->left[1] + r->right[1]) / 6.0, x2 = x1 * 9.0;
    if (rs_bracket_root_pos(r, &x1, &x2)) { ps[1] = -1.0; err += 1; }
    else if (rs_zbrent(r, x1, x2, 1.0e-8, &ps[1])) { ps[1] = -1.0; err += 1; }
  }
  double L[3] = {r->left[0], r->left[1], r->left[2]}, R[3] = {r->right[0], r->right[1], r->right[2]};
  hydro_wave_full(g, 0, ps[1], L, &ps[2], &ps[0]);
  double rhostar, temp;
  if ((ps[2] > 0) && (fabs(ps[2] / r->cr) > 1.e-6)) hydro_wave_full(g, 0, ps[1], L, &temp, &rhostar);
  else if ((ps[2] < 0) && (fabs(ps[2] / r->cr) > 1.e-6)) hydro_wave_full(g, 1, ps[1], R, &temp, &rhostar);
  else if (fabs(ps[2] / r->cr) <= 1.e-6) {
    hydro_wave_full(g, 0, ps[1], L, &temp, &rhostar);
    hydro_wave_full(g, 1, ps[1], R, &temp, &ps[0]);
    rhostar = (rhostar + ps[0]) / 2.0;
  } else { ps[0] = -1.0; return 1; }
  ps[0] = rhostar;
  if (err != 0) { ps[1] = ps[0] = ps[2] = -1.9; return 1; }
  rs_check_wave_locations(r);
  return 0;
}
/* riemann_Euler::solve_rarerare (riemann.cpp:829-885) */
static int rs_solve_rarerare(rs_euler *r) {
  const double g = r->s->gamma, cl = r->cl, cr = r->cr;
  double *ps = r->pstar;
  const double *L = r->left, *R = r->right;
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
/* eqns_Euler::SetAvgState (eqns_hydro_adiabatic.cpp:439-453, called once by the riemann_Euler constructor,
 * riemann.cpp:171): the solver's reference velocities are a tenth of the sound speed of RefVec, all three */
static double rs_refvel(const pion_oracle *s) {
  return 0.1 * sqrt(s->gamma * s->cfg.refvec[PG] / s->cfg.refvec[RO]);
}
/* riemann_Euler::solve_cavitation (riemann.cpp:892-960) */
static int rs_solve_cavitation(rs_euler *r) {
  const double g = r->s->gamma, cl = r->cl, cr = r->cr;
  double *ps = r->pstar;
  const double *L = r->left, *R = r->right;
  if ((L[2] - cl) >= 0.) { for (int i = 0; i < 5; i++) ps[i] = L[i]; return 0; }
  double temp = 2. / (g - 1.);
  if ((L[2] + temp * cl) >= 0.) {
    ps[2] = (2. * cl + L[2] * (g - 1.)) / (g + 1.);
    ps[0] = L[0] * exp(2. / (g - 1.) * log(ps[2] / cl));
    ps[1] = exp(g * log(ps[0] / L[0])) * L[1];
    return 0;
  }
  if ((R[2] - temp * cr) >= 0.) {
    /* eq_refvec[eqRO|eqPG|eqVX]: the solver's direction-rotated velocity index */
    ps[0] = r->s->cfg.refvec[RO] * BASEPG;
    ps[1] = r->s->cfg.refvec[PG] * BASEPG;
    ps[2] = rs_refvel(r->s) * BASEPG;
    return 0;
  }
  if ((R[2] + cr) > 0.) {
    ps[2] = (-2. * cr + R[2] * (g - 1.)) / (g + 1.);
    ps[0] = R[0] * exp(2. / (g - 1.) * log(-ps[2] / cr));
    ps[1] = exp(g * log(ps[0] / R[0])) * R[1];
    return 0;
  }
  if ((R[2] + cr) <= 0.) { for (int i = 0; i < 5; i++) ps[i] = R[i]; return 0; }
  return 1;
}
/* riemann_Euler::JMs_riemann_solve (riemann.cpp:245-463); l, r, ans are grid-frame primitive vectors */
static int rs_euler_solve(pion_oracle *s, const double *l, const double *rgt, double *ans, int mode) {
  rs_euler r;
  r.s = s;
  const int ix[5] = {RO, PG, s->eVX, s->eVY, s->eVZ};
  for (int v = 0; v < 5; v++) { r.left[v] = l[ix[v]]; r.right[v] = rgt[ix[v]]; r.pstar[v] = 0.0; }
  const double g = s->gamma;
  /* "same state" shortcut: the sum runs over the UNROTATED components with the unrotated reference vector */
  double diff = 0.;
  const double refvel = rs_refvel(s);
  for (int i = 0; i < 5; i++) diff += fabs(rgt[i] - l[i]) / (fabs(i >= 2 ? refvel : s->cfg.refvec[i]) + TINYVALUE);
  if (diff < 1.e-6) {
    for (int i = 0; i < 5; i++) ans[i] = (l[i] + rgt[i]) / 2.;
    return 0;
  }
  r.cl = sqrt(g * r.left[1] / r.left[0]);
  r.cr = sqrt(g * r.right[1] / r.right[0]);
  int err = 0, fail = 0;
  if ((r.right[2] - r.left[2]) <= 2. * (r.cl + sqrt((g - 1.) / 2. / g) * r.cr) / (g - 1.)) {
    if (mode == 1) {
      err = rs_linear_solver(&r);
      if (err) { r.pstar[1] = r.pstar[0] = r.pstar[2] = TINYVALUE; fail = 1; }
    } else if (mode == 2) {
      err = rs_exact_solver(&r);
      if (err) { r.pstar[1] = r.pstar[0] = TINYVALUE; fail = 1; }
    } else {
      err = rs_linear_solver(&r);
      if (err) r.pstar[1] = r.pstar[0] = r.pstar[2] = TINYVALUE;
      if (err != 0 || rs_linearOK(&r) != 0) {
        err = rs_exact_solver(&r);
        if (err) { r.pstar[1] = r.pstar[0] = TINYVALUE; fail = 1; }
      }
    }
  } else if ((r.right[2] - r.left[2]) <= 2. * (r.cl + r.cr) / (g - 1.)) {
    err = rs_solve_rarerare(&r);
    if (err) { r.pstar[1] = r.pstar[0] = r.pstar[2] = -1.9e99; fail = 1; }
  } else {
    err = rs_solve_cavitation(&r);
    if (err) { r.pstar[1] = r.pstar[0] = r.pstar[2] = -1.9e100; fail = 1; }
  }
  if (!fail) {
    if (r.pstar[2] > 0) { r.pstar[3] = r.left[3]; r.pstar[4] = r.left[4]; }
    else { r.pstar[3] = r.right[3]; r.pstar[4] = r.right[4]; }
    if (r.pstar[1] <= TINYVALUE) r.pstar[1] = BASEPG * s->cfg.refvec[PG];
    if (r.pstar[0] <= TINYVALUE) r.pstar[0] = BASEPG * s->cfg.refvec[RO];
  }
  /* the reference copies rs_pstar[i] -> ans[i] by NATURAL index: rs_pstar is indexed with the rotated eq* indices */
  for (int v = 0; v < 5; v++) ans[ix[v]] = r.pstar[v];
  return fail;
}
/* FV_solver_Hydro_Euler::inviscid_flux (solver_eqn_hydro_adi.cpp:94-205) */
static void euler_inviscid_flux(pion_oracle *s, const double *Pl, const double *Pr, double *flux, double *pstar) {
  double ustar[PO_MAXVAR];
  for (int v = 0; v < s->nv; v++) { ustar[v] = 0.0; flux[v] = 0.0; pstar[v] = 0.0; }
  if (s->cfg.solver == PO_FLUX_LF) { /* :142-148 */
    lax_friedrichs_flux(s, Pl, Pr, flux);
    for (int v = 0; v < 5; v++) pstar[v] = 0.5 * (Pl[v] + Pr[v]);
  } else if (s->cfg.solver >= 1 && s->cfg.solver <= 3) { /* :160-168: linear / exact / hybrid Riemann solver, then PtoFlux */
    if (rs_euler_solve(s, Pl, Pr, pstar, s->cfg.solver)) s->nfail_riemann++;
    euler_PtoU(s, pstar, ustar);
    euler_PUtoFlux(s, pstar, ustar, flux);
  } else if (s->cfg.solver == PO_FLUX_ROE) {
    hydro_RoeCV(s, Pl, Pr, s->HC_etamax, pstar, flux);
  } else if (s->cfg.solver == PO_FLUX_HLL) {
    hydro_HLL(s, Pl, Pr, flux, ustar);
    euler_UtoP(s, ustar, pstar);
  } else if (s->cfg.solver == PO_FLUX_FVS) {
    hydro_FVS(s, Pl, Pr, flux, pstar);
  } else if (s->cfg.solver == PO_FLUX_ROE_PV) {
    /* :178-187: interface state, then PtoFlux = (virtual) PtoU + PUtoFlux (eqns_base.cpp:230-240) */
    hydro_RoePV(s, Pl, Pr, pstar);
    euler_PtoU(s, pstar, ustar);
    euler_PUtoFlux(s, pstar, ustar, flux);
  } else {
    fprintf(stderr, "pion_oracle: Euler solver %d not restated\n", s->cfg.solver);
    abort();
  }
}
/* ---------------------------------------------------------------------------------------------------------
 * riemann_MHD: the linear MHD Riemann solver (solverType 1 with the MHD equations; Falle, Komissarov & Joarder 1998
 * with the Roe & Balsara eigenvector normalisation), Riemann_solvers/riemannMHD.cpp.
 * Solver-frame variables (enum rsvars, riemannMHD.h:56-65): RRO, RPG, RVX, RVY, RVZ, RBY, RBZ (+ RBX as a parameter).
 * ------------------------------------------------------------------------------------------------------- */
/* eqns_mhd_ideal::SetAvgState (eqns_mhd_adiabatic.cpp:501-543), called once by the riemann_MHD constructor
 * (riemannMHD.cpp:120) in direction XX: the solver's reference vector is RefVec[RO], RefVec[PG], a tenth of the fast
 * speed of RefVec (rotated so that B lies in the x-z plane) three times, |B(RefVec)| three times */
/* eqns_mhd_ideal::cfast (eqns_mhd_adiabatic.cpp:246-257) of an unrotated 8-vector */
static double rs_mhd_cfast_nat(const double *rv, double g) {
  const double ch = sqrt(g * rv[PG] / rv[RO]);
  double t1 = ch * ch + (rv[BX] * rv[BX] + rv[BY] * rv[BY] + rv[BZ] * rv[BZ]) / rv[RO];
  double t2 = 4. * ch * ch * rv[BX] * rv[BX] / rv[RO];
  t2 = fmax(MACHINEACCURACY, t1 * t1 - t2);
  return sqrt((t1 + sqrt(t2)) / 2.);
}
/* eqns_mhd_ideal::rotateXY (:423-437) */
static void rs_mhd_rotateXY(double *rv, double th) {
  const double ct = cos(th), st = sin(th);
  double vx = rv[VX] * ct - rv[VY] * st, vy = rv[VX] * st + rv[VY] * ct;
  rv[VX] = vx; rv[VY] = vy;
  vx = rv[BX] * ct - rv[BY] * st; vy = rv[BX] * st + rv[BY] * ct;
  rv[BX] = vx; rv[BY] = vy;
}
static void rs_mhd_refvec(const pion_oracle *s, double *refvel01, double *refB) {
  double rv[8];
  for (int v = 0; v < 8; v++) rv[v] = s->cfg.refvec[v];
  double angle = rv[BY] * rv[BY] + rv[BX] * rv[BX], refvel;
  if (angle > 10. * MACHINEACCURACY) {
    angle = M_PI / 2. - asin(rv[BY] / sqrt(angle));
    if (rv[BX] < 0) angle = -angle;
    rs_mhd_rotateXY(rv, angle);
    refvel = rs_mhd_cfast_nat(rv, s->gamma);
    rs_mhd_rotateXY(rv, -angle);
  } else {
    refvel = rs_mhd_cfast_nat(rv, s->gamma); /* maxspeed == cfast (eqns_mhd_adiabatic.h:120-123) */
  }
  *refB = sqrt(rv[BX] * rv[BX] + rv[BY] * rv[BY] + rv[BZ] * rv[BZ]);
  *refvel01 = 0.1 * refvel;
}
/* riemann_MHD::JMs_riemann_solve, mode 1 (riemannMHD.cpp:165-400) with get_sound_speeds (:555-765), get_eigenvalues
 * (:768-777), RoeBalsara_evectors (:965-1115), calculate_wave_strengths (:813-846), get_pstar (:849-960).
 * l, r, ans: grid-frame primitive vectors; returns 1 where the reference calls rep.error. */
static int rs_mhd_solve(pion_oracle *s, const double *l, const double *rgt, double *ans) {
  enum { RRO = 0, RPG = 1, RVX = 2, RVY = 3, RVZ = 4, RBY = 5, RBZ = 6, RBX = 7 };
  enum { FN = 0, AN = 1, SN = 2, CT = 3, SP = 4, AP = 5, FP = 6 };
  const int ix[8] = {RO, PG, s->eVX, s->eVY, s->eVZ, s->eBY, s->eBZ, s->eBX}; /* code2solvervars (:443-475) */
  const double g = s->gamma;
  const double smallB = MACHINEACCURACY, tinyB = smallB * smallB * smallB;
  double L[8], R[8], M[8], ps[8], ev[7], pdiff[7], str[7], lev[7][7], rev[7][7];
  for (int v = 0; v < 8; v++) { L[v] = l[ix[v]]; R[v] = rgt[ix[v]]; M[v] = 0.5 * (L[v] + R[v]); ps[v] = 0.0; }
  for (int v = 0; v < s->nv; v++) ans[v] = 0.0; /* RS_pstar[v >= 8] is never written: zero */
  const double ansBX = M[RBX];
  ps[RBX] = ansBX;
  /* same-state shortcut (:227-268): solver-frame differences over the UNROTATED entries 0..6 of the reference vector */
  double refvel01, refB;
  rs_mhd_refvec(s, &refvel01, &refB);
  const double refn[7] = {s->cfg.refvec[RO], s->cfg.refvec[PG], refvel01, refvel01, refvel01, refB, refB};
  double diff = 0.;
  for (int i = 0; i < 7; i++) diff += fabs(R[i] - L[i]) / (fabs(refn[i]) + TINYVALUE);
  int fail = 0;
  if (diff < 1.e-6) {
    for (int v = 0; v < 7; v++) ps[v] = M[v];
  } else {
    /* get_sound_speeds */
    const double ch = sqrt(g * M[RPG] / M[RRO]);
    const double bx = ansBX / sqrt(M[RRO]);
    const double ca = fabs(bx);
    const double bt = sqrt((M[RBY] * M[RBY] + M[RBZ] * M[RBZ]) / M[RRO]);
    double betay, betaz;
    if (bt > tinyB) { betay = M[RBY] / sqrt(M[RRO]) / bt; betaz = M[RBZ] / sqrt(M[RRO]) / bt; }
    else { betay = 1. / sqrt(2.); betaz = 1. / sqrt(2.); }
    if ((ch / ((ca < bt) ? bt : ca)) < sqrt(smallB)) fail = 1;
    double temp1 = ch * ch + bx * bx + bt * bt;
    double temp2 = 4. * ch * ch * bx * bx;
    if ((temp2 = temp1 * temp1 - temp2) < MACHINEACCURACY) temp2 = MACHINEACCURACY;
    double cf = sqrt((temp1 + sqrt(temp2)) / 2.);
    if ((temp2 = temp1 - sqrt(temp2)) < MACHINEACCURACY) temp2 = MACHINEACCURACY;
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
      fail = 1; /* "Near Triple degeneracy point": rep.error */
    }
    if ((cf <= 0.) || (cs < 0.) || (ca < 0.) || (ch <= 0.)) fail = 1;
    if (!fail) {
      /* get_eigenvalues */
      ev[FN] = M[RVX] - cf; ev[FP] = M[RVX] + cf; ev[AN] = M[RVX] - ca; ev[AP] = M[RVX] + ca;
      ev[SN] = M[RVX] - cs; ev[SP] = M[RVX] + cs; ev[CT] = M[RVX];
      /* RoeBalsara_evectors */
      const double r2 = sqrt(2.);
      const int sBx = (ansBX < 0.) ? -1 : 1;
      const double sro = sqrt(M[RRO]);
      lev[FN][RRO] = 0.0; lev[FN][RVX] = -alphaf * cf; lev[FN][RVY] = alphas * cs * sBx * betay; lev[FN][RVZ] = alphas * cs * sBx * betaz;
      lev[FN][RPG] = alphaf / M[RRO]; lev[FN][RBY] = alphas * ch * betay / sro; lev[FN][RBZ] = alphas * ch * betaz / sro;
      lev[AN][RRO] = 0.; lev[AN][RVX] = 0.; lev[AN][RVY] = sBx * betaz / r2; lev[AN][RVZ] = -sBx * betay / r2;
      lev[AN][RPG] = 0.; lev[AN][RBY] = betaz / sro / r2; lev[AN][RBZ] = -betay / sro / r2;
      lev[SN][RRO] = 0.0; lev[SN][RVX] = -alphas * cs; lev[SN][RVY] = -alphaf * cf * sBx * betay; lev[SN][RVZ] = -alphaf * cf * sBx * betaz;
      lev[SN][RPG] = alphas / M[RRO]; lev[SN][RBY] = -alphaf * ch * betay / sro; lev[SN][RBZ] = -alphaf * ch * betaz / sro;
      lev[CT][RRO] = 1.; lev[CT][RVX] = 0.; lev[CT][RVY] = 0.; lev[CT][RVZ] = 0.; lev[CT][RPG] = -1 / ch / ch; lev[CT][RBY] = 0.; lev[CT][RBZ] = 0.;
      lev[SP][RRO] = 0.0; lev[SP][RVX] = -lev[SN][RVX]; lev[SP][RVY] = -lev[SN][RVY]; lev[SP][RVZ] = -lev[SN][RVZ];
      lev[SP][RPG] = lev[SN][RPG]; lev[SP][RBY] = lev[SN][RBY]; lev[SP][RBZ] = lev[SN][RBZ];
      lev[AP][RRO] = 0.; lev[AP][RVX] = 0.; lev[AP][RVY] = lev[AN][RVY]; lev[AP][RVZ] = lev[AN][RVZ];
      lev[AP][RPG] = 0.; lev[AP][RBY] = -lev[AN][RBY]; lev[AP][RBZ] = -lev[AN][RBZ];
      lev[FP][RRO] = 0.0; lev[FP][RVX] = -lev[FN][RVX]; lev[FP][RVY] = -lev[FN][RVY]; lev[FP][RVZ] = -lev[FN][RVZ];
      lev[FP][RPG] = lev[FN][RPG]; lev[FP][RBY] = lev[FN][RBY]; lev[FP][RBZ] = lev[FN][RBZ];
      rev[FN][RRO] = alphaf * M[RRO]; rev[FN][RVX] = lev[FN][RVX]; rev[FN][RVY] = lev[FN][RVY]; rev[FN][RVZ] = lev[FN][RVZ];
      rev[FN][RPG] = alphaf * M[RRO] * ch * ch; rev[FN][RBY] = lev[FN][RBY] * M[RRO]; rev[FN][RBZ] = lev[FN][RBZ] * M[RRO];
      rev[AN][RRO] = 0.; rev[AN][RVX] = 0.; rev[AN][RVY] = lev[AN][RVY]; rev[AN][RVZ] = lev[AN][RVZ];
      rev[AN][RPG] = 0.; rev[AN][RBY] = lev[AN][RBY] * M[RRO]; rev[AN][RBZ] = lev[AN][RBZ] * M[RRO];
      rev[SN][RRO] = alphas * M[RRO]; rev[SN][RVX] = lev[SN][RVX]; rev[SN][RVY] = lev[SN][RVY]; rev[SN][RVZ] = lev[SN][RVZ];
      rev[SN][RPG] = alphas * M[RRO] * ch * ch; rev[SN][RBY] = lev[SN][RBY] * M[RRO]; rev[SN][RBZ] = lev[SN][RBZ] * M[RRO];
      rev[CT][RRO] = 1.0; rev[CT][RVX] = 0.; rev[CT][RVY] = 0.; rev[CT][RVZ] = 0.; rev[CT][RPG] = 0.; rev[CT][RBY] = 0.; rev[CT][RBZ] = 0.;
      rev[SP][RRO] = rev[SN][RRO]; rev[SP][RVX] = -rev[SN][RVX]; rev[SP][RVY] = -rev[SN][RVY]; rev[SP][RVZ] = -rev[SN][RVZ];
      rev[SP][RPG] = rev[SN][RPG]; rev[SP][RBY] = rev[SN][RBY]; rev[SP][RBZ] = rev[SN][RBZ];
      rev[AP][RRO] = 0.; rev[AP][RVX] = 0.; rev[AP][RVY] = rev[AN][RVY]; rev[AP][RVZ] = rev[AN][RVZ];
      rev[AP][RPG] = 0.; rev[AP][RBY] = -rev[AN][RBY]; rev[AP][RBZ] = -rev[AN][RBZ];
      rev[FP][RRO] = rev[FN][RRO]; rev[FP][RVX] = -rev[FN][RVX]; rev[FP][RVY] = -rev[FN][RVY]; rev[FP][RVZ] = -rev[FN][RVZ];
      rev[FP][RPG] = rev[FN][RPG]; rev[FP][RBY] = rev[FN][RBY]; rev[FP][RBZ] = rev[FN][RBZ];
      const double a22 = 1. / (2. * ch * ch);
      for (int i = 0; i < 7; i++) { lev[FN][i] *= a22; lev[SN][i] *= a22; lev[SP][i] *= a22; lev[FP][i] *= a22; }
      /* calculate_wave_strengths: strength = left eigenvector . (right - left), summed in rsvars order */
      for (int i = 0; i < 7; i++) pdiff[i] = R[i] - L[i];
      for (int w = 0; w < 7; w++) {
        double t = 0.0;
        for (int i = 0; i < 7; i++) t += lev[w][i] * pdiff[i];
        str[w] = t;
      }
      /* get_pstar: cross the waves with negative speed from the left; at a (nearly) stationary contact average
       * with the state reached from the right */
      int i = 0;
      const double evalacc = 1.e-4;
      for (int j = 0; j < 7; j++) ps[j] = L[j];
      while ((i < 7) && (ev[i] < 0.)) {
        for (int j = 0; j < 7; j++) ps[j] += str[i] * rev[i][j];
        i++;
      }
      if (fabs(M[RVX]) < (evalacc * ch)) {
        i = 6;
        for (int j = 0; j < 7; j++) pdiff[j] = R[j];
        while ((i >= 0) && (ev[i] > 0.)) {
          for (int j = 0; j < 7; j++) pdiff[j] -= str[i] * rev[i][j];
          i--;
        }
        for (int v = 0; v < 7; v++) ps[v] = 0.5 * (ps[v] + pdiff[v]);
      }
      if (ps[RPG] < 0.) ps[RPG] = refn[RPG] * BASEPG;
      if (ps[RRO] < 0.) ps[RRO] = refn[RRO] * BASEPG;
    }
  }
  for (int v = 0; v < 8; v++) ans[ix[v]] = ps[v]; /* solver2codevars (:480-512) */
  return fail;
}
/* FV_solver_mhd_ideal_adi::inviscid_flux (solver_eqn_mhd_adi.cpp:102-198) */
static void mhd_ideal_inviscid_flux(pion_oracle *s, long cl, long cr, const double *Pl, const double *Pr, double *flux,
                                    double *pstar) {
  double ustar[PO_MAXVAR];
  for (int v = 0; v < s->nv; v++) { ustar[v] = 0.0; flux[v] = 0.0; pstar[v] = 0.0; }
  if (s->cfg.solver == PO_FLUX_LF) { /* :132-136 */
    lax_friedrichs_flux(s, Pl, Pr, flux);
    for (int v = 0; v < s->nv; v++) pstar[v] = 0.5 * (Pl[v] + Pr[v]);
  } else if (s->cfg.solver >= 1 && s->cfg.solver <= 3) { /* :160-166: riemann_MHD only knows the linear solve: modes 2, 3 are fatal */
    if (s->cfg.solver != 1 || rs_mhd_solve(s, Pl, Pr, pstar)) s->nfail_riemann++;
    double us[PO_MAXVAR] = {0};
    mhd_ideal_PtoU(s, pstar, us); /* PtoFlux = PtoU + PUtoFlux (eqns_base.cpp:230-240); pstar[SI] is zero */
    mhd_PUtoFlux(s, pstar, us, flux);
  } else if (s->cfg.solver == PO_FLUX_ROE) {
    mhd_RoeCV(s, Pl, Pr, s->HC_etamax, pstar, flux);
  } else if (s->cfg.solver == PO_FLUX_HLLD) {
    double DivVl = s->divv[cl], DivVr = s->divv[cr], Gradl = s->gradp[cl], Gradr = s->gradp[cr];
    if ((DivVl < 0. && Gradl > 5.) || (DivVr < 0. && Gradr > 5.)) mhd_HLL(s, Pl, Pr, flux, ustar);
    else mhd_HLLD(s, Pl, Pr, flux, ustar);
    for (int v = 8; v < s->nv; v++) { flux[v] = 0.0; ustar[v] = 0.0; }
    mhd_UtoP(s, ustar, pstar);
  } else if (s->cfg.solver == PO_FLUX_HLL) {
    mhd_HLL(s, Pl, Pr, flux, ustar);
    mhd_UtoP(s, ustar, pstar);
  } else {
    fprintf(stderr, "pion_oracle: MHD solver %d not restated\n", s->cfg.solver);
    abort();
  }
}
/* FV_solver_mhd_mixedGLM_adi::inviscid_flux (solver_eqn_mhd_adi.cpp:662-772) */
static void glm_inviscid_flux(pion_oracle *s, long cl, long cr, const double *Pl, const double *Pr, double *flux,
                              double *pstar) {
  double left[PO_MAXVAR], right[PO_MAXVAR];
  for (int v = 0; v < s->nv; v++) { left[v] = Pl[v]; right[v] = Pr[v]; }
  double psistar = 0.5 * (left[SI] + right[SI] - (right[s->eBX] - left[s->eBX]));
  double bxstar = 0.5 * (left[s->eBX] + right[s->eBX] - (right[SI] - left[SI]));
  left[SI] = right[SI] = 0.0;
  left[s->eBX] = right[s->eBX] = bxstar;
  mhd_ideal_inviscid_flux(s, cl, cr, left, right, flux, pstar);
  flux[ERG] += s->chyp * bxstar * psistar;
  flux[s->eBBX] = s->chyp * psistar;
  flux[PSI] = s->chyp * bxstar;
}
/* FV_solver_Hydro_Euler::AVFalle (solver_eqn_hydro_adi.cpp:283-333) */
static void euler_AVFalle(const pion_oracle *s, const double *Pl, const double *Pr, const double *pstar, double *flux) {
  double prefactor = chydro(s, pstar) * s->cfg.etav * pstar[RO];
  double momvisc = prefactor * (Pr[s->eVX] - Pl[s->eVX]);
  double ergvisc = momvisc * pstar[s->eVX];
  flux[s->eMX] -= momvisc;
  momvisc = prefactor * (Pr[s->eVY] - Pl[s->eVY]);
  flux[s->eMY] -= momvisc;
  ergvisc += momvisc * pstar[s->eVY];
  momvisc = prefactor * (Pr[s->eVZ] - Pl[s->eVZ]);
  flux[s->eMZ] -= momvisc;
  ergvisc += momvisc * pstar[s->eVZ];
  flux[ERG] -= ergvisc;
}
/* FV_solver_mhd_ideal_adi::AVFalle (solver_eqn_mhd_adi.cpp:209-288); FV_etaB==FV_etav
 * (solver_eqn_base.cpp:82) */
static void mhd_AVFalle(const pion_oracle *s, const double *Pl, const double *Pr, const double *Pstar, double *flux) {
  double etav = s->cfg.etav, etaB = s->cfg.etav;
  double prefactor = cfast_components(0.5 * (Pl[RO] + Pr[RO]), 0.5 * (Pl[PG] + Pr[PG]), 0.5 * (Pl[s->eBX] + Pr[s->eBX]),
                                      0.5 * (Pl[s->eBY] + Pr[s->eBY]), 0.5 * (Pl[s->eBZ] + Pr[s->eBZ]), s->gamma) *
                     etav * Pstar[RO];
  double momvisc = prefactor * (Pr[s->eVX] - Pl[s->eVX]);
  double ergvisc = momvisc * Pstar[s->eVX];
  flux[s->eMX] -= momvisc;
  momvisc = prefactor * (Pr[s->eVY] - Pl[s->eVY]);
  flux[s->eMY] -= momvisc;
  ergvisc += momvisc * Pstar[s->eVY];
  momvisc = prefactor * (Pr[s->eVZ] - Pl[s->eVZ]);
  flux[s->eMZ] -= momvisc;
  ergvisc += momvisc * Pstar[s->eVZ];
  prefactor *= etaB / (etav * Pstar[RO]);
  momvisc = prefactor * (Pr[s->eBY] - Pl[s->eBY]);
  flux[s->eBBY] -= momvisc;
  ergvisc += momvisc * Pstar[s->eBY];
  momvisc = prefactor * (Pr[s->eBZ] - Pl[s->eBZ]);
  flux[s->eBBZ] -= momvisc;
  ergvisc += momvisc * Pstar[s->eBZ];
  flux[ERG] -= ergvisc;
}
/* FV_solver_base::InterCellFlux (solver_eqn_base.cpp:152-204) incl.
 * pre/post_calc_viscous_terms (:213-272) and set_interface_tracer_flux (:281-342) */
static void inter_cell_flux(pion_oracle *s, long cl, long cr, const double *lp, const double *rp, double *f) {
  double pstar[PO_MAXVAR];
  int av = s->cfg.artviscosity;
  if (av == PO_AV_HCORR || av == PO_AV_HCORR_FKJ98) s->HC_etamax = select_Hcorr_eta(s, cl, cr);
  if (s->cfg.eqntype == PO_EQEUL) euler_inviscid_flux(s, lp, rp, f, pstar);
  else if (s->cfg.eqntype == PO_EQMHD) mhd_ideal_inviscid_flux(s, cl, cr, lp, rp, f, pstar);
  else glm_inviscid_flux(s, cl, cr, lp, rp, f, pstar);
  if (av == PO_AV_FKJ98 || av == PO_AV_HCORR_FKJ98) {
    if (s->cfg.eqntype == PO_EQEUL) euler_AVFalle(s, lp, rp, pstar, f);
    else mhd_AVFalle(s, lp, rp, pstar, f);
  }
  if (s->ntr > 0) {
    double corrector[PO_MAXVAR];
    for (int v = 0; v < s->nv; v++) corrector[v] = 1.0;
    if (f[RHO] > 0.0) {
      if (s->have_mp) mp_sCMA(s, corrector, lp);
      for (int t = 0; t < s->ntr; t++) f[s->ftr + t] = lp[s->ftr + t] * f[RHO] * corrector[s->ftr + t];
    } else if (f[RHO] < 0.0) {
      if (s->have_mp) mp_sCMA(s, corrector, rp);
      for (int t = 0; t < s->ntr; t++) f[s->ftr + t] = rp[s->ftr + t] * f[RHO] * corrector[s->ftr + t];
    } else {
      for (int t = 0; t < s->ntr; t++) f[s->ftr + t] = 0.0;
    }
  }
}

/* ------------------------------------------------------------------ */
/* coord_sys/VectorOps (Cartesian)                                     */
/* ------------------------------------------------------------------ */
/* centre-of-volume radius of a cell: cylindrical VectorOps.h:414-418, spherical
 * VectorOps_spherical.h:188-197; R3 (spherical, :172-178) */
static inline double R_com_cyl(const pion_oracle *s, long c) {
  double R = dpos(s, c, 1);
  return R + s->dx * s->dx / 12. / R;
}
static inline double R_com_sph(const pion_oracle *s, long c) {
  double R = dpos(s, c, 0);
  double delta2 = s->dx / R;
  delta2 *= delta2;
  return R * (1.0 + 0.25 * delta2) / (1.0 + delta2 / 12.0);
}
static inline double R3_sph(const pion_oracle *s, long c) {
  double R = dpos(s, c, 0);
  return R + s->dx * s->dx / 12.0 / R;
}
/* is `axis` the radial axis of a curvilinear grid?  (Rcyl = YY in 2-D axisymmetry, Rsph = XX) */
static inline int radial_axis(const pion_oracle *s, int axis) {
  return (s->cfg.coord_sys == PO_COORD_CYL && axis == 1) || (s->cfg.coord_sys == PO_COORD_SPH && axis == 0);
}
static inline double R_com(const pion_oracle *s, long c) {
  return (s->cfg.coord_sys == PO_COORD_CYL) ? R_com_cyl(s, c) : R_com_sph(s, c);
}
/* BaseVectorOps::AvgFalle, AVG_MINMOD variant (VectorOps.cpp:40-59) */
static inline double avg_falle(double a, double b) {
  if (a * b <= VERY_TINY_VALUE) return 0.0;
  double r = a / b;
  return (r > 0.0) ? fmin(r, 1.0) * b : 0.0;
}
/* VectorOps_Cart::SetSlope (VectorOps.cpp:578-617) */
static void set_slope(const pion_oracle *s, long c, int axis, double *dpdx, int OA) {
  int nv = s->nv;
  if (OA == OA1) {
    for (int v = 0; v < nv; v++) dpdx[v] = 0.;
    return;
  }
  long cp = nextpt(s, c, 2 * axis + 1), cn = nextpt(s, c, 2 * axis);
  if (cp < 0) cp = nextpt(s, cn, 2 * axis + 1);
  if (cn < 0) cn = nextpt(s, cp, 2 * axis);
  const double *Pc = s->Ph + c * nv, *Pn = s->Ph + cn * nv, *Pp = s->Ph + cp * nv;
  double dx = s->dx;
  if (radial_axis(s, axis)) {
    /* VectorOps_Cyl::SetSlope case Rcyl (VectorOps.cpp:1158-1182), VectorOps_Sph::SetSlope
     * (VectorOps_spherical.cpp:372-380): divided differences between centres of volume; a missing
     * neighbour gives a zero one-sided slope (cyl :1159,:1167) */
    int noneg = nextpt(s, c, 2 * axis) < 0, nopos = nextpt(s, c, 2 * axis + 1) < 0;
    for (int v = 0; v < nv; v++) {
      double slpn = noneg ? 0.0 : (Pc[v] - Pn[v]) / (R_com(s, c) - R_com(s, cn));
      double slpp = nopos ? 0.0 : (Pp[v] - Pc[v]) / (R_com(s, cp) - R_com(s, c));
      dpdx[v] = avg_falle(slpn, slpp);
    }
    return;
  }
  for (int v = 0; v < nv; v++) {
    double slpn = (Pc[v] - Pn[v]) / dx;
    double slpp = (Pp[v] - Pc[v]) / dx;
    dpdx[v] = avg_falle(slpn, slpp);
  }
}
/* VectorOps_Cart::SetEdgeState (VectorOps.cpp:535-571) */
static void set_edge_state(const pion_oracle *s, long c, int positive, const double *dpdx, double *edge, int OA) {
  int nv = s->nv;
  const double *Pc = s->Ph + c * nv;
  double dx = s->dx;
  if (OA == OA1) {
    for (int v = 0; v < nv; v++) edge[v] = Pc[v];
  } else if (radial_axis(s, s->dir)) {
    /* VectorOps_Cyl::SetEdgeState RPcyl/RNcyl (VectorOps.cpp:1073-1079), VectorOps_Sph (:312-318) */
    double R = dpos(s, c, s->dir);
    double del = positive ? (R + dx * 0.5 - R_com(s, c)) : (R - dx * 0.5 - R_com(s, c));
    for (int v = 0; v < nv; v++) edge[v] = Pc[v] + dpdx[v] * del;
  } else if (positive) {
    for (int v = 0; v < nv; v++) edge[v] = Pc[v] + dpdx[v] * dx * 0.5;
  } else {
    for (int v = 0; v < nv; v++) edge[v] = Pc[v] - dpdx[v] * dx * 0.5;
  }
}
/* VectorOps_Cart::Divergence on Ph velocities (VectorOps.cpp:377-439) */
static double divergence_v(const pion_oracle *s, long c) {
  double divv = 0.0;
  int nv = s->nv;
  for (int v = 0; v < s->ndim; v++) {
    long n = nextpt(s, c, 2 * v), p = nextpt(s, c, 2 * v + 1);
    if (n < 0) n = c;
    if (p < 0) p = c;
    double d = (n == c || p == c) ? s->dx : 2.0 * s->dx;
    if (s->cfg.coord_sys == PO_COORD_CYL && v == 1) { /* VectorOps_Cyl::Divergence (VectorOps.cpp:948-954) */
      double rn = R_com_cyl(s, n), rp = R_com_cyl(s, p);
      divv += 2.0 * (rp * s->Ph[p * nv + VX + v] - rn * s->Ph[n * nv + VX + v]) / (rp * rp - rn * rn);
      continue;
    }
    divv += (s->Ph[p * nv + VX + v] - s->Ph[n * nv + VX + v]) / d;
  }
  return divv;
}
/* VectorOps_Cart::GradZone / CentralDiff on Ph[PG] (VectorOps.cpp:282-368) */
static double grad_zone_p(const pion_oracle *s, long c, int ax) {
  int nv = s->nv;
  long n = nextpt(s, c, 2 * ax), p = nextpt(s, c, 2 * ax + 1);
  if (n < 0) n = c;
  if (p < 0) p = c;
  double min_v = fmin(s->Ph[p * nv + PG], s->Ph[n * nv + PG]);
  return fabs(s->Ph[p * nv + PG] - s->Ph[n * nv + PG]) / min_v;
}

/* ------------------------------------------------------------------ */
/* FV_solver_base::preprocess_data (solver_eqn_base.cpp:353-415)       */
/* ------------------------------------------------------------------ */
/* set_Hcorrection (:579-599): eta = 0.5(|du| + |dc|) from the edge states */
static void set_Hcorrection(pion_oracle *s, long c, int axis, const double *eL, const double *eR) {
  double eta = 0.5 * (fabs(eR[s->eVX] - eL[s->eVX]) + fabs(maxspeed(s, eR) - maxspeed(s, eL)));
  s->hcorr[c * 3 + axis] = eta;
}
/* calc_Hcorrection (:423-573): same column walk as dynamics_dU_column */
static void calc_Hcorrection(pion_oracle *s, int csp) {
  int nv = s->nv;
  double slope_a[PO_MAXVAR], slope_b[PO_MAXVAR], edgeL[PO_MAXVAR], edgeR[PO_MAXVAR];
  for (int idim = 0; idim < s->ndim; idim++) {
    set_direction(s, idim);
    int a1 = (idim + 1) % 3, a2 = (idim + 2) % 3;
    for (int i2 = 0; i2 < s->NGa[a2]; i2++)
      for (int i1 = 0; i1 < s->NGa[a1]; i1++) {
        int ijk[3];
        ijk[idim] = 0; ijk[a1] = i1; ijk[a2] = i2;
        long cpt = cidx(s, ijk[0], ijk[1], ijk[2]);
        long npt = cpt + s->stride[idim];
        double *slope_cpt = slope_a, *slope_npt = slope_b, *tmp;
        for (int v = 0; v < nv; v++) { slope_cpt[v] = 0.; edgeL[v] = 0.; }
        int n = s->NGa[idim];
        for (int i = 0; i < n - 2; i++) {
          set_edge_state(s, cpt, 1, slope_cpt, edgeL, csp);
          set_slope(s, npt, idim, slope_npt, csp);
          set_edge_state(s, npt, 0, slope_npt, edgeR, csp);
          set_Hcorrection(s, cpt, idim, edgeL, edgeR);
          cpt = npt;
          npt += s->stride[idim];
          tmp = slope_cpt; slope_cpt = slope_npt; slope_npt = tmp;
        }
        set_edge_state(s, cpt, 1, slope_cpt, edgeL, csp);
        for (int v = 0; v < nv; v++) slope_npt[v] = 0.;
        set_edge_state(s, npt, 0, slope_npt, edgeR, csp);
        set_Hcorrection(s, cpt, idim, edgeL, edgeR);
      }
  }
  set_direction(s, 0);
}
static void preprocess_data(pion_oracle *s, int csp) {
  int av = s->cfg.artviscosity;
  if (av == PO_AV_HCORR || av == PO_AV_HCORR_FKJ98) calc_Hcorrection(s, csp);
  if (s->cfg.solver == PO_FLUX_HLLD) {
    for (long c = 0; c < s->ncell; c++) {
      s->divv[c] = divergence_v(s, c);
      double gradp = 0.0;
      for (int i = 0; i < s->ndim; i++) gradp += grad_zone_p(s, c, i);
      s->gradp[c] = gradp;
    }
  }
}

/* ------------------------------------------------------------------ */
/* sources, dU, column sweep                                           */
/* ------------------------------------------------------------------ */
/* FV_solver_mhd_ideal_adi::MHDsource (solver_eqn_mhd_adi.cpp:396-443) and the
 * GLM addition (:782-813).  Cell-centre Ph values, scattered into both cells. */
static void mhd_source(pion_oracle *s, long cl, long cr, double dt) {
  if (s->cfg.eqntype == PO_EQEUL) return;
  int nv = s->nv;
  double dx = s->dx;
  const double *L = s->Ph + cl * nv, *R = s->Ph + cr * nv;
  double *dUl = s->dU + cl * nv, *dUr = s->dU + cr * nv;
  double sm = 0.0;
  if (s->cfg.eqntype == PO_EQGLM) sm = 0.5 * (L[SI] + R[SI]);
  double bm = 0.5 * (L[s->eBX] + R[s->eBX]);
  double uB_l = L[s->eBX] * L[s->eVX] + L[s->eBY] * L[s->eVY] + L[s->eBZ] * L[s->eVZ];
  double uB_r = R[s->eBX] * R[s->eVX] + R[s->eBY] * R[s->eVY] + R[s->eBZ] * R[s->eVZ];
  double pl[PO_MAXVAR], pr[PO_MAXVAR];
  for (int v = 0; v < nv; v++) pl[v] = pr[v] = 0.0;
  pl[s->eMX] = L[s->eBX]; pl[s->eMY] = L[s->eBY]; pl[s->eMZ] = L[s->eBZ]; pl[ERG] = uB_l;
  pl[s->eBBX] = L[s->eVX]; pl[s->eBBY] = L[s->eVY]; pl[s->eBBZ] = L[s->eVZ];
  pr[s->eMX] = R[s->eBX]; pr[s->eMY] = R[s->eBY]; pr[s->eMZ] = R[s->eBZ]; pr[ERG] = uB_r;
  pr[s->eBBX] = R[s->eVX]; pr[s->eBBY] = R[s->eVY]; pr[s->eBBZ] = R[s->eVZ];
  if (radial_axis(s, s->dir)) { /* cyl_FV_solver_mhd_ideal_adi::MHDsource case Rcyl (solver_eqn_mhd_adi.cpp:1087-1098) */
    double rp = dpos(s, cl, 1) + dx * 0.5, rn = rp - dx;
    for (int v = 0; v < nv; v++) dUl[v] -= dt * bm * (pl[v]) * 2.0 * rp / (rp * rp - rn * rn);
    rn = rp;
    rp += dx;
    for (int v = 0; v < nv; v++) dUr[v] += dt * bm * (pr[v]) * 2.0 * rn / (rp * rp - rn * rn);
  } else {
    for (int v = 0; v < nv; v++) {
      dUl[v] -= dt * bm * (pl[v]) / dx;
      dUr[v] += dt * bm * (pr[v]) / dx;
    }
  }
  if (s->cfg.eqntype == PO_EQGLM) {
    for (int v = 0; v < nv; v++) pl[v] = pr[v] = 0.0;
    pl[ERG] = L[s->eVX] * L[SI];
    pl[PSI] = L[s->eVX];
    pr[ERG] = R[s->eVX] * R[SI];
    pr[PSI] = R[s->eVX];
    for (int v = 0; v < nv; v++) {
      dUl[v] -= dt * sm * pl[v] / dx;
      dUr[v] += dt * sm * pr[v] / dx;
    }
  }
}
/* dU_Cell: Euler uses the dt argument (solver_eqn_hydro_adi.cpp:342-363), MHD
 * uses FV_dt (solver_eqn_mhd_adi.cpp:368-387); DivStateVectorComponent
 * VectorOps.cpp:624-644.  Cartesian: no geometric source. */
static void dU_cell(pion_oracle *s, long c, const double *fn, const double *fp, const double *dpdx, int OA, double dt) {
  int nv = s->nv;
  double dx = s->dx;
  double mult = (s->cfg.eqntype == PO_EQEUL) ? dt : s->FV_dt;
  double *dU = s->dU + c * nv;
  double u1[PO_MAXVAR];
  if (!radial_axis(s, s->dir)) {
    for (int v = 0; v < nv; v++) u1[v] = (fn[v] - fp[v]) / dx;
  } else if (s->cfg.coord_sys == PO_COORD_CYL) {
    /* VectorOps_Cyl::DivStateVectorComponent Rcyl (VectorOps.cpp:1228-1232) */
    double rp = dpos(s, c, 1) + dx * 0.5, rn = rp - dx;
    for (int v = 0; v < nv; v++) u1[v] = 2.0 * (rn * fn[v] - rp * fp[v]) / (rp * rp - rn * rn);
    /* geometric_source: Euler solver_eqn_hydro_adi.cpp:560-590, MHD solver_eqn_mhd_adi.cpp:1001-1036,
     * GLM :1156-1190 */
    const double *Ph = s->Ph + c * nv;
    double R = dpos(s, c, 1);
    double ptot = Ph[PG], dptot = dpdx[PG];
    if (s->cfg.eqntype != PO_EQEUL) {
      ptot += (Ph[s->eBX] * Ph[s->eBX] + Ph[s->eBY] * Ph[s->eBY] + Ph[s->eBZ] * Ph[s->eBZ]) / 2.;
      dptot = (dpdx[PG] + Ph[s->eBX] * dpdx[s->eBX] + Ph[s->eBY] * dpdx[s->eBY] + Ph[s->eBZ] * dpdx[s->eBZ]);
    }
    if (OA == OA1) u1[s->eMX] += ptot / R;
    else u1[s->eMX] += (ptot + (R - R_com_cyl(s, c)) * dptot) / R;
    if (s->cfg.eqntype == PO_EQGLM) {
      if (OA == OA1) u1[s->eBBX] += s->chyp * Ph[SI] / R;
      else u1[s->eBBX] += s->chyp * (Ph[SI] + (R - R_com_cyl(s, c)) * dpdx[SI]) / R;
    }
  } else {
    /* VectorOps_Sph::DivStateVectorComponent (VectorOps_spherical.cpp:462-467) +
     * sph_FV_solver_Hydro_Euler::geometric_source (solver_eqn_hydro_adi.cpp:648-668) */
    double rc = dpos(s, c, 0);
    double rp = rc + 0.5 * dx, rn = rp - dx;
    rc = (pow(rp, 3.0) - pow(rn, 3.0)) / 3.0;
    for (int v = 0; v < nv; v++) u1[v] = (rn * rn * fn[v] - rp * rp * fp[v]) / rc;
    const double *Ph = s->Ph + c * nv;
    if (OA == OA1) u1[s->eMX] += 2.0 * Ph[PG] / R3_sph(s, c);
    else u1[s->eMX] += 2.0 * ((Ph[PG] - dpdx[PG] * R_com_sph(s, c)) / R3_sph(s, c) + dpdx[PG]);
  }
  for (int v = 0; v < nv; v++) dU[v] += mult * u1[v];
}
/* time_integrator::dynamics_dU_column (time_integrator.cpp:645-873) */
static void dynamics_dU_column(pion_oracle *s, long start, int axis, double dt, int csp) {
  int nv = s->nv;
  double Fa[PO_MAXVAR], Fb[PO_MAXVAR], sa[PO_MAXVAR], sb[PO_MAXVAR], edgeL[PO_MAXVAR], edgeR[PO_MAXVAR];
  double *Fr_prev = Fa, *Fr_this = Fb, *slope_cpt = sa, *slope_npt = sb, *tmp;
  long cpt = start, npt = start + s->stride[axis];
  for (int v = 0; v < nv; v++) { slope_cpt[v] = 0.; slope_npt[v] = 0.; Fr_prev[v] = 0.; Fr_this[v] = 0.; edgeL[v] = 0.; edgeR[v] = 0.; }
  int n = s->NGa[axis];
  for (int i = 0; i < n - 2; i++) {
    set_edge_state(s, cpt, 1, slope_cpt, edgeL, csp);
    set_slope(s, npt, axis, slope_npt, csp);
    set_edge_state(s, npt, 0, slope_npt, edgeR, csp);
    inter_cell_flux(s, cpt, npt, edgeL, edgeR, Fr_this);
    mhd_source(s, cpt, npt, dt);
    dU_cell(s, cpt, Fr_prev, Fr_this, slope_cpt, csp, dt);
    tmp = Fr_prev; Fr_prev = Fr_this; Fr_this = tmp;
    tmp = slope_cpt; slope_cpt = slope_npt; slope_npt = tmp;
    cpt = npt;
    npt += s->stride[axis];
  }
  /* last pair: right cell first order (:805-814) */
  set_edge_state(s, cpt, 1, slope_cpt, edgeL, csp);
  for (int v = 0; v < nv; v++) slope_npt[v] = 0.;
  set_edge_state(s, npt, 0, slope_npt, edgeR, csp);
  inter_cell_flux(s, cpt, npt, edgeL, edgeR, Fr_this);
  mhd_source(s, cpt, npt, dt);
  dU_cell(s, cpt, Fr_prev, Fr_this, slope_cpt, csp, dt);
}
/* time_integrator::set_dynamics_dU (time_integrator.cpp:553-636) */
static void set_dynamics_dU(pion_oracle *s, double dt, int step) {
  int space_ooa = (step == OA1) ? OA1 : OA2;
  for (int i = 0; i < s->ndim; i++) {
    set_direction(s, i);
    int a1 = (i + 1) % 3, a2 = (i + 2) % 3;
    for (int i2 = 0; i2 < s->NGa[a2]; i2++)
      for (int i1 = 0; i1 < s->NGa[a1]; i1++) {
        int ijk[3];
        ijk[i] = 0; ijk[a1] = i1; ijk[a2] = i2;
        dynamics_dU_column(s, cidx(s, ijk[0], ijk[1], ijk[2]), i, dt, space_ooa);
      }
  }
  set_direction(s, 0);
}
/* time_integrator::calc_dynamics_dU (time_integrator.cpp:498-544) */
static int calc_dynamics_dU(pion_oracle *s, double dt, int step) {
  preprocess_data(s, step);
  set_dynamics_dU(s, dt, step);
  return 0;
}

/* ------------------------------------------------------------------ */
/* microphysics source term: mp_only_cooling + adaptive RKCK           */
/* ------------------------------------------------------------------ */
/* mp_only_cooling::Edot_WSS09CIE_heat_cool_metallines (mp_only_cooling.cpp:470-521) */
static double mp_Edot_metallines(const pion_oracle *s, double rho, double T) {
  size_t ihi = s->nT - 1, ilo = 0, imid = 0;
  do {
    imid = ilo + floor((ihi - ilo) / 2.0);
    if (s->tT[imid] < T) ilo = imid;
    else ihi = imid;
  } while (ihi - ilo > 1);
  int iT = ilo;
  double dT = T - s->tT[iT];
  double rho2 = rho * rho, rate = 0.0;
  rate = -(s->t_Cfbdn[iT] + dT * s->s_Cfbdn[iT]) * rho2 * s->inv_Mu2_elec_H;
  rate = fmin(rate, -(s->t_Ccie[iT] + dT * s->s_Ccie[iT]) * rho2 * s->inv_Mu2);
  rate -= (s->t_Crrh[iT] + dT * s->s_Crrh[iT]) * rho2 * s->inv_Mu2_elec_H;
  rate -= (s->t_Cffhe[iT] + dT * s->s_Cffhe[iT]) * rho2 * s->inv_Mu2_elec_H;
  rate += 8.01e-12 * (s->t_rrhp[iT] + dT * s->s_rrhp[iT]) * rho2 * s->inv_Mu2_elec_H;
  return rate;
}
/* Natural cubic spline as the reference gets it from GSL's cspline (tools/interpolate.cpp:59-118): second
 * derivatives c = y''/2 from the symmetric tridiagonal system (Thomas algorithm), evaluation by bisection +
 * the cubic in (x - x_i).  Same arithmetic as oracle/shims/gsl_shim.c, which stands in for GSL in oracle/_ref. */
static void spline_init(int n, const double *xa, const double *ya, double *c) {
  int m = n - 2;
  double *diag = (double *)malloc(m * sizeof(double)), *off = (double *)malloc(m * sizeof(double)),
         *g = (double *)malloc(m * sizeof(double));
  for (int i = 0; i < m; i++) {
    double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
    double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (yd_ip1 / h_ip1 - yd_i / h_i);
  }
  for (int i = 1; i < m; i++) {
    double w = off[i - 1] / diag[i - 1];
    diag[i] -= w * off[i - 1];
    g[i] -= w * g[i - 1];
  }
  c[0] = 0.0;
  c[n - 1] = 0.0;
  if (m > 0) {
    c[m] = g[m - 1] / diag[m - 1];
    for (int i = m - 1; i-- > 0;) c[i + 1] = (g[i] - off[i] * c[i + 2]) / diag[i];
  }
  free(diag); free(off); free(g);
}
static double spline_eval(const pion_oracle *s, double x) {
  int lo = 0, hi = s->ns - 1;
  while (hi > lo + 1) {
    int mid = (lo + hi) / 2;
    if (s->sx[mid] > x) hi = mid;
    else lo = mid;
  }
  double dx = s->sx[lo + 1] - s->sx[lo], dy = s->sy[lo + 1] - s->sy[lo];
  double c_i = s->sc[lo], c_ip1 = s->sc[lo + 1];
  double b_i = dy / dx - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  double d_i = (c_ip1 - c_i) / (3.0 * dx);
  double delx = x - s->sx[lo];
  return s->sy[lo] + delx * (b_i + delx * (c_i + delx * d_i));
}
/* cooling_function_SD93CIE::cooling_rate_SD93CIE (cooling_SD93_cie.cpp:666-704) */
static double mp_rate_SD93CIE(const pion_oracle *s, double T) {
  if (T < 0.0 || !isfinite(T)) return HUGE_VAL;
  double rate = 0.0;
  T = log10(T);
  const double MinTemp = s->sx[0], MaxTemp = s->sx[s->ns - 1];
  if (T > MaxTemp) rate = s->sy[s->ns - 1] + s->s_maxslope * (T - MaxTemp);
  else if (T < MinTemp) rate = s->sy[0] + s->s_minslope * (T - MinTemp);
  else rate = spline_eval(s, T);
  return exp(2.3025850929940459 * rate); /* pconst.ln10(), constants.h:44 */
}
/* mp_only_cooling::Edot (mp_only_cooling.cpp:383-420) and the Edot_* it dispatches to (:427-521);
 * KI02: CoolingFn::CoolingRate with WhichFunction 2 (cooling.cpp:325-399, MinTemp 5 K) */
static double mp_Edot(const pion_oracle *s, double rho, double T) {
  switch (s->cfg.cooling) {
    case 2: {
      const double nH = rho / s->Mu;
      if (T <= 0.0 || isnan(T) || isinf(T)) return -0.0;
      double rate = 0.0;
      if (T > 5.0) rate += nH * nH * (2.0e-19 * exp(-1.184e5 / (T + 1.0e3)) + 2.8e-28 * sqrt(T) * exp(-92.0 / T));
      rate -= nH * 2.0e-26;
      return -rate;
    }
    case 4: return -(rho * rho / s->Mu_elec / s->Mu_ion) * mp_rate_SD93CIE(s, T);
    case 5: return (rho * rho) * (2.733e-21 * exp(-0.782991 * log(T)) / s->Mu_elec / s->Mu - mp_rate_SD93CIE(s, T) / s->Mu_elec / s->Mu_ion);
    case 7: return 2e-26 * rho / s->Mu - (rho * rho / s->Mu / s->Mu) * mp_rate_SD93CIE(s, T);
    case 6: return (rho * rho) * (2.733e-21 * exp(-0.782991 * log(T)) / s->Mu_elec / s->Mu - mp_rate_SD93CIE(s, T) / s->Mu / s->Mu);
    default: return mp_Edot_metallines(s, rho, T);
  }
}
/* mp_only_cooling::dPdt (:227-236) */
static inline double mp_dPdt(const pion_oracle *s, double E) {
  return mp_Edot(s, s->mp_rho, E * (s->mp_gamma - 1.0) * s->Mu_tot_over_kB / s->mp_rho);
}
/* Integrator_Base::Step_RK5CK for one variable (integrator.cpp:285-371) */
static void step_rk5ck(const pion_oracle *s, double p0, double dt, double *pf, double *dp) {
  static const double b21 = 0.2, b31 = 3. / 40., b32 = 9. / 40., b41 = 0.3, b42 = -0.9, b43 = 1.2, b51 = -11. / 54.,
                      b52 = 2.5, b53 = -70. / 27., b54 = 35. / 27., b61 = 1631. / 55296., b62 = 175. / 512.,
                      b63 = 575. / 13824., b64 = 44275. / 110592., b65 = 253. / 4096., c1 = 37. / 378., c3 = 250. / 621.,
                      c4 = 125. / 594., c6 = 512. / 1771.;
  const double dc1 = c1 - 2825. / 27648., dc3 = c3 - 18575. / 48384., dc4 = c4 - 13525. / 55296., dc5 = -277. / 14336.,
               dc6 = c6 - 0.25;
  double k1, k2, k3, k4, k5, k6, ptemp;
  k1 = mp_dPdt(s, p0);
  ptemp = 0.0;
  ptemp += fabs(k1) * dt / (p0 + 1.0e-100);
  if (ptemp < 1.e-6) {
    *pf = p0 + k1 * dt;
    *dp = k1 * dt;
    return;
  }
  k1 *= dt;
  ptemp = p0 + b21 * k1;
  k2 = mp_dPdt(s, ptemp);
  k2 *= dt;
  ptemp = p0 + b31 * k1 + b32 * k2;
  k3 = mp_dPdt(s, ptemp);
  k3 *= dt;
  ptemp = p0 + b41 * k1 + b42 * k2 + b43 * k3;
  k4 = mp_dPdt(s, ptemp);
  k4 *= dt;
  ptemp = p0 + b51 * k1 + b52 * k2 + b53 * k3 + b54 * k4;
  k5 = mp_dPdt(s, ptemp);
  k5 *= dt;
  ptemp = p0 + b61 * k1 + b62 * k2 + b63 * k3 + b64 * k4 + b65 * k5;
  k6 = mp_dPdt(s, ptemp);
  k6 *= dt;
  *pf = p0 + c1 * k1 + c3 * k3 + c4 * k4 + c6 * k6;
  *dp = dc1 * k1 + dc3 * k3 + dc4 * k4 + dc5 * k5 + dc6 * k6;
}
/* Integrator_Base::Stepper_RKCK, BISECTION_STEPPER variant (integrator.cpp:401-530) */
static int stepper_rkck(const pion_oracle *s, double p0, double t0, double htry, double errtol, double *p1, double *hdid,
                        double *hnext) {
  int rval = 0, ct = 0;
  double h = htry, tnew, eps = 1.e-100, maxerr, err = 0.0, ptemp = 0.0;
  if (h < 0) return 1;
  do {
    step_rk5ck(s, p0, h, &ptemp, &err);
    maxerr = 0;
    if (!isfinite(err) || !isfinite(ptemp) || ptemp < 0.0) {
      maxerr = fmax(maxerr, 1000.0);
    } else {
      err /= fabs(ptemp) + eps;
      err = fabs(err / errtol);
      maxerr = fmax(maxerr, err);
    }
    if (maxerr > 1.) h /= 2.0;
    tnew = t0 + h;
    if (tnew == t0) return -2;
    ct++;
  } while (maxerr > 1.0 && ct < 50);
  if (maxerr > 1.0) rval += ct + (int)(fabs(maxerr));
  *hnext = h * 2.0;
  *hdid = h;
  *p1 = ptemp;
  if (isnan(*p1) || isinf(*p1)) { *p1 = -1.e100; rval++; }
  return rval;
}
/* Integrator_Base::Int_Adaptive_RKCK (integrator.cpp:540-606) */
static int int_adaptive_rkck(const pion_oracle *s, double p0, double t0, double dt, double errtol, double *pf, double *tf) {
  double t = t0, p1 = p0, p2 = 0.0;
  *tf = t0 + dt;
  double h = dt, hdid = 0.0, hnext = 0.0;
  int err = 0, ct = 0, ctmax = 25;
  do {
    err += stepper_rkck(s, p1, t, h, errtol, &p2, &hdid, &hnext);
    t += hdid;
    h = fmin(hnext, *tf - t);
    ct++;
    p1 = p2;
  } while (t < *tf && (err == 0) && (ct < ctmax));
  *pf = p1;
  *tf = t;
  return err;
}
/* mp_only_cooling::TimeUpdateMP (mp_only_cooling.cpp:167-221) */
static int mp_TimeUpdateMP(pion_oracle *s, const double *p_in, double *p_out, double dt) {
  s->mp_rho = p_in[RO];
  s->mp_gamma = s->gamma;
  for (int v = 0; v < s->nv; v++) p_out[v] = p_in[v];
  double Eint0 = p_in[PG] / (s->gamma - 1.0);
  double T = p_out[PG] * s->Mu_tot_over_kB / p_out[RO];
  if (T < s->MinT) { mp_set_temp(s, p_out, s->MinT); T = s->MinT; }
  if (T > s->MaxT) { mp_set_temp(s, p_out, s->MaxT); T = s->MaxT; }
  double Eint = Eint0, tout = 0.0;
  int err = int_adaptive_rkck(s, Eint, 0.0, dt, 1.0e-2, &Eint, &tout);
  if (err) return err; /* fatal in the reference (:207) */
  p_out[PG] = Eint * (s->gamma - 1);
  double Tf = p_out[PG] * s->Mu_tot_over_kB / p_out[RO];
  if (Tf > s->MaxT) p_out[PG] *= s->MaxT / Tf;
  else if (Tf < s->MinT) p_out[PG] *= s->MinT / Tf;
  return 0;
}
/* mp_only_cooling::timescales (mp_only_cooling.cpp:333-358) */
static double mp_timescales(const pion_oracle *s, const double *p_in) {
  double Eint = p_in[PG] / (s->gamma - 1.0);
  double T = p_in[PG] * s->Mu_tot_over_kB / p_in[RO];
  double mintime = 1.0e99;
  if (T >= 1.1 * s->MinT) {
    double rate = fmax(fabs(mp_Edot(s, p_in[RO], T)), fabs(mp_Edot(s, p_in[RO], fmax(s->MinT, 0.5 * T))));
    mintime = fmin(mintime, Eint / rate);
  }
  return mintime;
}
/* time_integrator::calc_noRT_microphysics_dU (time_integrator.cpp:438-489) */
static int calc_microphysics_dU(pion_oracle *s, double delt) {
  if (!s->have_mp) return 0;
  int nv = s->nv, err = 0;
  double p[PO_MAXVAR], ui[PO_MAXVAR], uf[PO_MAXVAR];
  for (long c = 0; c < s->ncell; c++) {
    if (!s->isdomain[c]) continue;
    err += mp_TimeUpdateMP(s, s->P + c * nv, p, delt);
    if (s->cfg.eqntype == PO_EQEUL) { euler_PtoU(s, s->P + c * nv, ui); euler_PtoU(s, p, uf); }
    else { mhd_PtoU(s, s->P + c * nv, ui); mhd_PtoU(s, p, uf); }
    for (int v = 0; v < nv; v++) s->dU[c * nv + v] += uf[v] - ui[v];
  }
  return err;
}

/* ------------------------------------------------------------------ */
/* state update                                                        */
/* ------------------------------------------------------------------ */
/* CellAdvanceTime: Euler solver_eqn_hydro_adi.cpp:372-451, MHD
 * solver_eqn_mhd_adi.cpp:452-504, GLM :822-844 (+GLMsource
 * eqns_mhd_adiabatic.cpp:650-660, using FV_dt) */
static void cell_advance_time(pion_oracle *s, const double *Pin, double *dU, double *Pf) {
  int nv = s->nv;
  double u1[PO_MAXVAR], Pint[PO_MAXVAR], corr[PO_MAXVAR], Pcopy[PO_MAXVAR];
  for (int v = 0; v < nv; v++) Pcopy[v] = Pin[v]; /* Pin may alias Pf */
  const double *src = Pcopy;
  if (s->have_mp) {
    mp_sCMA(s, corr, Pcopy);
    for (int t = 0; t < nv; t++) Pint[t] = Pcopy[t] * corr[t];
    src = Pint;
  }
  if (s->cfg.eqntype == PO_EQEUL) euler_PtoU(s, src, u1);
  else mhd_PtoU(s, src, u1);
  for (int v = 0; v < nv; v++) u1[v] += dU[v];
  if (s->cfg.eqntype == PO_EQEUL) euler_UtoP(s, u1, Pf);
  else mhd_UtoP(s, u1, Pf);
  for (int v = 0; v < nv; v++) dU[v] = 0.;
  if (s->have_mp) {
    mp_sCMA(s, corr, Pf);
    for (int t = 0; t < nv; t++) Pf[t] = Pf[t] * corr[t];
  }
  if (s->cfg.eqntype == PO_EQGLM) Pf[SI] *= exp(-s->FV_dt * s->chyp * s->cr);
}
/* time_integrator::grid_update_state_vector (time_integrator.cpp:881-958) */
static int grid_update_state_vector(pion_oracle *s, double dt, int step, int ooa) {
  (void)dt;
  int nv = s->nv;
  for (long c = 0; c < s->ncell; c++) {
    if (!s->isdomain[c]) {
      for (int v = 0; v < nv; v++) s->dU[c * nv + v] = 0.0;
    } else {
      cell_advance_time(s, s->P + c * nv, s->dU + c * nv, s->Ph + c * nv);
    }
    if (s->have_mp) {
      double T = mp_temperature(s, s->Ph + c * nv);
      if (T > s->cfg.max_temperature) mp_set_temp(s, s->Ph + c * nv, s->cfg.max_temperature);
    }
    if (step == ooa)
      for (int v = 0; v < nv; v++) s->P[c * nv + v] = s->Ph[c * nv + v];
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* time step                                                           */
/* ------------------------------------------------------------------ */
/* CellTimeStep: Euler solver_eqn_hydro_adi.cpp:460-500; MHD
 * solver_eqn_mhd_adi.cpp:516-574 (rotation to the axis of smallest |B|) */
static double cell_time_step(pion_oracle *s, const double *P) {
  double temp;
  if (s->cfg.eqntype == PO_EQEUL) {
    temp = 0.0;
    for (int v = 0; v < s->ndim; v++) temp += P[VX + v] * P[VX + v];
    temp = sqrt(temp);
    temp += chydro(s, P);
  } else {
    temp = fabs(P[VX]);
    if (s->ndim > 1) temp = fmax(temp, fabs(P[VY]));
    if (s->ndim > 2) temp = fmax(temp, fabs(P[VZ]));
    double cf = 0.0;
    if (s->ndim == 1) {
      temp += cfast(s, P);
    } else {
      int newdir = 0;
      if (fabs(P[BY]) < fabs(P[BX])) {
        newdir = 1;
        if (fabs(P[BZ]) < fabs(P[BY])) newdir = 2;
      } else if (fabs(P[BZ]) < fabs(P[BX])) newdir = 2;
      /* rotate(u1,XX,newdir) then cfast == cfast with B_x := B_newdir; the
       * other two components only enter through their squares */
      int a = newdir;
      double bx = P[BX + a], by = P[BX + (a + 1) % 3], bz = P[BX + (a + 2) % 3];
      double ch = sqrt(s->gamma * P[PG] / P[RO]);
      double temp1 = ch * ch + (bx * bx + by * by + bz * bz) / P[RO];
      double temp2 = 4. * ch * ch * bx * bx / P[RO];
      temp2 = fmax(MACHINEACCURACY, temp1 * temp1 - temp2);
      cf = sqrt((temp1 + sqrt(temp2)) / 2.);
      temp += cf;
    }
  }
  double fdt = s->dx / temp;
  fdt *= s->cfg.cfl;
  s->FV_dt = fdt;
  return fdt;
}
/* calc_timestep::calc_dynamics_dt (calc_timestep.cpp:271-333) */
static double calc_dynamics_dt(pion_oracle *s) {
  double dt = 1.e100;
  int nv = s->nv;
  for (long c = 0; c < s->ncell; c++) {
    if (!s->isgd[c]) continue;
    if (s->tsflag[c] && !s->iswind[c]) { /* c->timestep && !c->isbd (calc_timestep.cpp:295) */
      double tempdt = cell_time_step(s, s->P + c * nv);
      dt = fmin(dt, tempdt);
    }
  }
  /* first step with stellar winds: limit dt by the wind speed (calc_timestep.cpp:318-323) */
  if (s->timestep == 0)
    for (int v = 0; v < s->cfg.n_wind; v++) dt = fmin(dt, 0.1 * s->cfg.cfl * s->dx / (s->cfg.wind[v].vinf * 1.0e5));
  return dt;
}
/* calc_timestep::calc_microphysics_dt / get_mp_timescales_no_radiation
 * (calc_timestep.cpp:342-463) */
static double calc_microphysics_dt(pion_oracle *s) {
  if (!s->have_mp) return 1.0e99;
  if (s->cfg.mp_timestep_limit == 0) return 1.0e99;
  /* limit 4 = recombination time only: mp_only_cooling::timescales returns 1e99 when tc is
   * false (mp_only_cooling.cpp:341, calc_timestep.cpp:441-443) */
  if (s->cfg.mp_timestep_limit == 4) return 1.0e99;
  double dt = 1.e99;
  int nv = s->nv;
  for (long c = 0; c < s->ncell; c++) {
    if (!s->isgd[c]) continue;
    if (!s->isdomain[c]) continue; /* isbd / internal-boundary cells skipped (:435) */
    double t = mp_timescales(s, s->Ph + c * nv);
    dt = fmin(dt, t);
  }
  return dt;
}
/* calc_timestep::calculate_timestep (calc_timestep.cpp:68-153) with
 * timestep_checking_and_limiting (:219-262) */
static double calculate_timestep(pion_oracle *s) {
  double t_dyn = calc_dynamics_dt(s);
  double t_mp = calc_microphysics_dt(s);
  s->dt = fmin(t_dyn, t_mp);
  double cr = 0.25 / s->dx;
  /* Set_GLM_Speeds (solver_eqn_mhd_adi.cpp:906-921): chyp = CFL*dx/t_dyn */
  if (s->cfg.eqntype == PO_EQGLM) { s->chyp = s->cfg.cfl * s->dx / t_dyn; s->cr = cr; }
  s->dt = fmin(s->dt, 1.3 * s->last_dt);
  if (s->cfg.op_criterion == 1) s->dt = fmin(s->dt, s->next_optime - s->simtime);
  s->dt = fmin(s->dt, s->cfg.finishtime - s->simtime);
  s->FV_dt = s->dt;
  return s->dt;
}

/* ------------------------------------------------------------------ */
/* boundaries                                                          */
/* ------------------------------------------------------------------ */
static void bc_push(bc_list *b, long c, int isedge) {
  b->cell[b->n] = c;
  b->isedge[b->n] = isedge;
  b->npt[b->n] = -1;
  b->n++;
}
/* UniformGrid::SetupBCs (uniform_grid.cpp:1009-1216): list order = increasing
 * cell id; X faces hold interior y,z only, Y faces add the x-ghost corners, Z
 * faces hold whole xy planes. */
static void setup_bc_lists(pion_oracle *s) {
  int nb = s->nbc;
  s->nbcs = 0;
  for (int d = 0; d < 2 * s->ndim; d++) {
    bc_list *b = &s->bcs[s->nbcs++];
    memset(b, 0, sizeof(*b));
    b->type = s->cfg.bc[d];
    b->dir = d;
    int a = d / 2;
    long cap = (long)nb;
    for (int q = 0; q < 3; q++)
      if (q != a) cap *= s->NGa[q];
    b->cell = (long *)malloc(cap * sizeof(long));
    b->npt = (long *)malloc(cap * sizeof(long));
    b->isedge = (int *)malloc(cap * sizeof(int));
    int lo[3], hi[3];
    for (int q = 0; q < 3; q++) { lo[q] = 0; hi[q] = s->NGa[q]; }
    /* restrict the perpendicular extents per face family */
    if (a == 0) {
      for (int q = 1; q < 3; q++) { lo[q] = s->nb[q]; hi[q] = s->NGa[q] - s->nb[q]; }
    } else if (a == 1) {
      lo[2] = s->nb[2]; hi[2] = s->NGa[2] - s->nb[2];
    }
    if (d & 1) { lo[a] = s->NGa[a] - nb; hi[a] = s->NGa[a]; }
    else { lo[a] = 0; hi[a] = nb; }
    for (int k = lo[2]; k < hi[2]; k++)
      for (int j = lo[1]; j < hi[1]; j++)
        for (int i = lo[0]; i < hi[0]; i++) {
          int ijk[3] = {i, j, k};
          int depth = (d & 1) ? (ijk[a] - (s->NGa[a] - nb) + 1) : (nb - ijk[a]);
          bc_push(b, cidx(s, i, j, k), -depth);
        }
  }
}
static void wind_assign(pion_oracle *s);
static void wind_update(pion_oracle *s);
/* walk `n` steps from c in direction dir */
static long walk(const pion_oracle *s, long c, int dir, int n) {
  for (int v = 0; v < n; v++) c = nextpt(s, c, dir);
  return c;
}
/* assign_update_bcs::assign_boundary_data (assign_update_bcs.cpp:40-120) and
 * the BC_assign_* functions of each boundary type */
static int assign_boundary_data(pion_oracle *s) {
  int nv = s->nv;
  for (int ib = 0; ib < s->nbcs; ib++) {
    bc_list *b = &s->bcs[ib];
    int ondir = (b->dir >= 0) ? (b->dir ^ 1) : -1;
    int a = (b->dir >= 0) ? b->dir / 2 : 0;
    switch (b->type) {
      case PO_BC_PERIODIC: /* periodic_boundaries.cpp:20-55 */
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q];
          s->isdomain[c] = 1;
          long t = walk(s, c, ondir, s->cfg.NG[a]);
          for (int v = 0; v < nv; v++) { s->P[c * nv + v] = s->P[t * nv + v]; s->Ph[c * nv + v] = s->P[t * nv + v]; s->dU[c * nv + v] = 0.0; }
          b->npt[q] = t;
        }
        break;
      case PO_BC_OUTFLOW:
      case PO_BC_ONEWAY_OUT: /* outflow_boundaries.cpp:20-76 (GLM_NEGATIVE_BOUNDARY, boundaries.h:21) */
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q];
          long t = walk(s, c, ondir, -b->isedge[q]);
          for (int v = 0; v < nv; v++) { s->P[c * nv + v] = s->P[t * nv + v]; s->Ph[c * nv + v] = s->P[t * nv + v]; }
          b->npt[q] = t;
          s->isdomain[c] = 0;
          if (s->cfg.eqntype == PO_EQGLM) {
            long t2 = t;
            for (int v = b->isedge[q] + 1; v < 0; v++) t2 = nextpt(s, t2, ondir);
            s->P[c * nv + SI] = -s->P[t2 * nv + SI];
            s->Ph[c * nv + SI] = -s->Ph[t2 * nv + SI];
          }
        }
        break;
      case PO_BC_INFLOW: { /* inflow_boundaries.cpp:20-60: refval = P of the LAST cell's source */
        long t = -1;
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q];
          t = walk(s, c, ondir, -b->isedge[q]);
          for (int v = 0; v < nv; v++) { s->P[c * nv + v] = s->P[t * nv + v]; s->Ph[c * nv + v] = s->P[t * nv + v]; s->dU[c * nv + v] = 0.0; }
          s->isdomain[c] = 0;
        }
        for (int v = 0; v < nv; v++) b->refval[v] = s->P[t * nv + v];
      } break;
      case PO_BC_REFLECTING: /* reflecting_boundaries.cpp:20-115 */
        for (int v = 0; v < nv; v++) b->refval[v] = 1.0;
        b->refval[VX + a] = -1.0;
        if (s->cfg.eqntype == PO_EQMHD || s->cfg.eqntype == PO_EQGLM) b->refval[BX + a] = -1.0;
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q];
          long t = walk(s, c, ondir, -b->isedge[q]);
          for (int v = 0; v < nv; v++) {
            s->P[c * nv + v] = s->P[t * nv + v] * b->refval[v];
            s->Ph[c * nv + v] = s->Ph[t * nv + v] * b->refval[v];
            s->dU[c * nv + v] = 0.0;
          }
          b->npt[q] = t;
        }
        break;
      case PO_BC_FIXED: { /* fixed_boundaries.cpp:20-85: refval from the first on-grid source */
        long t = -1;
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q];
          s->isdomain[c] = 0;
          t = walk(s, c, ondir, -b->isedge[q]);
          if (s->isgd[t]) break;
        }
        for (int v = 0; v < nv; v++) b->refval[v] = s->P[t * nv + v];
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q];
          for (int v = 0; v < nv; v++) { s->P[c * nv + v] = b->refval[v]; s->Ph[c * nv + v] = b->refval[v]; s->dU[c * nv + v] = 0.; }
        }
      } break;
      case PO_BC_DMACH: /* double_Mach_ref_boundaries.cpp:20-84 */
        b->refval[RO] = 1.4; b->refval[PG] = 1.0; b->refval[VX] = 0.0; b->refval[VY] = 0.0; b->refval[VZ] = 0.0;
        for (int v = s->ftr; v < nv; v++) b->refval[v] = -1.0;
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q];
          s->isdomain[c] = 0;
          double bpos = 10.0 * s->simtime / sin(M_PI / 3.0) + 1.0 / 6.0 + dpos(s, c, 1) / tan(M_PI / 3.0);
          if (dpos(s, c, 0) <= bpos) {
            double *p = s->P + c * nv, *ph = s->Ph + c * nv;
            p[RO] = 8.0; p[PG] = 116.5; p[VX] = 7.14470958; p[VY] = -4.125; p[VZ] = 0.0;
            for (int v = s->ftr; v < nv; v++) p[v] = 1.0;
            ph[RO] = 8.0; ph[PG] = 116.5; ph[VX] = 7.14470958; ph[VY] = -4.125; ph[VZ] = 0.0;
            for (int v = s->ftr; v < nv; v++) ph[v] = 1.0;
          } else {
            for (int v = 0; v < nv; v++) { s->P[c * nv + v] = b->refval[v]; s->Ph[c * nv + v] = b->refval[v]; }
          }
        }
        break;
      case PO_BC_DMACH2: { /* double_Mach_ref_boundaries.cpp:90-150 */
        b->refval[RO] = 8.0; b->refval[PG] = 116.5; b->refval[VX] = 7.14470958; b->refval[VY] = -4.125; b->refval[VZ] = 0.0;
        for (int v = s->ftr; v < nv; v++) b->refval[v] = 1.0;
        long cap = (long)s->nb[1] * s->NGa[0];
        b->cell = (long *)malloc(cap * sizeof(long));
        b->npt = (long *)malloc(cap * sizeof(long));
        b->isedge = (int *)malloc(cap * sizeof(int));
        b->n = 0;
        long c = cidx(s, s->nb[0], s->nb[1], s->nb[2]); /* FirstPt() */
        do {
          if (dpos(s, c, 0) <= 1. / 6.) {
            long t = c;
            while ((t = nextpt(s, t, YN)) >= 0) {
              for (int v = 0; v < nv; v++) { s->P[t * nv + v] = b->refval[v]; s->Ph[t * nv + v] = b->refval[v]; }
              bc_push(b, t, 0);
            }
          }
        } while ((c = nextpt(s, c, XP)) >= 0 && (dpos(s, c, 0) <= 1. / 6.));
      } break;
      case PO_BC_MPI: /* BCMPI (MCMD_boundaries.cpp:57-236): ghost cells are filled by the halo exchange,
                       * which the caller (tests: torch.distributed/gloo) performs between the seam calls */
        break;
      case PO_BC_STWIND: /* stellar_wind_boundaries.cpp:29-250 */
        wind_assign(s);
        break;
      default:
        fprintf(stderr, "pion_oracle: BC type %d not restated\n", b->type);
        return 1;
    }
  }
  return 0;
}
/* ------------------------------------------------------------------ */
/* stellar wind internal boundary (STWIND), constant sources only      */
/* ------------------------------------------------------------------ */
/* stellar_wind::set_wind_cell_reference_state (stellar_wind_BC.cpp:375-596) for a
 * WINDTYPE_CONSTANT source in Cartesian coordinates; gamma is the literal 5./3. that
 * add_cell passes (:339). */
static void wind_reference_state(const pion_oracle *s, const po_wind_source *w, long c, double dist, double *p) {
  const double kB = 1.38064852e-16, m_p = 1.672621898e-24, Msun = 1.9891e33, year = 3.1558150e7; /* constants.h */
  const double gamma = 5. / 3.;
  const int nd = s->ndim;
  const double Mdot = w->mdot * Msun / year, Vinf = w->vinf * 1.0e5, v_rot = w->vrot * 1.0e5; /* add_source :163-166 */
  for (int v = 0; v < s->nv; v++) p[v] = 0.0;
  int set_rho = 1;
  if (dist < 0.75 * w->radius && nd > 1) { p[RO] = 1.0e-31; p[PG] = 1.0e-31; set_rho = 0; }
  if (nd == 2) { /* ndim==2 && COORD_CRT (:407-413) */
    p[RO] = Mdot / (Vinf * 2.0 * M_PI * dist);
    p[PG] = kB * w->temp / m_p;
    p[PG] *= exp((gamma - 1.0) * log(2.0 * M_PI * w->rstar * Vinf / Mdot));
    p[PG] *= exp((gamma)*log(p[RO]));
  } else if (set_rho) {
    p[RO] = 1.0 / (dist);
    p[RO] *= p[RO];
    p[RO] *= Mdot / (Vinf * 4.0 * M_PI);
    p[PG] = kB * w->temp / m_p;
    p[PG] *= exp((gamma - 1.0) * log(4.0 * M_PI * w->rstar * w->rstar * Vinf / Mdot));
    p[PG] *= exp((gamma)*log(p[RO]));
  }
  double x = dpos(s, c, 0) - w->dpos[0], y = (nd > 1) ? dpos(s, c, 1) - w->dpos[1] : 0.0,
         z = (nd > 2) ? dpos(s, c, 2) - w->dpos[2] : 0.0;
  const double d2 = exp(2 * log(dist)); /* pconst.pow_fast(dist,2) */
  if (nd == 1) {
    p[VX] = Vinf * x / dist; p[VY] = 0.0; p[VZ] = 0.0;
  } else if (nd == 2) {
    p[VX] = Vinf * x / dist;
    p[VY] = Vinf * y / dist;
    p[VZ] = v_rot * w->rstar * y / d2;
  } else {
    p[VX] = Vinf * x / dist;
    p[VY] = Vinf * y / dist;
    p[VZ] = Vinf * z / dist;
    p[VX] += -v_rot * w->rstar * y / d2;
    p[VY] += v_rot * w->rstar * x / d2;
  }
  if (s->cfg.eqntype == PO_EQMHD || s->cfg.eqntype == PO_EQGLM) {
    double B_s = w->bsrf / sqrt(4.0 * M_PI);
    double D_s = w->rstar / dist;
    double D_2 = D_s * D_s;
    double beta_B_sint = (v_rot / Vinf) * B_s * D_s;
    if (nd == 2) {
      p[BX] = B_s * D_2 * fabs(x) / dist;
      p[BY] = B_s * D_2 / dist;
      p[BY] = (x > 0.0) ? y * p[BY] : -y * p[BY];
      beta_B_sint = beta_B_sint * y / dist;
      p[BZ] = (x > 0.0) ? -beta_B_sint : beta_B_sint;
    } else if (nd == 3) {
      p[BX] = B_s * D_2 / dist;
      p[BX] = (z > 0.0) ? x * p[BX] : -x * p[BX];
      p[BY] = B_s * D_2 / dist;
      p[BY] = (z > 0.0) ? y * p[BY] : -y * p[BY];
      p[BZ] = B_s * D_2 * fabs(z) / dist;
      beta_B_sint *= sqrt(x * x + y * y) / dist;
      beta_B_sint = (z > 0.0) ? -beta_B_sint : beta_B_sint;
      p[BX] += -beta_B_sint * y / dist;
      p[BY] += beta_B_sint * x / dist;
    }
  }
  if (s->cfg.eqntype == PO_EQGLM) p[SI] = 0.0;
  for (int v = 0; v < s->ntr; v++) p[s->ftr + v] = w->tr[v];
  /* SET_NEGATIVE_PRESSURE_TO_FIXED_TEMPERATURE (:583-594); Tmin = EP.MinTemperature */
  if (s->have_mp) {
    if (mp_temperature(s, p) < s->cfg.min_temperature) mp_set_temp(s, p, s->cfg.min_temperature);
  } else {
    p[PG] = fmax(p[PG], s->cfg.min_temperature * p[RO] * kB * 0.78625 / m_p);
  }
}
/* BC_assign_STWIND + BC_assign_STWIND_add_cells2src + stellar_wind::add_cell
 * (stellar_wind_boundaries.cpp:29-250, stellar_wind_BC.cpp:247-347) */
static void wind_assign(pion_oracle *s) {
  int nv = s->nv;
  for (int id = 0; id < s->cfg.n_wind; id++) {
    const po_wind_source *w = &s->cfg.wind[id];
    for (long c = 0; c < s->ncell; c++) { /* FirstPt_All .. NextPt_All: ghost cells included */
      double d = 0.0;
      for (int a = 0; a < s->ndim; a++) d += pow(w->dpos[a] - dpos(s, c, a), 2.0);
      d = sqrt(d);
      if (d <= w->radius) {
        s->wind_cell = (long *)realloc(s->wind_cell, (s->wind_n + 1) * sizeof(long));
        s->wind_p = (double *)realloc(s->wind_p, (size_t)(s->wind_n + 1) * nv * sizeof(double));
        s->wind_cell[s->wind_n] = c;
        s->isdomain[c] = 0; /* isbd = true, isdomain = false (:268-269) */
        s->iswind[c] = 1;
        s->tsflag[c] = (d < 0.8 * w->radius) ? 0 : 1; /* c->timestep (:273-276) */
        wind_reference_state(s, w, c, d, s->wind_p + (size_t)s->wind_n * nv);
        s->wind_n++;
      }
    }
  }
}
/* BC_update_STWIND -> stellar_wind::set_cell_values (:642-670): P and Ph, every call */
static void wind_update(pion_oracle *s) {
  int nv = s->nv;
  for (long q = 0; q < s->wind_n; q++) {
    long c = s->wind_cell[q];
    for (int v = 0; v < nv; v++) s->P[c * nv + v] = s->wind_p[q * nv + v];
    for (int v = 0; v < nv; v++) s->Ph[c * nv + v] = s->wind_p[q * nv + v];
  }
}

/* assign_update_bcs::TimeUpdateExternalBCs (assign_update_bcs.cpp:182-246) and
 * the BC_update_* functions; TimeUpdateInternalBCs (:134-176) only acts on
 * STWIND and runs first (time_integrator.cpp:104-107). */
static int time_update_bcs(pion_oracle *s, int cstep, int maxstep) {
  int nv = s->nv;
  if (s->wind_n) wind_update(s); /* TimeUpdateInternalBCs: BC_update_STWIND */
  for (int ib = 0; ib < s->nbcs; ib++) {
    bc_list *b = &s->bcs[ib];
    int ondir = (b->dir >= 0) ? (b->dir ^ 1) : -1;
    int a = (b->dir >= 0) ? b->dir / 2 : 0;
    switch (b->type) {
      case PO_BC_PERIODIC: /* periodic_boundaries.cpp:68-88 */
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q], t = b->npt[q];
          for (int v = 0; v < nv; v++) { s->Ph[c * nv + v] = s->Ph[t * nv + v]; s->dU[c * nv + v] = 0.; }
          if (cstep == maxstep)
            for (int v = 0; v < nv; v++) s->P[c * nv + v] = s->P[t * nv + v];
        }
        break;
      case PO_BC_OUTFLOW: /* outflow_boundaries.cpp:109-160 */
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q], gc = b->npt[q];
          for (int v = 0; v < nv; v++) { s->Ph[c * nv + v] = s->Ph[gc * nv + v]; s->dU[c * nv + v] = 0.; }
          if (cstep == maxstep)
            for (int v = 0; v < nv; v++) s->P[c * nv + v] = s->P[gc * nv + v];
          if (s->cfg.eqntype == PO_EQGLM) {
            for (int v = b->isedge[q] + 1; v < 0; v++) gc = nextpt(s, gc, ondir);
            s->P[c * nv + SI] = -s->P[gc * nv + SI];
            s->Ph[c * nv + SI] = -s->Ph[gc * nv + SI];
          }
        }
        break;
      case PO_BC_ONEWAY_OUT: { /* oneway_out_boundaries.cpp:38-115 */
        int Vnorm = VX + a;
        int norm_sign = (b->dir & 1) ? 1 : -1;
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q], gc = b->npt[q];
          s->isdomain[c] = 0;
          for (int v = 0; v < nv; v++) { s->Ph[c * nv + v] = s->Ph[gc * nv + v]; s->dU[c * nv + v] = 0.; }
          s->Ph[c * nv + Vnorm] = norm_sign * fmax(0.0, s->Ph[c * nv + Vnorm] * norm_sign);
          if (cstep == maxstep) {
            for (int v = 0; v < nv; v++) s->P[c * nv + v] = s->P[gc * nv + v];
            s->P[c * nv + Vnorm] = norm_sign * fmax(0.0, s->P[c * nv + Vnorm] * norm_sign);
          }
          if (s->cfg.eqntype == PO_EQGLM) {
            for (int v = b->isedge[q] + 1; v < 0; v++) gc = nextpt(s, gc, ondir);
            s->P[c * nv + SI] = -s->P[gc * nv + SI];
            s->Ph[c * nv + SI] = -s->Ph[gc * nv + SI];
          }
        }
      } break;
      case PO_BC_INFLOW: /* inflow_boundaries.cpp:83-100 */
      case PO_BC_FIXED:  /* fixed_boundaries.cpp:91-107 */
      case PO_BC_DMACH2: /* double_Mach_ref_boundaries.cpp:214-230 */
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q];
          for (int v = 0; v < nv; v++) { s->dU[c * nv + v] = 0.; s->P[c * nv + v] = b->refval[v]; s->Ph[c * nv + v] = b->refval[v]; }
        }
        break;
      case PO_BC_REFLECTING: /* reflecting_boundaries.cpp:123-145 */
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q], t = b->npt[q];
          for (int v = 0; v < nv; v++) s->Ph[c * nv + v] = s->Ph[t * nv + v] * b->refval[v];
          for (int v = 0; v < nv; v++) s->dU[c * nv + v] = 0.;
          if (cstep == maxstep)
            for (int v = 0; v < nv; v++) s->P[c * nv + v] = s->P[t * nv + v] * b->refval[v];
        }
        break;
      case PO_BC_DMACH: /* double_Mach_ref_boundaries.cpp:169-208; simtime = start of step */
        for (long q = 0; q < b->n; q++) {
          long c = b->cell[q];
          double *ph = s->Ph + c * nv;
          double bpos = 10.0 * s->simtime / sin(M_PI / 3.0) + 1.0 / 6.0 + dpos(s, c, 1) / tan(M_PI / 3.0);
          if (dpos(s, c, 0) <= bpos) {
            ph[RO] = 8.0; ph[PG] = 116.5; ph[VX] = 7.14470958; ph[VY] = -4.125; ph[VZ] = 0.0;
            for (int v = s->ftr; v < nv; v++) ph[v] = 1.0;
          } else {
            for (int v = 0; v < nv; v++) ph[v] = b->refval[v];
          }
          for (int v = 0; v < nv; v++) s->dU[c * nv + v] = 0.0;
          if (cstep == maxstep)
            for (int v = 0; v < nv; v++) s->P[c * nv + v] = ph[v];
        }
        break;
      case PO_BC_MPI:
        break;
      case PO_BC_STWIND: /* updated by TimeUpdateInternalBCs above; skipped here (:238) */
        break;
      default:
        return 1;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* time integration                                                    */
/* ------------------------------------------------------------------ */
/* time_integrator::first_order_update / second_order_update
 * (time_integrator.cpp:151-243) */
static int order_update(pion_oracle *s, double dt, int space, int ooa) {
  s->FV_dt = dt; /* spatial_solver->Setdt(dt) */
  calc_microphysics_dU(s, dt);
  calc_dynamics_dU(s, dt, space);
  grid_update_state_vector(s, dt, space, ooa);
  return 0;
}
/* time_integrator::advance_time (time_integrator.cpp:72-142) */
static double advance_time(pion_oracle *s) {
  if (s->cfg.tmOOA == OA1 && s->cfg.spOOA == OA1) {
    order_update(s, s->dt, OA1, OA1);
    time_update_bcs(s, OA1, OA1);
  } else {
    order_update(s, 0.5 * s->dt, OA1, OA2);
    time_update_bcs(s, OA1, OA2);
    order_update(s, s->dt, OA2, OA2);
    time_update_bcs(s, OA2, OA2);
  }
  s->simtime += s->dt;
  s->last_dt = s->dt;
  s->timestep++;
  return s->dt;
}

/* ------------------------------------------------------------------ */
/* C API                                                               */
/* ------------------------------------------------------------------ */
pion_oracle *po_create(const pion_oracle_config *cfg) {
  pion_oracle *s = (pion_oracle *)calloc(1, sizeof(pion_oracle));
  s->cfg = *cfg;
  /* "Force Nbc=1 if using Lax-Friedrichs flux" (setup_fixed_grid.cpp:188-190) */
  if (cfg->solver == PO_FLUX_LF) s->cfg.spOOA = s->cfg.tmOOA = OA1;
  s->nv = cfg->nvar;
  s->ndim = cfg->ndim;
  s->ntr = cfg->ntracer;
  s->ftr = cfg->nvar - cfg->ntracer;
  s->gamma = cfg->gamma;
  /* setup_fixed_grid::setup_grid (setup_fixed_grid.cpp:183-186) */
  s->nbc = (s->cfg.spOOA == OA2) ? 2 : 1;
  for (int a = 0; a < 3; a++) {
    s->nb[a] = (a < s->ndim) ? s->nbc : 0;
    s->NGa[a] = (a < s->ndim) ? cfg->NG[a] + 2 * s->nbc : 1;
    if (a >= s->ndim) s->cfg.NG[a] = 1;
  }
  s->stride[0] = 1;
  s->stride[1] = s->NGa[0];
  s->stride[2] = (long)s->NGa[0] * s->NGa[1];
  s->ncell = (long)s->NGa[0] * s->NGa[1] * s->NGa[2];
  /* UniformGrid::set_cell_size: G_range[XX]/G_ng[XX] */
  s->dx = (cfg->xmax[0] - cfg->xmin[0]) / cfg->NG[0];
  size_t n = (size_t)s->ncell * s->nv;
  s->P = (double *)calloc(n, sizeof(double));
  s->Ph = (double *)calloc(n, sizeof(double));
  s->dU = (double *)calloc(n, sizeof(double));
  s->hcorr = (double *)calloc((size_t)s->ncell * 3, sizeof(double));
  s->divv = (double *)calloc(s->ncell, sizeof(double));
  s->gradp = (double *)calloc(s->ncell, sizeof(double));
  s->isgd = (unsigned char *)calloc(s->ncell, 1);
  s->isdomain = (unsigned char *)calloc(s->ncell, 1);
  s->tsflag = (unsigned char *)calloc(s->ncell, 1);
  s->iswind = (unsigned char *)calloc(s->ncell, 1);
  for (long c = 0; c < s->ncell; c++) {
    int ijk[3], in = 1;
    cijk(s, c, ijk);
    for (int a = 0; a < s->ndim; a++)
      if (ijk[a] < s->nb[a] || ijk[a] >= s->NGa[a] - s->nb[a]) in = 0;
    s->isgd[c] = in;
    s->isdomain[c] = in; /* uniform_grid.cpp:343-356 */
    s->tsflag[c] = in;
  }
  set_direction(s, 0);
  s->simtime = cfg->starttime;
  s->last_dt = 1.e100; /* sim_params.cpp:53 */
  s->dt = 0.0;
  s->timestep = 0;
  s->next_optime = 0.0;
  setup_bc_lists(s);
  for (int i = 0; i < cfg->n_internal_bc; i++) {
    bc_list *b = &s->bcs[s->nbcs++];
    memset(b, 0, sizeof(*b));
    b->type = cfg->internal_bc[i];
    b->dir = -1;
  }
  if (cfg->cooling) {
    /* mp_only_cooling constructor (mp_only_cooling.cpp:96-160); m_p, k_B from constants.h */
    const double m_p = 1.672621898e-24, kB = 1.38064852e-16; /* constants.h:53,64 */
    s->have_mp = 1;
    s->Mu = 1.40 * m_p;
    double Mu_tot = 0.609 * m_p;
    s->Mu_tot_over_kB = Mu_tot / kB;
    s->Mu_elec = 1.167 * m_p;
    s->Mu_ion = 1.273 * m_p;
    s->inv_Mu2 = 1.0 / (s->Mu * s->Mu);
    s->inv_Mu2_elec_H = 1.0 / (s->Mu_elec * s->Mu);
    s->MaxT = cfg->max_temperature;
    s->MinT = cfg->min_temperature;
    if (s->MinT < 1.0 || s->MinT > 1.0e6) s->MinT = 1.0;
    if (s->MaxT < 1.0e2 || s->MaxT > 3.0e10) s->MaxT = 1.0e8;
    if (cfg->cooling >= 4 && cfg->cooling <= 7) {
      s->ns = cfg->n_spline;
      s->sx = (double *)malloc(s->ns * sizeof(double));
      s->sy = (double *)malloc(s->ns * sizeof(double));
      s->sc = (double *)calloc(s->ns, sizeof(double));
      memcpy(s->sx, cfg->spline_logT, s->ns * sizeof(double));
      memcpy(s->sy, cfg->spline_logL, s->ns * sizeof(double));
      s->s_minslope = cfg->spline_min_slope;
      s->s_maxslope = cfg->spline_max_slope;
      spline_init(s->ns, s->sx, s->sy, s->sc);
    }
    int nT = (cfg->cooling == 8) ? cfg->n_table : 0;
    s->nT = nT;
    double **dst[6] = {&s->tT, &s->t_rrhp, &s->t_Crrh, &s->t_Cffhe, &s->t_Cfbdn, &s->t_Ccie};
    const double *src[6] = {cfg->table_T, cfg->table_rrhp, cfg->table_C_rrh, cfg->table_C_ffhe, cfg->table_C_fbdn, cfg->table_C_cie};
    for (int q = 0; q < 6; q++) {
      *dst[q] = (double *)malloc(nT * sizeof(double));
      memcpy(*dst[q], src[q], nT * sizeof(double));
    }
    double **sl[5] = {&s->s_rrhp, &s->s_Crrh, &s->s_Cffhe, &s->s_Cfbdn, &s->s_Ccie};
    double *tb[5] = {s->t_rrhp, s->t_Crrh, s->t_Cffhe, s->t_Cfbdn, s->t_Ccie};
    for (int q = 0; q < 5; q++) {
      *sl[q] = (double *)calloc(nT, sizeof(double));
      /* gen_mpoc_lookup_tables slopes (mp_only_cooling.cpp:566-579) */
      for (int i = 0; i < nT - 1; i++) (*sl[q])[i] = (tb[q][i + 1] - tb[q][i]) / (s->tT[i + 1] - s->tT[i]);
    }
  }
  return s;
}

void po_destroy(pion_oracle *s) {
  if (!s) return;
  free(s->P); free(s->Ph); free(s->dU); free(s->hcorr); free(s->divv); free(s->gradp);
  free(s->isgd); free(s->isdomain); free(s->tsflag); free(s->iswind); free(s->wind_cell); free(s->wind_p);
  for (int i = 0; i < s->nbcs; i++) { free(s->bcs[i].cell); free(s->bcs[i].npt); free(s->bcs[i].isedge); }
  free(s->tT); free(s->t_rrhp); free(s->t_Crrh); free(s->t_Cffhe); free(s->t_Cfbdn); free(s->t_Ccie);
  free(s->s_rrhp); free(s->s_Crrh); free(s->s_Cffhe); free(s->s_Cfbdn); free(s->s_Ccie);
  free(s->sx); free(s->sy); free(s->sc);
  free(s);
}

int po_info(pion_oracle *s, int *info, double *dinfo) {
  for (int a = 0; a < 3; a++) { info[a] = s->NGa[a]; info[3 + a] = s->cfg.NG[a]; }
  info[6] = s->nv; info[7] = s->ndim; info[8] = s->nbc; info[9] = s->cfg.eqntype; info[10] = s->cfg.solver;
  info[11] = s->cfg.artviscosity; info[12] = s->ntr; info[13] = s->cfg.coord_sys; info[14] = s->timestep;
  info[15] = s->cfg.tmOOA; info[16] = s->cfg.spOOA;
  dinfo[0] = s->dx; dinfo[1] = s->gamma; dinfo[2] = s->cfg.cfl; dinfo[3] = s->cfg.etav; dinfo[4] = s->simtime;
  dinfo[5] = s->dt; dinfo[6] = s->last_dt; dinfo[7] = s->cfg.finishtime;
  for (int a = 0; a < 3; a++) { dinfo[8 + a] = s->cfg.xmin[a]; dinfo[11 + a] = s->cfg.xmax[a]; }
  dinfo[14] = s->cfg.min_temperature; dinfo[15] = s->cfg.max_temperature; dinfo[16] = s->cfg.starttime;
  dinfo[17] = s->chyp; dinfo[18] = s->cr;
  return 0;
}
static double *which_arr(pion_oracle *s, int which) { return which == 0 ? s->P : which == 1 ? s->Ph : s->dU; }
int po_get_state(pion_oracle *s, int which, double *out) {
  const double *a = which_arr(s, which);
  for (long c = 0; c < s->ncell; c++)
    for (int v = 0; v < s->nv; v++) out[v * s->ncell + c] = a[c * s->nv + v];
  return 0;
}
int po_set_state(pion_oracle *s, int which, const double *in) {
  double *a = which_arr(s, which);
  for (long c = 0; c < s->ncell; c++)
    for (int v = 0; v < s->nv; v++) a[c * s->nv + v] = in[v * s->ncell + c];
  return 0;
}
int po_get_flags(pion_oracle *s, int *out) {
  for (long c = 0; c < s->ncell; c++)
    out[c] = (s->isgd[c] ? 1 : 0) | ((!s->isgd[c] || s->iswind[c]) ? 2 : 0) | (s->isdomain[c] ? 4 : 0) | 8 | (s->tsflag[c] ? 16 : 0);
  return 0;
}
int po_get_extra(pion_oracle *s, int what, int axis, double *out) {
  for (long c = 0; c < s->ncell; c++)
    out[c] = (what == 0) ? s->divv[c] : (what == 1) ? s->gradp[c] : s->hcorr[c * 3 + axis];
  return 0;
}
/* sim_init::Init after ReadData (sim_init.cpp:215-280) */
int po_init_after_state(pion_oracle *s) {
  int nv = s->nv;
  for (long c = 0; c < s->ncell; c++) {
    if (!s->isgd[c]) continue;
    for (int v = 0; v < nv; v++) s->Ph[c * nv + v] = s->P[c * nv + v];
    if (s->cfg.eqntype == PO_EQGLM && s->timestep == 0) s->P[c * nv + SI] = s->Ph[c * nv + SI] = 0.;
  }
  int err = assign_boundary_data(s);
  err += time_update_bcs(s, s->cfg.tmOOA, s->cfg.tmOOA);
  if (s->cfg.op_criterion == 1) {
    s->next_optime = s->simtime + s->cfg.opfreq_time;
    double tmp = ((s->simtime / s->cfg.opfreq_time) - floor(s->simtime / s->cfg.opfreq_time)) * s->cfg.opfreq_time;
    s->next_optime -= tmp;
  }
  return err;
}
double po_calc_timestep(pion_oracle *s) { return calculate_timestep(s); }
double po_advance(pion_oracle *s) { return advance_time(s); }
double po_dynamics_dt(pion_oracle *s) { return calc_dynamics_dt(s); }
double po_microphysics_dt(pion_oracle *s) { return calc_microphysics_dt(s); }
/* constants::equalD (constants.cpp:48-69) */
static int po_equalD(double a, double b) {
  if (a == b) return 1;
  if (fabs(a) + fabs(b) < 1.0e-100) return 1;
  return (fabs(a - b) / (fabs(a) + fabs(b) + 1.0e-100)) < 1.0e-12;
}
/* the output-criterion bookkeeping of sim_init::output_data (sim_init.cpp:733-742), called by
 * sim_control::Time_Int after every step (sim_control.cpp:252): with op_criterion == 1 an output time that
 * has been reached is consumed, next_optime += opfreq_time */
static void output_bookkeeping(pion_oracle *s) {
  if (s->cfg.op_criterion != 1 || s->timestep == 0) return;
  const int maxtime = s->simtime >= s->cfg.finishtime;
  if (po_equalD(s->simtime, s->next_optime) || maxtime) s->next_optime += s->cfg.opfreq_time;
}
int po_run(pion_oracle *s, int nsteps, double *dts) {
  for (int i = 0; i < nsteps; i++) {
    double dt = calculate_timestep(s);
    if (!(dt > 0)) return i;
    advance_time(s);
    if (dts) dts[i] = dt;
    output_bookkeeping(s);
  }
  return nsteps;
}
int po_update_bcs(pion_oracle *s, int cstep, int maxstep) { return time_update_bcs(s, cstep, maxstep); }
int po_dynamics_dU(pion_oracle *s, double dt, int step) {
  s->FV_dt = dt;
  return calc_dynamics_dU(s, dt, step);
}
int po_microphysics_dU(pion_oracle *s, double dt) { return calc_microphysics_dU(s, dt); }
int po_update_state(pion_oracle *s, double dt, int step, int ooa) {
  s->FV_dt = dt;
  return grid_update_state_vector(s, dt, step, ooa);
}
void po_set_dt(pion_oracle *s, double dt) { s->dt = dt; s->FV_dt = dt; }
void po_set_glm_speeds(pion_oracle *s, double tdyn, double dx, double cr) {
  s->chyp = s->cfg.cfl * dx / tdyn;
  s->cr = cr;
}
void po_set_time(pion_oracle *s, double simtime, double last_dt, int timestep) {
  s->simtime = simtime;
  s->last_dt = last_dt;
  s->timestep = timestep;
}
int po_intercell_flux(pion_oracle *s, int axis, const double *Pl, const double *Pr, double divv_l, double gradp_l,
                      double divv_r, double gradp_r, double hc_etamax, double *flux) {
  long cl = cidx(s, s->nb[0], s->nb[1], s->nb[2]), cr = cl + 1;
  double sv[4] = {s->divv[cl], s->gradp[cl], s->divv[cr], s->gradp[cr]};
  s->divv[cl] = divv_l; s->gradp[cl] = gradp_l; s->divv[cr] = divv_r; s->gradp[cr] = gradp_r;
  int av = s->cfg.artviscosity;
  set_direction(s, axis);
  if (av == PO_AV_HCORR || av == PO_AV_HCORR_FKJ98) {
    /* caller-supplied eta: bypass select_Hcorr_eta */
    double pstar[PO_MAXVAR];
    s->HC_etamax = hc_etamax;
    s->cfg.artviscosity = (av == PO_AV_HCORR) ? PO_AV_NONE : PO_AV_FKJ98;
    inter_cell_flux(s, cl, cr, Pl, Pr, flux);
    s->cfg.artviscosity = av;
    (void)pstar;
  } else {
    s->HC_etamax = 0.0;
    inter_cell_flux(s, cl, cr, Pl, Pr, flux);
  }
  set_direction(s, 0);
  s->divv[cl] = sv[0]; s->gradp[cl] = sv[1]; s->divv[cr] = sv[2]; s->gradp[cr] = sv[3];
  return 0;
}
int po_error_counts(pion_oracle *s, long *out) {
  out[0] = s->neg_rho;
  out[1] = s->neg_pg;
  return 0;
}
