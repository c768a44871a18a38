/* TEST INFRASTRUCTURE ONLY -- the parity oracle, never linked into the product.
 *
 * pion_oracle.h: C API of the plain-C restatement of the reference's
 * finite-volume dynamics update (see pion_oracle.c for file:line citations).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Parity pinning: pion_oracle is checked bit-for-bit (tests/test_oracle_vs_ref.py)
 * against oracle/_ref/libpion_ref.so, i.e. the UNMODIFIED reference translation
 * units compiled here, and against golden vectors generated from that library
 * (tests/golden/, generating script tests/golden/make_golden.py).
 */
#ifndef PION_ORACLE_H
#define PION_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define PO_MAXVAR 16

/* integer codes are the reference's own (source/constants.h:166-246,
 * source/boundaries/boundaries.h:32-52) */
enum { PO_EQEUL = 1, PO_EQMHD = 2, PO_EQGLM = 3 };
enum { PO_COORD_CRT = 1, PO_COORD_CYL = 2, PO_COORD_SPH = 3 };
enum { PO_FLUX_LF = 0, PO_FLUX_ROE = 4, PO_FLUX_ROE_PV = 5, PO_FLUX_FVS = 6, PO_FLUX_HLLD = 7, PO_FLUX_HLL = 8 };  /* 5, 6: Euler only */
enum { PO_AV_NONE = 0, PO_AV_FKJ98 = 1, PO_AV_HCORR = 3, PO_AV_HCORR_FKJ98 = 4 };
enum {
  PO_BC_PERIODIC = 1, PO_BC_OUTFLOW = 2, PO_BC_INFLOW = 3, PO_BC_REFLECTING = 4,
  PO_BC_FIXED = 5, PO_BC_DMACH = 8, PO_BC_DMACH2 = 9, PO_BC_MPI = 10, PO_BC_ONEWAY_OUT = 13, PO_BC_STWIND = 14
};

/* one constant stellar-wind source: arguments of stellar_wind::add_source
 * (grid/stellar_wind_BC.cpp:125-140), units as in the parameter file */
typedef struct po_wind_source {
  double dpos[3];   /* cm */
  double radius;    /* cm */
  double mdot;      /* Msun/yr */
  double vinf, vrot;/* km/s */
  double temp;      /* K */
  double rstar;     /* cm */
  double bsrf;      /* Gauss */
  double tr[4];     /* tracer values */
} po_wind_source;

typedef struct pion_oracle_config {
  int ndim;
  int NG[3];
  int nvar;
  int ntracer;
  int eqntype;
  int coord_sys;
  int solver;
  int artviscosity;
  int spOOA, tmOOA;
  double gamma, cfl, etav;
  double xmin[3], xmax[3];
  int bc[6];          /* XN,XP,YN,YP,ZN,ZP */
  int n_internal_bc;  /* e.g. {PO_BC_DMACH2} */
  int internal_bc[4];
  double refvec[PO_MAXVAR];
  double starttime, finishtime;
  int op_criterion;   /* 1: limit dt to hit opfreq_time multiples (sim_init.cpp:270) */
  double opfreq_time;
  /* microphysics: mp_only_cooling (cooling>0 && no chemistry) */
  int cooling;        /* EP.cooling flag of mp_only_cooling.cpp:42-48: 0 none, 2 KI02, 4 SD93_CIE, 5 SD93_PLUS_HEATING,
                         6 WSS09_CIE_PLUS_HEATING, 7 WSS09_CIE_ONLY_COOLING, 8 WSS09_CIE_LINE_HEAT_COOL */
  int mp_timestep_limit;
  double min_temperature, max_temperature;
  int n_table;        /* 200 */
  const double *table_T, *table_rrhp, *table_C_rrh, *table_C_ffhe, *table_C_fbdn, *table_C_cie;
  /* internal boundary PO_BC_STWIND: constant wind sources (SWP, sim_params.h) */
  int n_wind;
  po_wind_source wind[2];
  /* EP.cooling 4..7: knots of the cooling-curve spline of cooling_function_SD93CIE (log10 T, log10 Lambda:
   * Tarray / Larray after setup_SD93_cie() [4, 5] or setup_WSS09_CIE() [6, 7]) and its power-law slopes
   * outside the table (cooling_SD93_cie.cpp:87-200,555-660) */
  int n_spline;
  const double *spline_logT, *spline_logL;
  double spline_min_slope, spline_max_slope;
} pion_oracle_config;

typedef struct pion_oracle pion_oracle;

pion_oracle *po_create(const pion_oracle_config *cfg);
void po_destroy(pion_oracle *s);
/* info as pref_info in ref_driver.cpp */
int po_info(pion_oracle *s, int *info, double *dinfo);
int po_get_state(pion_oracle *s, int which, double *out);       /* SoA padded */
int po_set_state(pion_oracle *s, int which, const double *in);
int po_get_flags(pion_oracle *s, int *out);
int po_get_extra(pion_oracle *s, int what, int axis, double *out);
int po_init_after_state(pion_oracle *s);
double po_calc_timestep(pion_oracle *s);
double po_advance(pion_oracle *s);
double po_dynamics_dt(pion_oracle *s);
double po_microphysics_dt(pion_oracle *s);
int po_run(pion_oracle *s, int nsteps, double *dts);
int po_update_bcs(pion_oracle *s, int cstep, int maxstep);
int po_dynamics_dU(pion_oracle *s, double dt, int step);
int po_microphysics_dU(pion_oracle *s, double dt);
int po_update_state(pion_oracle *s, double dt, int step, int ooa);
void po_set_dt(pion_oracle *s, double dt);
void po_set_glm_speeds(pion_oracle *s, double tdyn, double dx, double cr);
void po_set_time(pion_oracle *s, double simtime, double last_dt, int timestep);
int po_intercell_flux(pion_oracle *s, int axis, const double *Pl, const double *Pr, double divv_l,
                      double gradp_l, double divv_r, double gradp_r, double hc_etamax, double *flux);
/* error counters: [0]=negative density events (fatal in the reference),
 * [1]=negative pressure fix-ups */
int po_error_counts(pion_oracle *s, long *out);

#ifdef __cplusplus
}
#endif
#endif
