// TEST INFRASTRUCTURE ONLY -- not product code, never linked into libpion_b200.
//
// oracle/ref_driver.cpp: a thin C-ABI harness around the UNMODIFIED reference
// translation units (compiled where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/libpion_ref.so).  It contains no physics:
// every number it returns is produced by the reference's own classes.
//
// It mirrors what the reference's two mains do, minus file I/O:
//   * parameter parsing + grid/BC/IC set-up as in source/ics/icgen.cpp:90-330
//   * Ph=P, psi=0, boundary assignment and first BC update as in
//     source/sim_control/sim_init.cpp:173-326 (sim_init::Init)
//   * the time loop body of source/sim_control/sim_control.cpp:220-266
//     (calculate_timestep + advance_time)
// and additionally exposes the grid-level seam-2 methods one at a time
// (calc_dynamics_dU, grid_update_state_vector, TimeUpdate*BCs,
// calc_microphysics_dU) so that tests can check each CUDA kernel separately.
//
// State is exchanged as structure-of-arrays doubles [var][k][j][i] over the
// full padded grid (ghost cells included), x fastest.

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

// The knots of the reference's cooling-curve spline are private data members of cooling_function_SD93CIE and
// the class has no accessor; the tree is read-only here, so the harness opens this ONE header up to read them
// (pref_cooling_spline below).  All standard headers are included above, nothing else is affected.
#define private public
#include "microphysics/cooling_SD93_cie.h"
#undef private

#include "defines/functionality_flags.h"
#include "defines/testing_flags.h"
#include "sim_constants.h"
#include "tools/reporting.h"
#include "tools/mem_manage.h"
#include "grid/grid_base_class.h"
#include "grid/cell_interface.h"
#include "sim_control/sim_control.h"
#include "ics/icgen_base.h"
#include "ics/icgen.h"
#include "ics/get_sim_info.h"
#include "dataIO/readparams.h"
#include "dataIO/dataio_text.h"
#include "microphysics/microphysics_base.h"
#include "microphysics/mp_only_cooling.h"
#include "spatial_solvers/solver_eqn_base.h"

using namespace std;

namespace {

struct NullBuf : public std::streambuf {
  int overflow(int c) override { return c; }
};
NullBuf g_nullbuf;
std::streambuf *g_cout_saved = 0;

void quiet_on() {
  if (getenv("PION_REF_VERBOSE")) return;
  if (!g_cout_saved) g_cout_saved = std::cout.rdbuf(&g_nullbuf);
}

// sim_control with its protected seam-2 methods made reachable.
class RefSim : public sim_control {
 public:
  vector<class GridBaseClass *> grid;
  class ReadParams *rp = 0;
  class ICsetup_base *ic = 0;
  int NGa[3] = {1, 1, 1};  // padded extents
  int ioff[3] = {0, 0, 0};

  RefSim() : sim_control() {}
  ~RefSim() {
    if (rp) delete rp;
    if (ic) delete ic;
  }

  GridBaseClass *g() { return grid[0]; }
  FV_solver_base *solver() { return spatial_solver; }

  int setup(const char *pfile, int run_ics) {
    int err = 0;
    MP = 0;
    // the reference keeps its wind-source list in a process-wide global that read_gridparams
    // appends to (ics/get_sim_info.cpp:875): start every instance from an empty list
    SWP.params.clear();
    SWP.Nsources = 0;
    {
      class get_sim_info siminfo;
      err += siminfo.read_gridparams(pfile, SimPM);
      if (err) return err;
    }
    // single-level bookkeeping, as icgen.cpp:131-146
    SimPM.levels.clear();
    SimPM.levels.resize(1);
    SimPM.grid_nlevels = 1;
    SimPM.levels[0].parent = 0;
    SimPM.levels[0].child = 0;
    SimPM.levels[0].Ncell = SimPM.Ncell;
    for (int v = 0; v < MAX_DIM; v++) SimPM.levels[0].NG[v] = SimPM.NG[v];
    for (int v = 0; v < MAX_DIM; v++) SimPM.levels[0].Range[v] = SimPM.Range[v];
    for (int v = 0; v < MAX_DIM; v++) SimPM.levels[0].Xmin[v] = SimPM.Xmin[v];
    for (int v = 0; v < MAX_DIM; v++) SimPM.levels[0].Xmax[v] = SimPM.Xmax[v];
    SimPM.levels[0].dx = SimPM.Range[XX] / SimPM.NG[XX];
    SimPM.levels[0].simtime = SimPM.simtime;
    SimPM.levels[0].dt = 0.0;
    SimPM.levels[0].multiplier = 1;

    grid.resize(1);
    grid[0] = 0;
    err += setup_grid(grid, SimPM);
    SimPM.dx = grid[0]->DX();
    SimPM.levels[0].grid = grid[0];
    err += set_equations(SimPM);
    spatial_solver->SetEOS(SimPM.gamma);
    err += setup_microphysics(SimPM);
    err += boundary_conditions(SimPM, grid);
    err += setup_raytracing(SimPM, grid[0]);
    err += setup_evolving_RT_sources(SimPM);
    err += update_evolving_RT_sources(SimPM, SimPM.simtime, grid[0]->RT);

    for (int a = 0; a < SimPM.ndim; a++) {
      NGa[a] = grid[0]->NG_All(static_cast<axes>(a));
      ioff[a] = grid[0]->iXmin_all(static_cast<axes>(a));
    }

    rp = new ReadParams;
    err += rp->read_paramfile(pfile);
    if (run_ics) {
      string seek = "ics";
      string ics = rp->find_parameter(seek);
      setup_ics_type(ics, &ic);
      ic->set_SimPM(&SimPM);
      err += ic->setup_data(rp, grid[0]);
    }
    return err;
  }

  // sim_init.cpp:215-262 (after ReadData)
  int init_after_state() {
    int err = 0;
    cell *c = grid[0]->FirstPt();
    do {
      for (int v = 0; v < SimPM.nvar; v++) c->Ph[v] = c->P[v];
    } while ((c = grid[0]->NextPt(c)) != 0);
    if (SimPM.eqntype == EQGLM && SimPM.timestep == 0) {
      c = grid[0]->FirstPt();
      do {
        c->P[SI] = c->Ph[SI] = 0.;
      } while ((c = grid[0]->NextPt(c)) != 0);
    }
    err += assign_boundary_data(SimPM, 0, grid[0]);
    err += TimeUpdateInternalBCs(SimPM, 0, grid[0], spatial_solver, SimPM.simtime, SimPM.tmOOA, SimPM.tmOOA);
    err += TimeUpdateExternalBCs(SimPM, 0, grid[0], spatial_solver, SimPM.simtime, SimPM.tmOOA, SimPM.tmOOA);
    // sim_init.cpp:270-280: next output time when outputting by sim-time
    if (SimPM.op_criterion == 1) {
      SimPM.next_optime = SimPM.simtime + SimPM.opfreq_time;
      double tmp = ((SimPM.simtime / SimPM.opfreq_time) - floor(SimPM.simtime / SimPM.opfreq_time)) * SimPM.opfreq_time;
      SimPM.next_optime -= tmp;
    }
    return err;
  }

  long index_of(const cell *c) const {
    long idx[3] = {0, 0, 0};
    for (int a = 0; a < SimPM.ndim; a++) idx[a] = (c->pos[a] - ioff[a]) / 2;
    return idx[0] + NGa[0] * (idx[1] + (long)NGa[1] * idx[2]);
  }
  long ncell_all() const { return (long)NGa[0] * NGa[1] * NGa[2]; }

  // sim_control.cpp:237-239
  double do_calc_timestep() {
    SimPM.levels[0].last_dt = SimPM.last_dt;
    int err = calculate_timestep(SimPM, grid[0], spatial_solver, 0);
    if (err) return -1.0;
    return SimPM.dt;
  }
  double do_advance() { return advance_time(0, grid[0]); }
  // Time_Int calls output_data after every step (sim_control.cpp:252).  The harness has no dataio object, so
  // it cannot call the reference's output_data itself; this is that function's op_criterion == 1 branch
  // (sim_init.cpp:733-742) with the reference's own equalD: an output time that has been reached is consumed.
  void output_bookkeeping() {
    if (SimPM.op_criterion != 1 || SimPM.timestep == 0) return;
    const bool maxtime = SimPM.simtime >= SimPM.finishtime;
    if (pconst.equalD(SimPM.simtime, SimPM.next_optime) || maxtime) SimPM.next_optime += SimPM.opfreq_time;
  }

  int do_update_bcs(int cstep, int maxstep) {
    int err = 0;
    err += TimeUpdateInternalBCs(SimPM, 0, grid[0], spatial_solver, SimPM.simtime, cstep, maxstep);
    err += TimeUpdateExternalBCs(SimPM, 0, grid[0], spatial_solver, SimPM.simtime, cstep, maxstep);
    return err;
  }
  int do_dynamics_dU(double dt, int step) {
    spatial_solver->Setdt(dt);
    return calc_dynamics_dU(dt, step, grid[0]);
  }
  int do_microphysics_dU(double dt) { return calc_microphysics_dU(dt, grid[0]); }
  int do_update_state(double dt, int step, int ooa) {
    spatial_solver->Setdt(dt);
    return grid_update_state_vector(dt, step, ooa, grid[0]);
  }
  double do_dynamics_dt() { return calc_dynamics_dt(SimPM, grid[0], spatial_solver); }
  double do_microphysics_dt() { return calc_microphysics_dt(SimPM, grid[0], 0); }
};

}  // namespace

extern "C" {

void *pref_create(const char *paramfile, int run_ics) {
  quiet_on();
  RefSim *s = new RefSim();
  int err = s->setup(paramfile, run_ics);
  if (err) {
    fprintf(stderr, "pref_create: reference set-up returned %d\n", err);
    delete s;
    return 0;
  }
  return s;
}

void pref_destroy(void *h) {
  RefSim *s = static_cast<RefSim *>(h);
  if (!s) return;
  GridBaseClass *g = s->grid.size() ? s->grid[0] : 0;
  delete s;
  if (g) delete g;
}

// info[0..2]=padded NG, [3..5]=interior NG, [6]=nvar, [7]=ndim, [8]=Nbc,
// [9]=eqntype, [10]=solverType, [11]=artviscosity, [12]=ntracer, [13]=coord_sys,
// [14]=timestep, [15]=tmOOA, [16]=spOOA
int pref_info(void *h, int *info, double *dinfo) {
  RefSim *s = static_cast<RefSim *>(h);
  for (int a = 0; a < 3; a++) {
    info[a] = s->NGa[a];
    info[3 + a] = (a < s->SimPM.ndim) ? s->SimPM.NG[a] : 1;
  }
  info[6] = s->SimPM.nvar;
  info[7] = s->SimPM.ndim;
  info[8] = s->SimPM.Nbc;
  info[9] = s->SimPM.eqntype;
  info[10] = s->SimPM.solverType;
  info[11] = s->SimPM.artviscosity;
  info[12] = s->SimPM.ntracer;
  info[13] = s->SimPM.coord_sys;
  info[14] = s->SimPM.timestep;
  info[15] = s->SimPM.tmOOA;
  info[16] = s->SimPM.spOOA;
  // dinfo: [0]=dx [1]=gamma [2]=CFL [3]=etav [4]=simtime [5]=dt [6]=last_dt
  //        [7]=finishtime [8..10]=Xmin [11..13]=Xmax [14]=MinTemperature
  //        [15]=MaxTemperature [16]=starttime
  dinfo[0] = s->SimPM.dx;
  dinfo[1] = s->SimPM.gamma;
  dinfo[2] = s->SimPM.CFL;
  dinfo[3] = s->SimPM.etav;
  dinfo[4] = s->SimPM.simtime;
  dinfo[5] = s->SimPM.dt;
  dinfo[6] = s->SimPM.last_dt;
  dinfo[7] = s->SimPM.finishtime;
  for (int a = 0; a < 3; a++) {
    dinfo[8 + a] = s->SimPM.Xmin[a];
    dinfo[11 + a] = s->SimPM.Xmax[a];
  }
  dinfo[14] = s->SimPM.EP.MinTemperature;
  dinfo[15] = s->SimPM.EP.MaxTemperature;
  dinfo[16] = s->SimPM.starttime;
  return 0;
}

int pref_refvec(void *h, double *out) {
  RefSim *s = static_cast<RefSim *>(h);
  for (int v = 0; v < s->SimPM.nvar; v++) out[v] = s->SimPM.RefVec[v];
  return 0;
}

// which: 0=P, 1=Ph, 2=dU.  SoA [var][k][j][i] over the padded grid.
int pref_get_state(void *h, int which, double *out) {
  RefSim *s = static_cast<RefSim *>(h);
  const long n = s->ncell_all();
  const int nv = s->SimPM.nvar;
  cell *c = s->g()->FirstPt_All();
  do {
    long ix = s->index_of(c);
    const pion_flt *src = (which == 0) ? c->P : (which == 1) ? c->Ph : c->dU;
    for (int v = 0; v < nv; v++) out[v * n + ix] = src[v];
  } while ((c = s->g()->NextPt_All(c)) != 0);
  return 0;
}

int pref_set_state(void *h, int which, const double *in) {
  RefSim *s = static_cast<RefSim *>(h);
  const long n = s->ncell_all();
  const int nv = s->SimPM.nvar;
  cell *c = s->g()->FirstPt_All();
  do {
    long ix = s->index_of(c);
    pion_flt *dst = (which == 0) ? c->P : (which == 1) ? c->Ph : c->dU;
    for (int v = 0; v < nv; v++) dst[v] = in[v * n + ix];
  } while ((c = s->g()->NextPt_All(c)) != 0);
  return 0;
}

// per-cell flags, one int per padded cell: bit0 isgd, bit1 isbd, bit2 isdomain,
// bit3 isleaf, bit4 timestep
int pref_get_flags(void *h, int *out) {
  RefSim *s = static_cast<RefSim *>(h);
  cell *c = s->g()->FirstPt_All();
  do {
    long ix = s->index_of(c);
    out[ix] = (c->isgd ? 1 : 0) | (c->isbd ? 2 : 0) | (c->isdomain ? 4 : 0) | (c->isleaf ? 8 : 0) |
              (c->timestep ? 16 : 0);
  } while ((c = s->g()->NextPt_All(c)) != 0);
  return 0;
}

// HLLD pre-processing scalars (divV, |grad p|/p) and H-correction eta, as
// stored by FV_solver_base::preprocess_data (solver_eqn_base.cpp:353-415).
int pref_get_extra(void *h, int what, int axis, double *out) {
  RefSim *s = static_cast<RefSim *>(h);
  cell *c = s->g()->FirstPt_All();
  do {
    long ix = s->index_of(c);
    if (what == 0) out[ix] = CI.get_DivV(c);
    else if (what == 1) out[ix] = CI.get_MagGradP(c);
    else out[ix] = CI.get_Hcorr(c, static_cast<axes>(axis));
  } while ((c = s->g()->NextPt_All(c)) != 0);
  return 0;
}

int pref_init_after_state(void *h) { return static_cast<RefSim *>(h)->init_after_state(); }
double pref_calc_timestep(void *h) { return static_cast<RefSim *>(h)->do_calc_timestep(); }
double pref_advance(void *h) { return static_cast<RefSim *>(h)->do_advance(); }
double pref_dynamics_dt(void *h) { return static_cast<RefSim *>(h)->do_dynamics_dt(); }
double pref_microphysics_dt(void *h) { return static_cast<RefSim *>(h)->do_microphysics_dt(); }

// nsteps of the sim_control::Time_Int loop body; returns steps taken.  dts (if
// non-null) receives the dt of each step.
int pref_run(void *h, int nsteps, double *dts) {
  RefSim *s = static_cast<RefSim *>(h);
  for (int i = 0; i < nsteps; i++) {
    double dt = s->do_calc_timestep();
    if (dt <= 0) return i;
    s->do_advance();
    if (dts) dts[i] = dt;
    s->output_bookkeeping();
  }
  return nsteps;
}

int pref_update_bcs(void *h, int cstep, int maxstep) {
  return static_cast<RefSim *>(h)->do_update_bcs(cstep, maxstep);
}
int pref_dynamics_dU(void *h, double dt, int step) { return static_cast<RefSim *>(h)->do_dynamics_dU(dt, step); }
int pref_microphysics_dU(void *h, double dt) { return static_cast<RefSim *>(h)->do_microphysics_dU(dt); }
int pref_update_state(void *h, double dt, int step, int ooa) {
  return static_cast<RefSim *>(h)->do_update_state(dt, step, ooa);
}
void pref_set_dt(void *h, double dt) {
  RefSim *s = static_cast<RefSim *>(h);
  s->SimPM.dt = dt;
  s->solver()->Setdt(dt);
}
void pref_set_glm_speeds(void *h, double tdyn, double dx, double cr) {
  static_cast<RefSim *>(h)->solver()->Set_GLM_Speeds(tdyn, dx, cr);
}
void pref_set_time(void *h, double simtime, double last_dt, int timestep) {
  RefSim *s = static_cast<RefSim *>(h);
  s->SimPM.simtime = simtime;
  s->SimPM.last_dt = last_dt;
  s->SimPM.timestep = timestep;
}

// One interface flux through the reference's FV_solver_base::InterCellFlux
// (solver_eqn_base.cpp:152) along `axis`, with the HLLD switch scalars and the
// H-correction eta supplied by the caller through two scratch cells.
int pref_intercell_flux(void *h, int axis, const double *Pl, const double *Pr, double divv_l, double gradp_l,
                        double divv_r, double gradp_r, double *flux) {
  RefSim *s = static_cast<RefSim *>(h);
  cell *cl = s->g()->FirstPt();
  cell *cr = s->g()->NextPt(cl, XP);
  double sl[2] = {0, 0}, sr[2] = {0, 0};
  const bool hlld = (s->SimPM.solverType == FLUX_RS_HLLD);
  if (hlld) {
    sl[0] = CI.get_DivV(cl); sl[1] = CI.get_MagGradP(cl);
    sr[0] = CI.get_DivV(cr); sr[1] = CI.get_MagGradP(cr);
    CI.set_DivV(cl, divv_l); CI.set_MagGradP(cl, gradp_l);
    CI.set_DivV(cr, divv_r); CI.set_MagGradP(cr, gradp_r);
  }
  const int nv = s->SimPM.nvar;
  vector<pion_flt> l(Pl, Pl + nv), r(Pr, Pr + nv), f(nv, 0.0);
  s->solver()->SetDirection(static_cast<axes>(axis));
  int err = s->solver()->InterCellFlux(s->SimPM, s->g(), cl, cr, &l[0], &r[0], &f[0], s->SimPM.gamma, s->SimPM.dx);
  s->solver()->SetDirection(XX);
  for (int v = 0; v < nv; v++) flux[v] = f[v];
  if (hlld) {
    CI.set_DivV(cl, sl[0]); CI.set_MagGradP(cl, sl[1]);
    CI.set_DivV(cr, sr[0]); CI.set_MagGradP(cr, sr[1]);
  }
  return err;
}

// The lookup tables mp_only_cooling builds at set-up (private member `lt`,
// microphysics/mp_only_cooling.cpp:528-556) re-evaluated through the SAME public
// rate functions of the reference's MP object, so that the oracle port and the GPU
// library can be fed bit-identical tables (the GSL spline behind
// cooling_rate_SD93CIE is a shim here, SURVEY 8c "parity unpinned" for its values).
int pref_cooling_tables(void *h, int n, double *T, double *rrhp, double *C_rrh, double *C_ffhe, double *C_fbdn,
                        double *C_cie) {
  RefSim *s = static_cast<RefSim *>(h);
  class mp_only_cooling *mp = dynamic_cast<class mp_only_cooling *>(MP);
  if (!mp || n < 2) return 1;
  const double Tmin = s->SimPM.EP.MinTemperature, Tmax = s->SimPM.EP.MaxTemperature;
  const double dlogT = (log10(Tmax) - log10(Tmin)) / (n - 1);
  for (int i = 0; i < n; i++) {
    T[i] = pow(10.0, log10(Tmin) + i * dlogT);
    rrhp[i] = mp->Hii_rad_recomb_rate(T[i]);
    C_rrh[i] = mp->Hii_total_cooling(T[i]);
    C_ffhe[i] = 6.72e-28 * sqrt(T[i]);
    C_fbdn[i] = 1.20e-22 * exp(-33610.0 / T[i] - (2180.0 * 2180.0 / T[i] / T[i])) * exp(-T[i] * T[i] / 5.0e10);
    C_cie[i] = mp->cooling_rate_SD93CIE(T[i]);
  }
  return 0;
}

// Knots (log10 T, log10 Lambda) and out-of-table slopes of the cooling-curve spline the reference's MP object
// set up (cooling_function_SD93CIE: setup_SD93_cie for EP_cooling 2..5, setup_WSS09_CIE for 6, 7,
// setup_WSS09_CIE_OnlyMetals for 8), so that the oracle port and the GPU library are fed the reference's own
// table.  slopes[0] = MinSlope, slopes[1] = MaxSlope.  Returns the number of knots (arrays may be null).
int pref_cooling_spline(void *h, double *logT, double *logL, double *slopes) {
  (void)h;
  class cooling_function_SD93CIE *cf = dynamic_cast<class cooling_function_SD93CIE *>(MP);
  if (!cf) return 0;
  const int n = cf->Nspl;
  for (int i = 0; i < n; i++) {
    if (logT) logT[i] = cf->Tarray[i];
    if (logL) logL[i] = cf->Larray[i];
  }
  if (slopes) { slopes[0] = cf->MinSlope; slopes[1] = cf->MaxSlope; }
  return n;
}

// The reference's own ASCII writer (dataio_text::OutputData -> output_ascii_data, dataio_text.cpp:130-146,477-555)
// on the current state: writes <base>.txt (counter < 0, as icgen writes initial conditions) or
// <base>.<counter, 8 digits>.txt.  This is the file pion_ugs_gpu --in-text consumes and --out-text reproduces.
int pref_output_text(void *h, const char *base, long counter) {
  RefSim *s = static_cast<RefSim *>(h);
  class dataio_text dio(s->SimPM);
  dio.SetSolver(s->solver());
  return dio.OutputData(base, s->grid, s->SimPM, counter);
}

}  // extern "C"

