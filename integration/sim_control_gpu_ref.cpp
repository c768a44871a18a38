// integration/sim_control_gpu_ref.cpp -- see sim_control_gpu_ref.h.
#include "sim_control_gpu_ref.h"

#include <cmath>
#include <cstring>

#include "microphysics/microphysics_base.h"
#include "microphysics/mp_only_cooling.h"
#include "tools/reporting.h"

using namespace std;

sim_control_gpu::sim_control_gpu() : sim_control(), ctx(0), device(0), no_dataio_ok(false) {}
sim_control_gpu::~sim_control_gpu() {
  if (ctx) pion_gpu_destroy(ctx);
  ctx = 0;
}

void sim_control_gpu::pull_time() {
  pion_gpu_get_time(ctx, &SimPM.simtime, &SimPM.dt, &SimPM.last_dt, &SimPM.timestep);
}

int sim_control_gpu::Init(string infile, int typeOfFile, int narg, string *args, vector<class GridBaseClass *> &grid) {
  int err = sim_control::Init(infile, typeOfFile, narg, args, grid);
  if (err) return err;
  return gpu_attach(grid[0]);
}

// cell-id order of the linked list is x fastest over the PADDED grid (uniform_grid.cpp:449-451):
// FirstPt_All() / NextPt_All() visit the cells in exactly the order of the SoA planes.
int sim_control_gpu::copy_grid(class GridBaseClass *grid, bool to_device) {
  const size_t nall = grid->Ncell_all();
  const int nv = SimPM.nvar;
  soa.resize(nall * nv);
  if (!to_device && pion_gpu_download(ctx, PION_STATE_P, soa.data())) return 1;
  size_t i = 0;
  for (cell *c = grid->FirstPt_All(); c != 0; c = grid->NextPt_All(c), i++) {
    if (i >= nall) rep.error("sim_control_gpu::copy_grid: more cells than Ncell_all", i);
    for (int v = 0; v < nv; v++) {
      if (to_device) soa[v * nall + i] = c->P[v];
      else c->P[v] = c->Ph[v] = soa[v * nall + i];
    }
  }
  if (i != nall) rep.error("sim_control_gpu::copy_grid: cell count", i);
  if (to_device && pion_gpu_upload(ctx, PION_STATE_P, soa.data())) return 1;
  return 0;
}

int sim_control_gpu::sync_grid_from_device(class GridBaseClass *grid) { return copy_grid(grid, false); }

int sim_control_gpu::gpu_attach(class GridBaseClass *grid) {
  pion_gpu_config c;
  memset(&c, 0, sizeof(c));
  c.device = device;
  c.ndim = SimPM.ndim;
  c.nvar = SimPM.nvar;
  c.ntracer = SimPM.ntracer;
  c.eqntype = SimPM.eqntype;
  c.coord_sys = SimPM.coord_sys;
  c.solver = SimPM.solverType;
  c.artviscosity = SimPM.artviscosity;
  c.spOOA = SimPM.spOOA;
  c.tmOOA = SimPM.tmOOA;
  c.gamma = SimPM.gamma;
  c.cfl = SimPM.CFL;
  c.etav = SimPM.etav;
  for (int a = 0; a < 3; a++) {
    c.NG[a] = (a < SimPM.ndim) ? SimPM.NG[a] : 1;
    c.xmin[a] = c.sim_xmin[a] = (a < SimPM.ndim) ? SimPM.Xmin[a] : 0.0;
    c.xmax[a] = (a < SimPM.ndim) ? SimPM.Xmax[a] : 1.0;
  }
  // the grid's boundary list (set up by boundary_conditions(), setup_fixed_grid.cpp:1230-1420): the 2*ndim
  // external faces in the order XN,XP,YN,YP,ZN,ZP, then the internal boundaries; itype is the BoundaryTypes
  // enum (boundaries/boundaries.h:31-60), whose values the PION_BC_* codes are
  const size_t nface = 2 * SimPM.ndim;
  if (grid->BC_bd.size() < nface) rep.error("sim_control_gpu: boundaries not set up before gpu_attach", grid->BC_bd.size());
  for (size_t f = 0; f < 6; f++) {
    c.bc[f] = (f < nface) ? grid->BC_bd[f]->itype : 0;
    c.ngbprocs[f] = -1;
  }
  c.n_internal_bc = static_cast<int>(grid->BC_bd.size() - nface);
  if (c.n_internal_bc > 4) rep.error("sim_control_gpu: more than 4 internal boundaries", c.n_internal_bc);
  for (int i = 0; i < c.n_internal_bc; i++) c.internal_bc[i] = grid->BC_bd[nface + i]->itype;
  for (int v = 0; v < SimPM.nvar && v < PION_GPU_MAXVAR; v++) c.refvec[v] = SimPM.RefVec[v];
  c.starttime = SimPM.starttime;
  c.finishtime = SimPM.finishtime;
  c.op_criterion = SimPM.op_criterion;
  c.opfreq_time = SimPM.opfreq_time;
  c.min_timestep = SimPM.min_timestep;
  // microphysics: only the cooling-without-chemistry module (mp_only_cooling) is on the device
  c.cooling = 0;
  if (MP) {
    class mp_only_cooling *mp = dynamic_cast<class mp_only_cooling *>(MP);
    if (!mp) rep.error("sim_control_gpu: only mp_only_cooling microphysics is built on the device", SimPM.EP.cooling);
    c.cooling = SimPM.EP.cooling;
    c.mp_timestep_limit = SimPM.EP.MP_timestep_limit;
    c.min_temperature = SimPM.EP.MinTemperature;
    c.max_temperature = SimPM.EP.MaxTemperature;
    // the 200-point lookup columns of gen_mpoc_lookup_tables (mp_only_cooling.cpp:528-556; `lt` is private),
    // rebuilt from the object's public rate functions on the same temperature grid
    const int n = 200;
    const double Tmin = SimPM.EP.MinTemperature, Tmax = SimPM.EP.MaxTemperature;
    const double dlogT = (log10(Tmax) - log10(Tmin)) / (n - 1);
    for (int q = 0; q < 6; q++) tab[q].resize(n);
    for (int i = 0; i < n; i++) {
      const double T = pow(10.0, log10(Tmin) + i * dlogT);
      tab[0][i] = T;
      tab[1][i] = mp->Hii_rad_recomb_rate(T);
      tab[2][i] = mp->Hii_total_cooling(T);
      tab[3][i] = 6.72e-28 * sqrt(T);
      tab[4][i] = 1.20e-22 * exp(-33610.0 / T - (2180.0 * 2180.0 / T / T)) * exp(-T * T / 5.0e10);
      tab[5][i] = mp->cooling_rate_SD93CIE(T);
    }
    c.n_table = n;
    c.table_T = tab[0].data();
    c.table_rrhp = tab[1].data();
    c.table_C_rrh = tab[2].data();
    c.table_C_ffhe = tab[3].data();
    c.table_C_fbdn = tab[4].data();
    c.table_C_cie = tab[5].data();
  }
  // stellar winds (struct stellarwind_list SWP, sim_params.h:129-164): constant sources only
  c.n_wind = static_cast<int>(SWP.params.size());
  if (c.n_wind > 2) rep.error("sim_control_gpu: at most 2 wind sources", c.n_wind);
  for (int i = 0; i < c.n_wind; i++) {
    const struct stellarwind_params *w = SWP.params[i];
    if (w->type != 0) rep.error("sim_control_gpu: only constant winds (type 0) are built on the device", w->type);
    pion_gpu_wind_source &d = c.wind[i];
    for (int a = 0; a < 3; a++) d.dpos[a] = (a < SimPM.ndim) ? w->dpos[a] : 0.0;
    d.radius = w->radius;
    d.mdot = w->Mdot;
    d.vinf = w->Vinf;
    d.vrot = w->Vrot;
    d.temp = w->Tstar;
    d.rstar = w->Rstar;
    d.bsrf = w->Bstar;
    for (int t = 0; t < SimPM.ntracer && t < PION_GPU_MAXTR; t++) d.tr[t] = w->tr[t];
  }
  c.rank = 0;
  c.nproc = 1;
  if (ctx) pion_gpu_destroy(ctx);
  ctx = pion_gpu_create(&c);
  if (!ctx) rep.error(pion_gpu_last_error(), 1);
  if (pion_gpu_set_time(ctx, SimPM.simtime, SimPM.last_dt, SimPM.timestep)) rep.error(pion_gpu_last_error(), 2);
  if (copy_grid(grid, true)) rep.error(pion_gpu_last_error(), 3);
  if (pion_gpu_init_after_upload(ctx)) rep.error(pion_gpu_last_error(), 4);
  return 0;
}

int sim_control_gpu::calculate_timestep(class SimParams &par, class GridBaseClass *, class FV_solver_base *, const int) {
  if (pion_gpu_set_time(ctx, par.simtime, par.last_dt, par.timestep)) rep.error(pion_gpu_last_error(), 1);
  double dt = 0.0;
  const int err = pion_gpu_calculate_timestep(ctx, &dt);
  if (err) rep.error(pion_gpu_last_error(), err);
  par.dt = dt;
  return err;
}

double sim_control_gpu::advance_time(const int, class GridBaseClass *) {
  double dt = 0.0;
  if (pion_gpu_advance_time(ctx, &dt)) rep.error(pion_gpu_last_error(), 1);
  pull_time();
  // fatal inside the reference's per-cell code (rep.error in UtoP / TimeUpdateMP): same here, one step late
  long long cnt[3], mpf = 0;
  if (pion_gpu_counters(ctx, cnt) || pion_gpu_mp_failures(ctx, &mpf)) rep.error(pion_gpu_last_error(), 2);
  if (cnt[0]) rep.error("UtoP: negative density", cnt[0]);
  if (mpf) rep.error("mp_only_cooling integration failed.", mpf);
  return dt;
}

int sim_control_gpu::output_data(vector<class GridBaseClass *> &grid) {
  if (!ctx) return sim_control::output_data(grid);  // the initial write from inside the base Init
  // same decision as sim_init::output_data (sim_init.cpp:671-744), taken by the library on its own copy of
  // the time state (its dt limiter reads next_optime); the base class repeats it on SimPM's copy below
  int due = 0;
  if (pion_gpu_output_due(ctx, SimPM.opfreq, &due)) rep.error(pion_gpu_last_error(), 1);
  const int ckpt = (SimPM.checkpoint_freq > 0) ? SimPM.checkpoint_freq : 250;
  const bool checkpoint = (SimPM.timestep != 0) && ((SimPM.timestep % ckpt) == 0);
  if (!dataio) {
    if (!no_dataio_ok) rep.error("sim_control_gpu::output_data: no dataio object", 0);
    if (SimPM.op_criterion == 1 && due && SimPM.timestep != 0) SimPM.next_optime += SimPM.opfreq_time;
    return 0;
  }
  if (due || checkpoint) {
    if (copy_grid(grid[0], false)) rep.error(pion_gpu_last_error(), 2);
  }
  return sim_control::output_data(grid);  // the reference's writers see an up-to-date linked list
}

int sim_control_gpu::Finalise(vector<class GridBaseClass *> &grid) {
  if (ctx && copy_grid(grid[0], false)) rep.error(pion_gpu_last_error(), 1);
  if (!dataio && no_dataio_ok) return 0;
  return sim_control::Finalise(grid);
}
