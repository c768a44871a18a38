// pion_b200/host/sim_control_gpu.cpp -- see sim_control_gpu.h.  Every numerical step is
// one C-ABI call into libpion_b200.so; there is no host-side arithmetic on the state.
#include "sim_control_gpu.h"

#include <chrono>
#include <cstdio>
#include <cstring>

namespace pion_b200 {

sim_control_gpu::sim_control_gpu() {}
sim_control_gpu::~sim_control_gpu() { Finalise(); }

int sim_control_gpu::fail(const char* where) {
  err_ = std::string(where) + ": " + pion_gpu_last_error();
  return 1;
}

void sim_control_gpu::padded_extents(int ext[3]) const {
  const int g = (SimPM.spOOA == 2) ? 2 : 1;  // setup_fixed_grid.cpp:183-184
  for (int a = 0; a < 3; a++) ext[a] = (a < SimPM.ndim) ? SimPM.NG[a] + 2 * g : 1;
}
size_t sim_control_gpu::state_size() const {
  int e[3];
  padded_extents(e);
  return (size_t)SimPM.nvar * e[0] * e[1] * e[2];
}

void sim_control_gpu::pull_time() {
  pion_gpu_get_time(ctx_, &SimPM.simtime, &SimPM.dt, &SimPM.last_dt, &SimPM.timestep);
}

int sim_control_gpu::Init(int device, const double* P_soa) {
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
    c.xmin[a] = c.sim_xmin[a] = SimPM.Xmin[a];
    c.xmax[a] = SimPM.Xmax[a];
  }
  for (int f = 0; f < 6; f++) {
    c.bc[f] = (f < 2 * SimPM.ndim) ? SimPM.BC[f] : 0;
    c.ngbprocs[f] = -1;
  }
  c.n_internal_bc = (int)SimPM.BC_internal.size();
  for (int i = 0; i < c.n_internal_bc && i < 4; i++) c.internal_bc[i] = SimPM.BC_internal[i];
  for (int v = 0; v < PION_GPU_MAXVAR; v++) c.refvec[v] = SimPM.RefVec[v];
  c.starttime = SimPM.starttime;
  c.finishtime = SimPM.finishtime;
  c.op_criterion = SimPM.op_criterion;
  c.opfreq_time = SimPM.opfreq_time;
  c.cooling = SimPM.EP.cooling;
  c.mp_timestep_limit = SimPM.EP.MP_timestep_limit;
  c.min_temperature = SimPM.EP.MinTemperature;
  c.max_temperature = SimPM.EP.MaxTemperature;
  c.n_table = (int)SimPM.table_T.size();
  c.table_T = SimPM.table_T.data();
  c.table_rrhp = SimPM.table_rrhp.data();
  c.table_C_rrh = SimPM.table_C_rrh.data();
  c.table_C_ffhe = SimPM.table_C_ffhe.data();
  c.table_C_fbdn = SimPM.table_C_fbdn.data();
  c.table_C_cie = SimPM.table_C_cie.data();
  c.n_wind = (int)SimPM.SWP.size();
  for (int i = 0; i < c.n_wind && i < 2; i++) c.wind[i] = SimPM.SWP[i];
  c.min_timestep = SimPM.min_timestep;
  c.rank = 0;
  c.nproc = 1;
  Finalise();
  ctx_ = pion_gpu_create(&c);
  if (!ctx_) return fail("setup_grid");
  if (pion_gpu_set_time(ctx_, SimPM.simtime, SimPM.last_dt, SimPM.timestep)) return fail("Init");
  if (pion_gpu_upload(ctx_, PION_STATE_P, P_soa)) return fail("ReadData");
  if (pion_gpu_init_after_upload(ctx_)) return fail("boundary_conditions/assign_boundary_data");
  SimPM.maxtime = false;
  return 0;
}

int sim_control_gpu::Finalise() {
  if (ctx_) pion_gpu_destroy(ctx_);
  ctx_ = nullptr;
  return 0;
}

int sim_control_gpu::calculate_timestep() {
  double dt;
  if (pion_gpu_calculate_timestep(ctx_, &dt)) return fail("calculate_timestep");
  SimPM.dt = dt;
  return 0;
}
double sim_control_gpu::calc_dynamics_dt() {
  double td = -1, tm;
  if (pion_gpu_calc_dt(ctx_, &td, &tm)) { fail("calc_dynamics_dt"); return -1.0; }
  return td;
}
double sim_control_gpu::calc_microphysics_dt() {
  double td, tm = -1;
  if (pion_gpu_calc_dt(ctx_, &td, &tm)) { fail("calc_microphysics_dt"); return -1.0; }
  return tm;
}
double sim_control_gpu::advance_time() {
  double dt = 0;
  if (pion_gpu_advance_time(ctx_, &dt)) { fail("advance_time"); return -1.0; }
  pull_time();
  return dt;
}
int sim_control_gpu::calc_microphysics_dU(double dt) {
  return pion_gpu_calc_microphysics_dU(ctx_, dt) ? fail("calc_microphysics_dU") : 0;
}
int sim_control_gpu::calc_dynamics_dU(double dt, int step) {
  if (pion_gpu_set_dt(ctx_, dt)) return fail("Setdt");  // spatial_solver->Setdt(dt)
  return pion_gpu_calc_dynamics_dU(ctx_, dt, step) ? fail("calc_dynamics_dU") : 0;
}
int sim_control_gpu::grid_update_state_vector(double dt, int step, int ooa) {
  return pion_gpu_grid_update_state_vector(ctx_, dt, step, ooa) ? fail("grid_update_state_vector") : 0;
}
// assign_update_bcs.cpp:134-181 (STWIND cells) and :191-246 (the faces in BC_bd order, then DMACH2)
int sim_control_gpu::TimeUpdateInternalBCs(double simtime, int cstep, int maxstep) {
  return pion_gpu_time_update_internal_bcs(ctx_, simtime, cstep, maxstep) ? fail("TimeUpdateInternalBCs") : 0;
}
int sim_control_gpu::TimeUpdateExternalBCs(double simtime, int cstep, int maxstep) {
  return pion_gpu_time_update_external_bcs(ctx_, simtime, cstep, maxstep) ? fail("TimeUpdateExternalBCs") : 0;
}
// sim_init::output_data (sim_init.cpp:671-760): the output criteria decide whether this step is saved; an
// output time that has been reached is consumed (next_optime += opfreq_time).  P_soa may be null (no copy).
int sim_control_gpu::output_data(double* P_soa, bool* saved) {
  int due = 0;
  if (pion_gpu_output_due(ctx_, SimPM.opfreq, &due)) return fail("output_data");
  if (saved) *saved = due != 0;
  if (due && P_soa) return pion_gpu_download(ctx_, PION_STATE_P, P_soa) ? fail("output_data") : 0;
  return 0;
}
int sim_control_gpu::download_state(double* P_soa) {
  return pion_gpu_download(ctx_, PION_STATE_P, P_soa) ? fail("download_state") : 0;
}
// the conditions that are fatal inside the reference's per-cell code (rep.error in UtoP / TimeUpdateMP)
int sim_control_gpu::check_fatal_counters() {
  long long cnt[3], mpf = 0;
  if (pion_gpu_counters(ctx_, cnt) || pion_gpu_mp_failures(ctx_, &mpf)) return fail("Time_Int");
  if (cnt[0]) { err_ = "UtoP: negative density (fatal in the reference, eqns_mhd_adiabatic.cpp:137)"; return 1; }
  if (mpf) { err_ = "mp_only_cooling integration failed."; return 1; }
  return 0;
}

// sim_control::check_eosim (sim_control.cpp:317-392), time criterion
int sim_control_gpu::check_eosim() {
  if (SimPM.simtime >= SimPM.finishtime) SimPM.maxtime = true;
  return 0;
}

int sim_control_gpu::Time_Int(long max_steps, bool verbose) {
  if (!ctx_) { err_ = "Time_Int before Init"; return 1; }
  SimPM.maxtime = false;
  const int step0 = SimPM.timestep;
  if (pion_gpu_sync(ctx_)) return fail("Time_Int");
  const auto t0 = std::chrono::steady_clock::now();
  long n = 0;
  while (!SimPM.maxtime && (max_steps < 0 || n < max_steps)) {
    if (calculate_timestep()) return 1;
    if (advance_time() < 0.0) return 1;
    if (verbose) printf("New time: %.10e\t dt=%.10e\t steps: %d\n", SimPM.simtime, SimPM.dt, SimPM.timestep);
    check_eosim();
    if (output_data(output_buffer_, nullptr)) return 1;  // sim_control.cpp:252
    n++;
    // the reference aborts inside the step; here the device counters are polled every few steps (one 24-byte
    // read-back) so that a failed run stops within `fatal_poll_steps` steps instead of running on to the end
    if (fatal_poll_steps > 0 && (n % fatal_poll_steps) == 0 && check_fatal_counters()) return 1;
  }
  if (pion_gpu_sync(ctx_)) return fail("Time_Int");
  wall_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const int steps = SimPM.timestep - step0;
  // the reference's closing lines (sim_control.cpp:270-277)
  printf("TOTALS ###: Nsteps: %d wall-time: %g time/step: %g\n", steps, wall_, wall_ / (steps > 0 ? steps : 1));
  printf("STEPS: %d\t%.6e\t%.6e\t%.6e\n", steps, wall_, wall_ / (steps > 0 ? steps : 1),
         (double)steps * (double)SimPM.Ncell() / wall_);
  return check_fatal_counters();
}

}  // namespace pion_b200
