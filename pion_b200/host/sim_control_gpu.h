// pion_b200/host/sim_control_gpu.h -- C++ host side above the C ABI.
//
// PION's extension seam is C++ inheritance: sim_control_pllel and sim_control_NG
// override the grid-level methods of time_integrator / calc_timestep /
// assign_update_bcs.  `sim_control_gpu` is the same shape of class
// for the B200 library: identical method names, argument meaning and error
// convention (int error count, 0 = success; failures carry the text the reference
// would pass to rep.error), with the device-resident grid behind the handle instead
// of a GridBaseClass*.  INTEGRATION.md shows the ~60-line adaptor that derives it
// from the reference's own `sim_control` (linked-list grid <-> SoA copy at Init /
// output_data); this header has no dependency on the reference's headers so that it
// builds and is tested stand-alone.
//
// Reference methods mirrored (paths relative to /root/reference/source):
//   sim_control/sim_init.cpp:58-300        Init
//   sim_control/sim_control.cpp:203-290    Time_Int (+ the cell-updates/s line, :270-277)
//   sim_control/sim_control.cpp:317-392    check_eosim
//   sim_control/calc_timestep.h:54-95      calculate_timestep, calc_dynamics_dt, calc_microphysics_dt
//   sim_control/time_integrator.h:52-179   advance_time, calc_microphysics_dU, calc_dynamics_dU,
//                                          grid_update_state_vector
//   boundaries/assign_update_bcs.h:51-73   TimeUpdateInternalBCs / TimeUpdateExternalBCs
#ifndef PION_B200_SIM_CONTROL_GPU_H
#define PION_B200_SIM_CONTROL_GPU_H

#include <string>
#include <vector>

#include "../../include/pion_b200.h"

namespace pion_b200 {

// The members of `class SimParams` (sim_params.h:200-285) the path reads, same names.
struct SimParamsGPU {
  int ndim = 3, eqntype = PION_EQGLM, coord_sys = PION_COORD_CRT, solverType = PION_FLUX_HLLD;
  int nvar = 9, ntracer = 0, artviscosity = 1, spOOA = 2, tmOOA = 2;
  int NG[3] = {1, 1, 1};
  double Xmin[3] = {0, 0, 0}, Xmax[3] = {1, 1, 1};
  double gamma = 5.0 / 3.0, CFL = 0.3, etav = 0.15;
  int BC[6] = {PION_BC_OUTFLOW, PION_BC_OUTFLOW, PION_BC_OUTFLOW, PION_BC_OUTFLOW, PION_BC_OUTFLOW, PION_BC_OUTFLOW};
  std::vector<int> BC_internal;
  double RefVec[PION_GPU_MAXVAR];
  double starttime = 0, finishtime = 1e30, simtime = 0, dt = 0, last_dt = 1e100;
  int timestep = 0, op_criterion = 0, opfreq = 0;
  double opfreq_time = 0;
  double min_timestep = 0;  // sim_params.h:227
  bool maxtime = false;
  // struct which_physics EP (sim_params.h:106-160), cooling-only microphysics
  struct {
    int cooling = 0, MP_timestep_limit = 0;
    double MinTemperature = 0, MaxTemperature = 1e99;
  } EP;
  // struct stellarwind_list SWP (sim_params.h:340-380): constant wind sources
  std::vector<pion_gpu_wind_source> SWP;
  // mp_only_cooling lookup tables (mp_only_cooling.cpp:528-556), 6 columns of n_table values
  std::vector<double> table_T, table_rrhp, table_C_rrh, table_C_ffhe, table_C_fbdn, table_C_cie;
  long Ncell() const { return (long)NG[0] * NG[1] * NG[2]; }
  SimParamsGPU() { for (double& r : RefVec) r = 1.0; }
};

class sim_control_gpu {
 public:
  sim_control_gpu();
  ~sim_control_gpu();
  sim_control_gpu(const sim_control_gpu&) = delete;
  sim_control_gpu& operator=(const sim_control_gpu&) = delete;

  SimParamsGPU SimPM;

  /// extents of the padded SoA state [nvar][NZ+2g][NY+2g][NX+2g] exchanged with the device
  void padded_extents(int ext[3]) const;
  size_t state_size() const;

  /// sim_init::Init: create the device grid on `device`, upload the primitive state
  /// (ghost cells need not be set), Ph=P, boundary assignment + first boundary update.
  int Init(int device, const double* P_soa);
  /// sim_control::Time_Int: { calculate_timestep; advance_time; check_eosim } until
  /// maxtime or `max_steps` steps; prints the reference's TOTALS / STEPS lines.
  int Time_Int(long max_steps = -1, bool verbose = false);
  /// sim_control::Finalise
  int Finalise();

  // ---- the grid-level seam, names as in the reference ----
  int calculate_timestep();                                  // calc_timestep.h:54
  double calc_dynamics_dt();                                 // calc_timestep.h:95
  double calc_microphysics_dt();                             // calc_timestep.h:67
  double advance_time();                                     // time_integrator.h:52 (returns dt)
  int calc_microphysics_dU(double dt);                       // time_integrator.h:92
  int calc_dynamics_dU(double dt, int step);                 // time_integrator.h:133
  int grid_update_state_vector(double dt, int step, int ooa);  // time_integrator.h:179
  int TimeUpdateInternalBCs(double simtime, int cstep, int maxstep);  // assign_update_bcs.h:51
  int TimeUpdateExternalBCs(double simtime, int cstep, int maxstep);  // assign_update_bcs.h:63
  /// sim_init::output_data (sim_init.cpp:671-760): applies the output criteria (op_criterion / opfreq /
  /// opfreq_time, consuming an output time that has been reached) and, when the step is to be saved, copies
  /// P back into the host array (null = bookkeeping only).  `saved` reports the decision.
  int output_data(double* P_soa, bool* saved = nullptr);
  /// unconditional device -> host copy of P (dataio->OutputData side of the seam)
  int download_state(double* P_soa);
  /// host buffer Time_Int hands to output_data after every step (null = bookkeeping only)
  void set_output_buffer(double* P_soa) { output_buffer_ = P_soa; }
  /// negative density / failed cooling integrations are fatal in the reference: Time_Int polls the device
  /// counters every this many steps (0 = only at the end)
  int fatal_poll_steps = 16;
  int check_fatal_counters();
  int check_eosim();                                         // sim_control.cpp:317

  /// last error text (what the reference would hand to rep.error)
  const std::string& error() const { return err_; }
  double wall_seconds() const { return wall_; }
  pion_gpu_ctx* handle() { return ctx_; }

 private:
  int fail(const char* where);
  void pull_time();
  pion_gpu_ctx* ctx_ = nullptr;
  double* output_buffer_ = nullptr;
  std::string err_;
  double wall_ = 0;
};

}  // namespace pion_b200
#endif
