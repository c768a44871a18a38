// integration/sim_control_gpu_ref.h -- the reference-side binding of libpion_b200, COMPILED against the
// reference's own headers (/root/reference/source, read-only): `class sim_control_gpu : public sim_control`.
//
// This is the file a PION maintainer adds to source/sim_control/ (INTEGRATION.md section 2).  It follows the
// precedent of sim_control_pllel (sim_control_MPI.cpp:482) and sim_control_NG (sim_control_NG.cpp:602-777):
// override the grid-level virtuals and leave everything else -- parameter parsing, grid / boundary / IC set-up,
// the time loop sim_control::Time_Int (sim_control.cpp:198-280), the output criteria and the writers
// (sim_init::output_data, dataio_*) -- to the reference.  Overridden:
//   sim_init::Init                       (sim_init.h:54)     base Init, then the device grid is created from it
//   calc_timestep::calculate_timestep    (calc_timestep.h:54)
//   time_integrator::advance_time        (time_integrator.h:52)
//   sim_init::output_data                (sim_init.h:114)    device -> linked-list grid when a write is due
//   sim_control::Finalise                (sim_control.h:62)
// It has no numerics of its own: every number comes out of libpion_b200.so through include/pion_b200.h.
//
// In THIS repository it is test infrastructure: it can only be built where /root/reference is present
// (oracle/Makefile target `gpuref`, linked with the oracle/_ref objects) and is exercised by
// tests/test_reference_binding.py on the GPU box.
#ifndef SIM_CONTROL_GPU_REF_H
#define SIM_CONTROL_GPU_REF_H

#include <string>
#include <vector>

#include "sim_control/sim_control.h"
#include "pion_b200.h"

class sim_control_gpu : public sim_control {
 public:
  sim_control_gpu();
  ~sim_control_gpu();

  /// CUDA device the grid lives on (set before Init / gpu_attach)
  void set_device(int d) { device = d; }

  /// sim_init::Init, then gpu_attach
  virtual int Init(string, int, int, string *, vector<class GridBaseClass *> &);

  /// Build the device-resident copy of an initialised grid: configuration from SimPM / SWP / the grid's
  /// boundary list, P uploaded, then the library repeats the post-ReadData part of Init (sim_init.cpp:215-280)
  /// on the device.  Called by Init; public so that a caller that sets the grid up the way icgen does
  /// (ics/icgen.cpp:90-330) can attach without a restart file.
  int gpu_attach(class GridBaseClass *grid);

  virtual int calculate_timestep(class SimParams &, class GridBaseClass *, class FV_solver_base *, const int);
  virtual double advance_time(const int, class GridBaseClass *);
  virtual int output_data(vector<class GridBaseClass *> &);
  virtual int Finalise(vector<class GridBaseClass *> &);

  /// device -> linked list (P and Ph of every cell, ghost cells included)
  int sync_grid_from_device(class GridBaseClass *grid);
  pion_gpu_ctx *handle() { return ctx; }
  /// the driver of THIS repository has no dataio object when it sets the grid up icgen-style: then
  /// output_data keeps the criteria bookkeeping and skips the writers
  void allow_missing_dataio(bool b) { no_dataio_ok = b; }

 protected:
  pion_gpu_ctx *ctx;
  int device;
  bool no_dataio_ok;
  std::vector<double> soa;
  std::vector<double> tab[6];  // mp_only_cooling lookup columns handed to the library
  int copy_grid(class GridBaseClass *grid, bool to_device);
  void pull_time();
};

#endif
