// TEST INFRASTRUCTURE ONLY.  integration/gpu_ref_driver.cpp -- C-ABI harness around `sim_control_gpu`
// (sim_control_gpu_ref.h), the reference-derived binding of libpion_b200: sets a grid up the way the
// reference's icgen does (ics/icgen.cpp:90-330: get_sim_info parameter file, setup_grid, set_equations,
// setup_microphysics, boundary_conditions, the reference's own IC classes), attaches the device grid and then
// runs the reference's UNMODIFIED time loop sim_control::Time_Int (sim_control.cpp:198-280), which calls the
// overridden calculate_timestep / advance_time / output_data.  No physics here.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "defines/functionality_flags.h"
#include "defines/testing_flags.h"
#include "sim_constants.h"
#include "tools/reporting.h"
#include "grid/grid_base_class.h"
#include "grid/cell_interface.h"
#include "ics/icgen_base.h"
#include "ics/icgen.h"
#include "ics/get_sim_info.h"
#include "dataIO/readparams.h"
#include "microphysics/microphysics_base.h"

#include "sim_control_gpu_ref.h"

using namespace std;

namespace {
struct NullBuf : public std::streambuf {
  int overflow(int c) override { return c; }
};
NullBuf g_nullbuf;
std::streambuf *g_saved = 0;
void quiet_on() {
  if (getenv("PION_REF_VERBOSE")) return;
  if (!g_saved) g_saved = std::cout.rdbuf(&g_nullbuf);
}

class GpuRefSim : public sim_control_gpu {
 public:
  vector<class GridBaseClass *> grid;
  class ReadParams *rp = 0;
  class ICsetup_base *ic = 0;
  int NGa[3] = {1, 1, 1};
  int ioff[3] = {0, 0, 0};
  ~GpuRefSim() {
    if (rp) delete rp;
    if (ic) delete ic;
  }
  int setup(const char *pfile, int dev) {
    int err = 0;
    MP = 0;
    SWP.params.clear();
    SWP.Nsources = 0;
    {
      class get_sim_info siminfo;
      err += siminfo.read_gridparams(pfile, SimPM);
      if (err) return err;
    }
    SimPM.levels.clear();
    SimPM.levels.resize(1);
    SimPM.grid_nlevels = 1;
    SimPM.levels[0].parent = 0;
    SimPM.levels[0].child = 0;
    SimPM.levels[0].Ncell = SimPM.Ncell;
    for (int v = 0; v < MAX_DIM; v++) {
      SimPM.levels[0].NG[v] = SimPM.NG[v];
      SimPM.levels[0].Range[v] = SimPM.Range[v];
      SimPM.levels[0].Xmin[v] = SimPM.Xmin[v];
      SimPM.levels[0].Xmax[v] = SimPM.Xmax[v];
    }
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
    string seek = "ics";
    string ics = rp->find_parameter(seek);
    setup_ics_type(ics, &ic);
    ic->set_SimPM(&SimPM);
    err += ic->setup_data(rp, grid[0]);
    if (err) return err;
    // what sim_init::Init does with the data it has read is done by the library on the device
    // (pion_gpu_init_after_upload); the writers are absent here (no Silo / FITS in this environment)
    set_device(dev);
    allow_missing_dataio(true);
    return gpu_attach(grid[0]);
  }
  long index_of(const cell *c) const {
    long idx[3] = {0, 0, 0};
    for (int a = 0; a < SimPM.ndim; a++) idx[a] = (c->pos[a] - ioff[a]) / 2;
    return idx[0] + NGa[0] * (idx[1] + (long)NGa[1] * idx[2]);
  }
};
}  // namespace

extern "C" {

void *pgr_create(const char *paramfile, int device) {
  quiet_on();
  GpuRefSim *s = new GpuRefSim();
  if (s->setup(paramfile, device)) {
    fprintf(stderr, "pgr_create: set-up failed\n");
    delete s;
    return 0;
  }
  return s;
}
void pgr_destroy(void *h) {
  GpuRefSim *s = static_cast<GpuRefSim *>(h);
  if (!s) return;
  GridBaseClass *g = s->grid.size() ? s->grid[0] : 0;
  delete s;
  if (g) delete g;
}
// the reference's own time loop, to SimPM.finishtime; returns its error code
int pgr_time_int(void *h) {
  GpuRefSim *s = static_cast<GpuRefSim *>(h);
  int err = s->Time_Int(s->grid);
  err += s->Finalise(s->grid);
  return err;
}
// info: [0..2] padded extents, [3] nvar, [4] timestep; dinfo: [0] simtime, [1] last dt
int pgr_info(void *h, int *info, double *dinfo) {
  GpuRefSim *s = static_cast<GpuRefSim *>(h);
  for (int a = 0; a < 3; a++) info[a] = s->NGa[a];
  info[3] = s->SimPM.nvar;
  info[4] = s->SimPM.timestep;
  dinfo[0] = s->SimPM.simtime;
  dinfo[1] = s->SimPM.last_dt;
  return 0;
}
// P of the reference's linked-list grid as SoA [var][k][j][i] over the padded grid
int pgr_get_state(void *h, double *out) {
  GpuRefSim *s = static_cast<GpuRefSim *>(h);
  const long n = (long)s->NGa[0] * s->NGa[1] * s->NGa[2];
  cell *c = s->grid[0]->FirstPt_All();
  do {
    const long ix = s->index_of(c);
    for (int v = 0; v < s->SimPM.nvar; v++) out[v * n + ix] = c->P[v];
  } while ((c = s->grid[0]->NextPt_All(c)) != 0);
  return 0;
}
int pgr_describe(void *h, char *buf, int n) {
  return pion_gpu_describe(static_cast<GpuRefSim *>(h)->handle(), buf, n);
}

}  // extern "C"
