// pion_b200/csrc/pion_b200.cu -- context, host-side orchestration and the C ABI
// of libpion_b200.so (see include/pion_b200.h for the reference method each
// entry point replaces).  The host logic mirrors, call for call,
//   sim_control/time_integrator.cpp:72-243   advance_time / first_order_update / second_order_update
//   sim_control/calc_timestep.cpp:68-262     calculate_timestep / timestep_checking_and_limiting
//   boundaries/assign_update_bcs.cpp:28-246  assign_boundary_data / TimeUpdate{Internal,External}BCs
//   sim_control/sim_init.cpp:215-280         Init after ReadData
// There is no CPU fallback: every numerical step is a kernel launch on the
// context's stream; the host only sequences launches and reads back scalars.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pion_b200.h"
#include "aux_kernels.cuh"
#include "cooling.cuh"

using namespace pion;

static thread_local std::string g_last_error;
static void set_error(const std::string& s) { g_last_error = s; }

#define CUDA_OK(call)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                        \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)
#define NCCL_OK(call)                                                                       \
  do {                                                                                      \
    ncclResult_t r_ = (call);                                                               \
    if (r_ != ncclSuccess) {                                                                \
      set_error(std::string(#call) + ": " + ncclGetErrorString(r_));                        \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

struct pion_gpu_ctx {
  pion_gpu_config cfg;
  GridD g;
  PhysParams pp;
  int nvar, nbase_, ntr;
  size_t arr_elems;  // doubles per state array
  double *P = nullptr, *Ph = nullptr, *dU = nullptr, *eta = nullptr;
  unsigned char *hll = nullptr, *hllf = nullptr, *mask = nullptr;
  unsigned long long* d_dtmin = nullptr;  // [0] running min for the next step, [1] scratch
  long long* d_counters = nullptr;
  unsigned long long* h_pinned = nullptr;  // pinned mirror for scalar read-back
  cudaStream_t stream = nullptr, comm_stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  // time state (SimParams: simtime, dt, last_dt, timestep, next_optime)
  double simtime = 0, dt = 0, last_dt = 1.e100, next_optime = 0;
  int timestep = 0;
  double FV_dt = 0, chyp = 0, cr = 0;
  bool ph_valid = true;      // false after a fused full step: Ph's interior is stale, P is the truth
  bool next_dt_valid = false;  // d_dtmin[0] holds min CellTimeStep(P) for the current P
  double bc_refval[10][PION_MAXVAR];
  long long launches = 0;
  // optional per-launch timing of the stage kernel (bench.py roofline leg)
  bool timing = false;
  bool timing_suspended = false;  // a split stage is timed as a whole by stage_and_bcs
  bool no_overlap = false;    // PION_B200_NO_OVERLAP=1: halo exchange after the whole stage (A/B tests)
  bool force_overlap = false; // PION_B200_OVERLAP=1: split the stage even with a single exchanged face
  bool force_gather = false;  // PION_B200_GATHER=1: run the gather kernel on the fused path too (A/B tests)
  std::vector<cudaEvent_t> tev;  // begin/end pairs
  std::vector<cudaEvent_t> tev_pool;  // recycled timing events (no cudaEventCreate inside a timed loop)
  const char* last_stage_kernel = "(no stage launched yet)";
  bool last_stage_split = false;  // the most recent fused stage ran as boundary shell + interior
  // multi-GPU
  ncclComm_t comm = nullptr;
  double* d_red = nullptr;  // 2 doubles for the dt all-reduce
  // microphysics (mp_only_cooling)
  CoolParams cool;
  double *d_tables = nullptr, *mp_dE = nullptr, *mp_dE2 = nullptr;  // cooling source of a stage (mp_dE2: the corrector's)
  const double* mp_dE_stage = nullptr;                              // the one the next fused stage adds
  long long mp_failures = 0;
  // stellar-wind internal boundary: cell list (device linear indices) and reference states [nvar][n]
  long wind_n = 0;
  long* d_wind_idx = nullptr;
  double* d_wind_val = nullptr;
  double *sendbuf[6] = {nullptr}, *recvbuf[6] = {nullptr};
  size_t halo_elems[6] = {0};
  // upload / download staging (two compact variables, allocated on first use) and its copy stream
  double* stage[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_packed[2] = {nullptr, nullptr};
  // TMA tensor maps over the two state arrays (3-D grids; the TMA sweep kernel, stage_sweep_tma.cuh)
  alignas(64) CUtensorMap tmapP[2], tmapPh[2];  // [stage order - 1]: the predictor and the corrector use different tiles
  bool have_tmap = false;
};

static inline int nblocks(long n, int block, int cap = 148 * 16) {
  long b = (n + block - 1) / block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

// ---------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------
extern "C" const char* pion_gpu_last_error(void) { return g_last_error.c_str(); }

// Tensor map of one state array for the TMA sweep kernel: a 4-D tensor (x, y, z, variable) over the
// pitched SoA layout of grid.cuh, box = one plane tile [nbase][TY+3][36].  The driver entry point is
// fetched through the runtime (no link-time dependency on libcuda).
static int make_state_tmap(const pion_gpu_ctx* c, double* base, int order, CUtensorMap* out) {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return 1;
    }
    enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  int cw, rh, nb, tx, ty;
  sweep_tma_box(c->cfg.eqntype, order, c->ntr, &cw, &rh, &nb, &tx, &ty);
  if ((c->g.xoff + c->g.nb[0]) % 2 || tx % 2) {  // every box must start on a 16-byte boundary in x
    set_error("TMA sweep: tile boxes would start on odd x offsets");
    return 1;
  }
  const GridD& g = c->g;
  nb += c->ntr;  // tracers ride along as extra tile variables
  const cuuint64_t dims[4] = {(cuuint64_t)g.sy, (cuuint64_t)g.NGa[1], (cuuint64_t)g.NGa[2], (cuuint64_t)nb};
  const cuuint64_t strides[3] = {(cuuint64_t)g.sy * 8, (cuuint64_t)g.sz * 8, (cuuint64_t)g.vs * 8};
  const cuuint32_t box[4] = {(cuuint32_t)cw, (cuuint32_t)rh, 1u, (cuuint32_t)nb};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 1;
  }
  return 0;
}

// launch KERNEL<EQ, EP_cooling> for the cooling function of this context (mp_only_cooling::Edot's cases)
#define PION_COOL_DISPATCH(mode, KERNEL, EQ, grid, block, smem, stream, args)          \
  switch (mode) {                                                                      \
    case 2: KERNEL<EQ, 2><<<grid, block, smem, stream>>>(args); break;                 \
    case 4: KERNEL<EQ, 4><<<grid, block, smem, stream>>>(args); break;                 \
    case 5: KERNEL<EQ, 5><<<grid, block, smem, stream>>>(args); break;                 \
    case 6: KERNEL<EQ, 6><<<grid, block, smem, stream>>>(args); break;                 \
    case 7: KERNEL<EQ, 7><<<grid, block, smem, stream>>>(args); break;                 \
    default: KERNEL<EQ, 8><<<grid, block, smem, stream>>>(args); break;                \
  }

static int check_config(const pion_gpu_config& c) {
  if (c.ndim < 1 || c.ndim > 3) { set_error("ndim must be 1..3"); return 1; }
  if (c.eqntype != PION_EQEUL && c.eqntype != PION_EQMHD && c.eqntype != PION_EQGLM) { set_error("unsupported eqntype"); return 1; }
  if (c.coord_sys != PION_COORD_CRT && c.coord_sys != PION_COORD_CYL && c.coord_sys != PION_COORD_SPH) { set_error("Bad Geometry in setup_grid()"); return 1; }
  // setup_fixed_grid.cpp:1133-1190 / solver constructors: axisymmetry is 2-D (z,R), spherical symmetry 1-D Euler
  if (c.coord_sys == PION_COORD_CYL && c.ndim != 2) { set_error("Cylindrical coordinates only implemented for 2d axial symmetry"); return 1; }
  if (c.coord_sys == PION_COORD_SPH && (c.ndim != 1 || c.eqntype != PION_EQEUL)) { set_error("Spherical coordinates only implemented for 1D Euler"); return 1; }
  if (c.coord_sys != PION_COORD_CRT && c.n_wind > 0) { set_error("stellar-wind boundary: only Cartesian grids are built"); return 1; }
  // 1, 2, 3: linear / exact / hybrid Riemann solvers (riemann.cpp) for the Euler equations; riemann_MHD only has the
  // linear solve (riemannMHD.cpp:176-183: modes 2 and 3 end in rep.error "MODE i: Don't know what to do")
  const bool euler_only = (c.solver == PION_FLUX_ROE_PV || c.solver == PION_FLUX_FVS || c.solver == PION_FLUX_RSEXACT || c.solver == PION_FLUX_RSHYBRID);
  if (c.solver != PION_FLUX_LF && c.solver != PION_FLUX_RSLINEAR && c.solver != PION_FLUX_ROE && c.solver != PION_FLUX_HLLD && c.solver != PION_FLUX_HLL && !euler_only) { set_error("solver must be 0 (Lax-Friedrichs), 1-3 (linear / exact / hybrid Riemann solver, Euler), 4 (Roe-CV), 5 (Roe-PV), 6 (FVS), 7 (HLLD) or 8 (HLL)"); return 1; }
  // solver_eqn_mhd_adi.cpp:132-198: the MHD solvers have no Roe-PV / FVS branch ("what sort of flux solver do you mean???")
  if (euler_only && c.eqntype != PION_EQEUL) { set_error("solvers 2, 3 (exact / hybrid Riemann solver: riemann_MHD only knows the linear solve), 5 (Roe-PV) and 6 (FVS) are for the Euler equations only"); return 1; }
  if (c.eqntype == PION_EQEUL && c.solver == PION_FLUX_HLLD) { set_error("HLLD needs MHD equations"); return 1; }
  if (c.artviscosity != 0 && c.artviscosity != 1 && c.artviscosity != 3 && c.artviscosity != 4) { set_error("artviscosity must be 0,1,3,4"); return 1; }
  if (!((c.spOOA == 1 && c.tmOOA == 1) || (c.spOOA == 2 && c.tmOOA == 2))) { set_error("Bad OOA requests; choose (1,1) or (2,2)"); return 1; }
  if (c.ntracer < 0 || c.ntracer > PION_MAXTR) { set_error("ntracer must be 0..4"); return 1; }
  int nb0 = (c.eqntype == PION_EQEUL) ? 5 : (c.eqntype == PION_EQMHD) ? 8 : 9;
  if (c.nvar != nb0 + c.ntracer) { set_error("nvar inconsistent with eqntype/ntracer"); return 1; }
  for (int d = 0; d < 2 * c.ndim; d++) {
    int t = c.bc[d];
    if (t == PION_BC_FIXED && c.ndim > 1 && d >= 2) { set_error("FIXED boundary on a Y/Z face: the reference's BC_assign_FIXED never terminates there (fixed_boundaries.cpp:50-58)"); return 1; }
    if (t != PION_BC_PERIODIC && t != PION_BC_OUTFLOW && t != PION_BC_INFLOW && t != PION_BC_REFLECTING &&
        t != PION_BC_FIXED && t != PION_BC_DMACH && t != PION_BC_ONEWAY_OUT && t != PION_BC_MPI) { set_error("unsupported boundary type"); return 1; }
    if (c.eqntype == PION_EQGLM && c.ndim == 1 && (t == PION_BC_OUTFLOW || t == PION_BC_ONEWAY_OUT)) { set_error("Psi outflow boundary condition doesn't work for 1D! (outflow_boundaries.cpp:57)"); return 1; }
  }
  // internal_bc[4] / bc_refval[6 + i] / wind[2] are fixed-size: reject counts that would index past them
  if (c.n_internal_bc < 0 || c.n_internal_bc > 4) { set_error("n_internal_bc must be 0..4"); return 1; }
  if (c.n_wind < 0 || c.n_wind > 2) { set_error("n_wind must be 0..2"); return 1; }
  if (c.min_timestep < 0.0) { set_error("min_timestep must be >= 0"); return 1; }
  for (int i = 0; i < c.n_internal_bc; i++) {
    if (c.internal_bc[i] == PION_BC_STWIND) {
      if (c.n_wind < 1 || c.n_wind > 2) { set_error("BC_assign_STWIND() No Sources! (n_wind must be 1 or 2)"); return 1; }
    } else if (c.internal_bc[i] != PION_BC_DMACH2) { set_error("unsupported internal boundary"); return 1; }
  }
  if (c.cooling) {
    // mp_only_cooling::Edot has cases 2, 4, 5, 6, 7, 8 (mp_only_cooling.cpp:383-420); every other flag -- DMcC (3)
    // included -- ends in rep.error("bad cooling flag") in the reference
    if (c.cooling != 2 && (c.cooling < 4 || c.cooling > 8)) { set_error("bad cooling flag in mp_only_cooling::Edot (EP_cooling must be 2, 4, 5, 6, 7 or 8)"); return 1; }
    if (c.cooling == 8 && (c.n_table < 2 || !c.table_T || !c.table_rrhp || !c.table_C_rrh || !c.table_C_ffhe || !c.table_C_fbdn || !c.table_C_cie)) {
      set_error("EP_cooling 8 needs the mp_only_cooling lookup tables (n_table, table_*)");
      return 1;
    }
    if (c.cooling >= 4 && c.cooling <= 7 && (c.n_spline < 3 || !c.spline_logT || !c.spline_logL)) {
      set_error("EP_cooling 4..7 need the knots of the cooling-curve spline (n_spline, spline_logT, spline_logL)");
      return 1;
    }
    if (c.mp_timestep_limit < 0 || c.mp_timestep_limit > 4) { set_error("Bad MP_timestep_limit"); return 1; }
  }
  return 0;
}

// BC_assign_STWIND + BC_assign_STWIND_add_cells2src + stellar_wind::add_cell +
// set_wind_cell_reference_state (boundaries/stellar_wind_boundaries.cpp:29-250,
// grid/stellar_wind_BC.cpp:125-596) for constant sources on a Cartesian grid: every cell
// (ghost cells included) whose centre lies within `radius` of the source joins the list,
// becomes !isdomain (mask = 0) and carries a fixed reference state.  Host code, run once.
static int build_wind_cells(pion_gpu_ctx* c) {
  const pion_gpu_config& cfg = c->cfg;
  const GridD& g = c->g;
  const double kB = 1.38064852e-16, m_p = 1.672621898e-24, Msun = 1.9891e33, year = 3.1558150e7;  // constants.h
  const double gamma = 5. / 3.;  // the literal add_cell passes (:339)
  const int nd = g.ndim, nv = c->nvar;
  std::vector<long> idx;
  std::vector<double> val;  // [cell][var] while building
  std::vector<unsigned char> mask((size_t)g.vs, 1);
  auto dpos = [&](int i, int a) { return cfg.xmin[a] + (2 * (i - g.nb[a]) + 1) * (0.5 * g.dx); };
  for (int id = 0; id < cfg.n_wind; id++) {
    const pion_gpu_wind_source& w = cfg.wind[id];
    const double Mdot = w.mdot * Msun / year, Vinf = w.vinf * 1.0e5, v_rot = w.vrot * 1.0e5;
    for (int k = 0; k < g.NGa[2]; k++)
      for (int j = 0; j < g.NGa[1]; j++)
        for (int i = 0; i < g.NGa[0]; i++) {
          const int ijk[3] = {i, j, k};
          double dist = 0.0;
          for (int a = 0; a < nd; a++) dist += pow(w.dpos[a] - dpos(ijk[a], a), 2.0);
          dist = sqrt(dist);
          if (!(dist <= w.radius)) continue;
          double p[PION_MAXVAR] = {0};
          bool set_rho = true;
          if (dist < 0.75 * w.radius && nd > 1) { p[0] = 1.0e-31; p[1] = 1.0e-31; set_rho = false; }
          if (nd == 2) {
            p[0] = Mdot / (Vinf * 2.0 * M_PI * dist);
            p[1] = kB * w.temp / m_p;
            p[1] *= exp((gamma - 1.0) * log(2.0 * M_PI * w.rstar * Vinf / Mdot));
            p[1] *= exp((gamma)*log(p[0]));
          } else if (set_rho) {
            p[0] = 1.0 / (dist);
            p[0] *= p[0];
            p[0] *= Mdot / (Vinf * 4.0 * M_PI);
            p[1] = kB * w.temp / m_p;
            p[1] *= exp((gamma - 1.0) * log(4.0 * M_PI * w.rstar * w.rstar * Vinf / Mdot));
            p[1] *= exp((gamma)*log(p[0]));
          }
          const double x = dpos(i, 0) - w.dpos[0], y = (nd > 1) ? dpos(j, 1) - w.dpos[1] : 0.0,
                       z = (nd > 2) ? dpos(k, 2) - w.dpos[2] : 0.0;
          const double d2 = exp(2 * log(dist));  // pconst.pow_fast(dist,2)
          if (nd == 1) {
            p[2] = Vinf * x / dist;
          } else if (nd == 2) {
            p[2] = Vinf * x / dist;
            p[3] = Vinf * y / dist;
            p[4] = v_rot * w.rstar * y / d2;
          } else {
            p[2] = Vinf * x / dist;
            p[3] = Vinf * y / dist;
            p[4] = Vinf * z / dist;
            p[2] += -v_rot * w.rstar * y / d2;
            p[3] += v_rot * w.rstar * x / d2;
          }
          if (cfg.eqntype != PION_EQEUL) {
            if (nd == 1) { set_error("1D spherical but MHD?"); return 1; }
            const double B_s = w.bsrf / sqrt(4.0 * M_PI), D_s = w.rstar / dist, D_2 = D_s * D_s;
            double beta = (v_rot / Vinf) * B_s * D_s;
            if (nd == 2) {
              p[5] = B_s * D_2 * fabs(x) / dist;
              p[6] = B_s * D_2 / dist;
              p[6] = (x > 0.0) ? y * p[6] : -y * p[6];
              beta = beta * y / dist;
              p[7] = (x > 0.0) ? -beta : beta;
            } else {
              p[5] = B_s * D_2 / dist;
              p[5] = (z > 0.0) ? x * p[5] : -x * p[5];
              p[6] = B_s * D_2 / dist;
              p[6] = (z > 0.0) ? y * p[6] : -y * p[6];
              p[7] = B_s * D_2 * fabs(z) / dist;
              beta *= sqrt(x * x + y * y) / dist;
              beta = (z > 0.0) ? -beta : beta;
              p[5] += -beta * y / dist;
              p[6] += beta * x / dist;
            }
          }
          if (cfg.eqntype == PION_EQGLM) p[8] = 0.0;
          for (int t = 0; t < c->ntr; t++) p[c->nbase_ + t] = w.tr[t];
          // SET_NEGATIVE_PRESSURE_TO_FIXED_TEMPERATURE (:583-594), Tmin = EP.MinTemperature
          if (cfg.cooling) {
            if (p[1] * c->pp.mu_tot_over_kB / p[0] < cfg.min_temperature) p[1] = p[0] * cfg.min_temperature / c->pp.mu_tot_over_kB;
          } else {
            p[1] = fmax(p[1], cfg.min_temperature * p[0] * kB * 0.78625 / m_p);
          }
          const long ci = gidx(g, i, j, k);
          idx.push_back(ci);
          mask[ci] = 0;
          for (int v = 0; v < nv; v++) val.push_back(p[v]);
        }
  }
  c->wind_n = (long)idx.size();
  if (c->wind_n == 0) return 0;
  std::vector<double> soa((size_t)nv * c->wind_n);
  for (long q = 0; q < c->wind_n; q++)
    for (int v = 0; v < nv; v++) soa[(size_t)v * c->wind_n + q] = val[(size_t)q * nv + v];
  CUDA_OK(cudaMalloc(&c->d_wind_idx, idx.size() * sizeof(long)));
  CUDA_OK(cudaMalloc(&c->d_wind_val, soa.size() * sizeof(double)));
  if (!c->mask) CUDA_OK(cudaMalloc(&c->mask, (size_t)g.vs));
  CUDA_OK(cudaMemcpy(c->d_wind_idx, idx.data(), idx.size() * sizeof(long), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(c->d_wind_val, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(c->mask, mask.data(), (size_t)g.vs, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" pion_gpu_ctx* pion_gpu_create(const pion_gpu_config* cfg) {
  if (!cfg) { set_error("null config"); return nullptr; }
  if (check_config(*cfg)) return nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: libpion_b200 has no CPU fallback");
    return nullptr;
  }
  if (cudaSetDevice(cfg->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
  pion_gpu_ctx* c = new pion_gpu_ctx();
  c->cfg = *cfg;
  // "Force Nbc=1 if using Lax-Friedrichs flux" (setup_fixed_grid.cpp:188-190): first order in space and time
  if (c->cfg.solver == PION_FLUX_LF) c->cfg.spOOA = c->cfg.tmOOA = 1;
  // ics/get_sim_info.cpp:452-468
  if (c->cfg.artviscosity == 0) c->cfg.etav = 0.0;
  if (c->cfg.artviscosity == 3) c->cfg.etav = 0.1;
  GridD& g = c->g;
  g.ndim = cfg->ndim;
  const int nbc = (c->cfg.spOOA == 2) ? 2 : 1;  // setup_fixed_grid.cpp:183-184
  for (int a = 0; a < 3; a++) {
    g.NG[a] = (a < g.ndim) ? cfg->NG[a] : 1;
    g.nb[a] = (a < g.ndim) ? nbc : 0;
    g.NGa[a] = g.NG[a] + 2 * g.nb[a];
  }
  g.xoff = 16 - g.nb[0];
  long pitch = ((long)g.xoff + g.NGa[0] + 15) / 16 * 16;
  g.sy = pitch;
  g.sz = pitch * g.NGa[1];
  g.vs = ((pitch * g.NGa[1] * g.NGa[2]) + 15) / 16 * 16;
  g.dx = (cfg->xmax[0] - cfg->xmin[0]) / g.NG[0];  // UniformGrid::set_cell_size
  g.coord = cfg->coord_sys;
  g.r0 = (cfg->coord_sys == PION_COORD_CYL) ? cfg->xmin[1] : cfg->xmin[0];
  c->nvar = cfg->nvar;
  c->ntr = cfg->ntracer;
  c->nbase_ = cfg->nvar - cfg->ntracer;
  c->arr_elems = (size_t)g.vs * c->nvar;
  PhysParams& pp = c->pp;
  pp.gamma = cfg->gamma;
  pp.etav = c->cfg.etav;
  pp.chyp = 0.0;
  pp.refvec_ro = cfg->refvec[0];
  // eqns_Euler::SetAvgState (eqns_hydro_adiabatic.cpp:439-453, riemann.cpp:171): the Riemann solver's three reference
  // velocities are a tenth of the sound speed of RefVec
  pp.rs_refvec[0] = cfg->refvec[0];
  pp.rs_refvec[1] = cfg->refvec[1];
  pp.rs_refvec[2] = pp.rs_refvec[3] = pp.rs_refvec[4] = 0.1 * sqrt(cfg->gamma * cfg->refvec[1] / cfg->refvec[0]);
  if (cfg->eqntype != PION_EQEUL) {
    // eqns_mhd_ideal::SetAvgState (eqns_mhd_adiabatic.cpp:501-543, called by the riemann_MHD constructor in direction XX):
    // reference velocity = a tenth of the fast speed of RefVec rotated about z so that B_y = 0, reference field = |B(RefVec)|
    // of the vector rotated there and back
    double rv[8];
    for (int v = 0; v < 8; v++) rv[v] = cfg->refvec[v];
    const double g = cfg->gamma;
    auto cfast = [&]() {
      const double ch = sqrt(g * rv[1] / rv[0]);
      const double t1 = ch * ch + (rv[5] * rv[5] + rv[6] * rv[6] + rv[7] * rv[7]) / rv[0];
      double t2 = 4. * ch * ch * rv[5] * rv[5] / rv[0];
      t2 = fmax(PION_MACHINEACCURACY, t1 * t1 - t2);
      return sqrt((t1 + sqrt(t2)) / 2.);
    };
    auto rot = [&](double th) {
      const double ct = cos(th), st = sin(th);
      double vx = rv[2] * ct - rv[3] * st, vy = rv[2] * st + rv[3] * ct;
      rv[2] = vx; rv[3] = vy;
      vx = rv[5] * ct - rv[6] * st; vy = rv[5] * st + rv[6] * ct;
      rv[5] = vx; rv[6] = vy;
    };
    double angle = rv[6] * rv[6] + rv[5] * rv[5], refvel;
    if (angle > 10. * PION_MACHINEACCURACY) {
      angle = M_PI / 2. - asin(rv[6] / sqrt(angle));
      if (rv[5] < 0) angle = -angle;
      rot(angle);
      refvel = cfast();
      rot(-angle);
    } else {
      refvel = cfast();
    }
    pp.rs_refvec[2] = 0.1 * refvel;
    pp.rs_refvec[3] = sqrt(rv[5] * rv[5] + rv[6] * rv[6] + rv[7] * rv[7]);
    pp.rs_refvec[4] = 0.0;
  }
  pp.min_temp = cfg->min_temperature;
  pp.max_temp = cfg->max_temperature;
  pp.have_mp = cfg->cooling ? 1 : 0;
  pp.mu_tot_over_kB = cfg->cooling ? (0.609 * 1.672621898e-24) / 1.38064852e-16 : 0.0;
  c->simtime = cfg->starttime;
  { const char* e = getenv("PION_B200_GATHER"); c->force_gather = e && e[0] == '1'; }
  { const char* e = getenv("PION_B200_NO_OVERLAP"); c->no_overlap = e && e[0] == '1'; }
  { const char* e = getenv("PION_B200_OVERLAP"); c->force_overlap = e && e[0] == '1'; }

  bool ok = true;
  ok &= cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
  {
    int plo = 0, phi = 0;  // the comm stream (boundary shell + halo exchange) outranks the interior update
    cudaDeviceGetStreamPriorityRange(&plo, &phi);
    ok &= cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, phi) == cudaSuccess;
  }
  ok &= cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming) == cudaSuccess;
  ok &= cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming) == cudaSuccess;
  ok &= cudaMalloc(&c->P, c->arr_elems * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&c->Ph, c->arr_elems * sizeof(double)) == cudaSuccess;
  ok &= cudaMalloc(&c->d_dtmin, 2 * sizeof(unsigned long long)) == cudaSuccess;
  ok &= cudaMalloc(&c->d_counters, 4 * sizeof(long long)) == cudaSuccess;
  ok &= cudaMallocHost(&c->h_pinned, 8 * sizeof(unsigned long long)) == cudaSuccess;
  if (cfg->solver == PION_FLUX_HLLD) ok &= cudaMalloc(&c->hll, (size_t)g.vs) == cudaSuccess;
  if (cfg->solver == PION_FLUX_ROE && (cfg->artviscosity == 3 || cfg->artviscosity == 4))
    ok &= cudaMalloc(&c->eta, (size_t)g.vs * 3 * sizeof(double)) == cudaSuccess;
  if (cfg->cooling) {
    // mp_only_cooling constructor + gen_mpoc_lookup_tables (mp_only_cooling.cpp:51-160, :528-579)
    const double m_p = 1.672621898e-24;  // constants.h:64
    const double Mu = 1.40 * m_p, Mu_elec = 1.167 * m_p, Mu_ion = 1.273 * m_p;
    CoolParams& cp = c->cool;
    cp.mode = cfg->cooling;
    cp.guess_ok = 0; cp.l2T0 = 0.f; cp.inv_l2step = 0.f;
    cp.nT = (cfg->cooling == 8) ? cfg->n_table : (cfg->cooling >= 4) ? cfg->n_spline : 0;
    cp.Mu = Mu; cp.Mu_elec = Mu_elec; cp.Mu_ion = Mu_ion;
    cp.smin = cfg->spline_min_slope; cp.smax = cfg->spline_max_slope;
    cp.inv_Mu2 = 1.0 / (Mu * Mu);
    cp.inv_Mu2_elec_H = 1.0 / (Mu_elec * Mu);
    cp.Mu_tot_over_kB = pp.mu_tot_over_kB;
    cp.MinT = cfg->min_temperature;
    cp.MaxT = cfg->max_temperature;
    if (cp.MinT < 1.0 || cp.MinT > 1.0e6) cp.MinT = 1.0;      // microphysics_base / mp_only_cooling limits
    if (cp.MaxT < 1.0e2 || cp.MaxT > 3.0e10) cp.MaxT = 1.0e8;
    const int n = cp.nT;
    std::vector<double> h((size_t)cool_ncol(cp.mode) * n + 1, 0.0);
    if (cp.mode == 8) {
      const double* src[6] = {cfg->table_T, cfg->table_rrhp, cfg->table_C_rrh, cfg->table_C_ffhe, cfg->table_C_fbdn, cfg->table_C_cie};
      for (int q = 0; q < 6; q++) memcpy(&h[(size_t)q * n], src[q], n * sizeof(double));
      for (int q = 1; q < 6; q++)
        for (int i = 0; i < n - 1; i++) h[(size_t)(5 + q) * n + i] = (h[(size_t)q * n + i + 1] - h[(size_t)q * n + i]) / (h[i + 1] - h[i]);
      // log-uniform T column (the reference builds T_i = T_0 (T_max/T_0)^(i/(n-1))): interval guess from log2 T.
      // Any strictly increasing table whose knots stay within a quarter step of that law qualifies; anything else
      // keeps the binary search.  (The guess is only a starting point: the kernel corrects it against the table.)
      cp.guess_ok = 0; cp.l2T0 = 0.f; cp.inv_l2step = 0.f;
      if (n >= 3 && h[0] > 0.0) {
        const double l0 = log2(h[0]), step = (log2(h[n - 1]) - l0) / (n - 1);
        bool ok_guess = step > 0.0;
        for (int i = 1; i < n && ok_guess; i++) ok_guess = h[i] > h[i - 1] && fabs(log2(h[i]) - (l0 + i * step)) < 0.25 * step;
        if (ok_guess) { cp.guess_ok = 1; cp.l2T0 = (float)l0; cp.inv_l2step = (float)(1.0 / step); }
      }
    } else if (cp.mode >= 4) {
      // natural cubic spline through the knots, as GSL's cspline builds it for the reference (tools/interpolate.cpp:
      // 59-118): c = y''/2 from the symmetric tridiagonal system, solved by forward elimination + back substitution
      double *x = &h[0], *y = &h[n], *cc = &h[2 * (size_t)n];
      memcpy(x, cfg->spline_logT, n * sizeof(double));
      memcpy(y, cfg->spline_logL, n * sizeof(double));
      const int m = n - 2;
      std::vector<double> diag(m), off(m), rhs(m);
      for (int i = 0; i < m; i++) {
        const double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
        off[i] = h_ip1;
        diag[i] = 2.0 * (h_ip1 + h_i);
        rhs[i] = 3.0 * ((y[i + 2] - y[i + 1]) / h_ip1 - (y[i + 1] - y[i]) / h_i);
      }
      for (int i = 1; i < m; i++) {
        const double w = off[i - 1] / diag[i - 1];
        diag[i] -= w * off[i - 1];
        rhs[i] -= w * rhs[i - 1];
      }
      cc[0] = cc[n - 1] = 0.0;
      cc[m] = rhs[m - 1] / diag[m - 1];
      for (int i = m - 1; i-- > 0;) cc[i + 1] = (rhs[i] - off[i] * cc[i + 2]) / diag[i];
    }
    ok &= cudaMalloc(&c->d_tables, h.size() * sizeof(double)) == cudaSuccess;
    ok &= cudaMalloc(&c->mp_dE, (size_t)g.vs * sizeof(double)) == cudaSuccess;
    if (cfg->tmOOA == 2) ok &= cudaMalloc(&c->mp_dE2, (size_t)g.vs * sizeof(double)) == cudaSuccess && cudaMemset(c->mp_dE2, 0, (size_t)g.vs * sizeof(double)) == cudaSuccess;
    if (ok) {
      ok &= cudaMemcpy(c->d_tables, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
      ok &= cudaMemset(c->mp_dE, 0, (size_t)g.vs * sizeof(double)) == cudaSuccess;
    }
    cp.tables = c->d_tables;
    c->cfg.table_T = c->cfg.table_rrhp = c->cfg.table_C_rrh = c->cfg.table_C_ffhe = c->cfg.table_C_fbdn = c->cfg.table_C_cie = nullptr;
    c->cfg.spline_logT = c->cfg.spline_logL = nullptr;
  }
  for (int ib = 0; ib < cfg->n_internal_bc; ib++) {
    if (cfg->internal_bc[ib] != PION_BC_STWIND) continue;
    if (build_wind_cells(c)) { pion_gpu_destroy(c); return nullptr; }
    break;  // one cell list covers every source (a second STWIND entry must not build it again)
  }
  if (!ok) {
    set_error(std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    pion_gpu_destroy(c);
    return nullptr;
  }
  {
    // 3-D Cartesian grids run the TMA sweep kernel (PION_B200_NO_TMA=1: the LDG sweep kernel, for A/B tests)
    const char* e = getenv("PION_B200_NO_TMA");
    if (g.ndim == 3 && g.coord == PION_COORD_CRT && !(e && e[0] == '1') && sweep_tma_fits(c->cfg.eqntype, c->ntr)) {
      for (int o = 1; o <= 2; o++)
        if (make_state_tmap(c, c->P, o, &c->tmapP[o - 1]) || make_state_tmap(c, c->Ph, o, &c->tmapPh[o - 1])) { pion_gpu_destroy(c); return nullptr; }
      c->have_tmap = true;
      if (c->hll && (cudaMalloc(&c->hllf, (size_t)g.vs) != cudaSuccess || cudaMemset(c->hllf, 0, (size_t)g.vs) != cudaSuccess)) {
        set_error("device allocation failed (face flags)");
        pion_gpu_destroy(c);
        return nullptr;
      }
    }
  }
  cudaMemsetAsync(c->P, 0, c->arr_elems * sizeof(double), c->stream);
  cudaMemsetAsync(c->Ph, 0, c->arr_elems * sizeof(double), c->stream);
  cudaMemsetAsync(c->d_counters, 0, 4 * sizeof(long long), c->stream);
  if (c->hll) cudaMemsetAsync(c->hll, 0, (size_t)g.vs, c->stream);
  if (c->eta) cudaMemsetAsync(c->eta, 0, (size_t)g.vs * 3 * sizeof(double), c->stream);
  cudaStreamSynchronize(c->stream);
  return c;
}

extern "C" void pion_gpu_destroy(pion_gpu_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (auto e : c->tev) cudaEventDestroy(e);
  for (auto e : c->tev_pool) cudaEventDestroy(e);
  if (c->comm) ncclCommDestroy(c->comm);
  cudaFree(c->P); cudaFree(c->Ph); cudaFree(c->dU); cudaFree(c->eta); cudaFree(c->hll); cudaFree(c->hllf); cudaFree(c->mask);
  cudaFree(c->d_dtmin); cudaFree(c->d_counters); cudaFree(c->d_red); cudaFree(c->d_tables); cudaFree(c->mp_dE); cudaFree(c->mp_dE2); cudaFree(c->d_wind_idx); cudaFree(c->d_wind_val);
  for (int f = 0; f < 6; f++) { cudaFree(c->sendbuf[f]); cudaFree(c->recvbuf[f]); }
  for (int q = 0; q < 2; q++) {
    cudaFree(c->stage[q]);
    if (c->ev_copied[q]) cudaEventDestroy(c->ev_copied[q]);
    if (c->ev_packed[q]) cudaEventDestroy(c->ev_packed[q]);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  delete c;
}

static int ensure_dU(pion_gpu_ctx* c) {
  if (c->dU) return 0;
  CUDA_OK(cudaMalloc(&c->dU, c->arr_elems * sizeof(double)));
  CUDA_OK(cudaMemsetAsync(c->dU, 0, c->arr_elems * sizeof(double), c->stream));
  return 0;
}

// ---------------------------------------------------------------------------
// upload / download: compact padded SoA on the host <-> pitched SoA on the device
// ---------------------------------------------------------------------------
static int ensure_staging(pion_gpu_ctx* c) {
  if (c->stage[0]) return 0;
  const size_t nv = (size_t)c->g.NGa[0] * c->g.NGa[1] * c->g.NGa[2];
  CUDA_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int q = 0; q < 2; q++) {
    CUDA_OK(cudaMalloc(&c->stage[q], nv * sizeof(double)));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_copied[q], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_packed[q], cudaEventDisableTiming));
  }
  return 0;
}

static double* state_ptr(pion_gpu_ctx* c, int which) {
  if (which == PION_STATE_P) return c->P;
  if (which == PION_STATE_PH) return c->ph_valid ? c->Ph : c->P;  // after a fused full step Ph == P
  return c->dU;
}

extern "C" int pion_gpu_upload(pion_gpu_ctx* c, int which, const double* soa) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (which == PION_STATE_DU && ensure_dU(c)) return 1;
  if (which == PION_STATE_PH && !c->ph_valid) {
    CUDA_OK(cudaMemcpyAsync(c->Ph, c->P, c->arr_elems * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->ph_valid = true;
  }
  double* dst = (which == PION_STATE_P) ? c->P : (which == PION_STATE_PH) ? c->Ph : c->dU;
  const GridD& g = c->g;
  const size_t rows = (size_t)g.NGa[1] * g.NGa[2];
  if (ensure_staging(c)) return 1;
  const size_t nv = rows * g.NGa[0];
  for (int v = 0; v < c->nvar; v++) {
    const int q = v & 1;
    // flat PCIe copy of variable v into staging buffer q (once the re-pitch of variable v-2 has drained it),
    // then the re-pitch kernel on the compute stream: the copy of v+1 overlaps the re-pitch of v
    if (v >= 2) CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_packed[q], 0));
    CUDA_OK(cudaMemcpyAsync(c->stage[q], soa + (size_t)v * nv, nv * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
    CUDA_OK(cudaEventRecord(c->ev_copied[q], c->copy_stream));
    CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copied[q], 0));
    k_repack_var<<<nblocks((long)rows, 1, 148 * 32), 256, 0, c->stream>>>(g, c->stage[q], dst + (size_t)v * g.vs, 1);
    CUDA_OK(cudaEventRecord(c->ev_packed[q], c->stream));
  }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  if (which == PION_STATE_P) c->next_dt_valid = false;
  return 0;
}

extern "C" int pion_gpu_download(pion_gpu_ctx* c, int which, double* soa) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (which == PION_STATE_DU && ensure_dU(c)) return 1;
  const double* src = state_ptr(c, which);
  const GridD& g = c->g;
  const size_t rows = (size_t)g.NGa[1] * g.NGa[2];
  if (ensure_staging(c)) return 1;
  const size_t nv = rows * g.NGa[0];
  for (int v = 0; v < c->nvar; v++) {
    const int q = v & 1;
    // pack variable v into staging buffer q (once the copy of variable v-2 has left it), flat PCIe copy out
    if (v >= 2) CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copied[q], 0));
    k_repack_var<<<nblocks((long)rows, 1, 148 * 32), 256, 0, c->stream>>>(g, c->stage[q], const_cast<double*>(src) + (size_t)v * g.vs, 0);
    CUDA_OK(cudaEventRecord(c->ev_packed[q], c->stream));
    CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->ev_packed[q], 0));
    CUDA_OK(cudaMemcpyAsync(soa + (size_t)v * nv, c->stage[q], nv * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    CUDA_OK(cudaEventRecord(c->ev_copied[q], c->copy_stream));
  }
  CUDA_OK(cudaStreamSynchronize(c->copy_stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---------------------------------------------------------------------------
// boundaries
// ---------------------------------------------------------------------------
static void fill_bc_args(pion_gpu_ctx* c, BCArgs& b, int face, int type, double* A0, double* A1, double simtime,
                         const double* refval) {
  b.g = c->g;
  b.A[0] = A0;
  b.A[1] = A1;
  b.narr = A1 ? 2 : 1;
  b.face = face;
  b.type = type;
  b.nvar = c->nvar;
  b.eq = c->cfg.eqntype;
  b.ftr = c->nbase_;
  for (int v = 0; v < PION_MAXVAR; v++) b.refval[v] = refval ? refval[v] : 0.0;
  b.simtime = simtime;
  for (int a = 0; a < 3; a++) b.sim_xmin[a] = c->cfg.xmin[a];  // local origin: positions are xmin + (2i+1)dx/2
}

static long face_cells(const GridD& g, int face) {
  long n = g.nb[face >> 1];
  for (int q = 0; q < 3; q++)
    if (q != (face >> 1)) n *= g.NGa[q];
  return n;
}

static int halo_exchange_axis(pion_gpu_ctx* c, int ax, double* A, cudaStream_t st);

// TimeUpdateInternalBCs (assign_update_bcs.cpp:134-181): of the internal boundaries only STWIND is updated
// here -- BC_update_STWIND writes P and Ph (stellar_wind_boundaries.cpp:300-341)
static int update_internal_bcs(pion_gpu_ctx* c, cudaStream_t st) {
  if (c->wind_n) {
    k_wind_set<<<nblocks(c->wind_n * c->nvar, 256), 256, 0, st>>>(c->g.vs, c->nvar, c->wind_n, c->d_wind_idx, c->d_wind_val, c->P, c->Ph);
    c->launches++;
  }
  return 0;
}

// TimeUpdateExternalBCs (assign_update_bcs.cpp:191-246) on the given arrays (A1 may be null): the six faces in
// BC_bd order, then DMACH2 (an "internal" boundary by position, but the reference updates it in this call)
static int update_external_bcs(pion_gpu_ctx* c, double* A0, double* A1, double simtime, cudaStream_t st) {
  const GridD& g = c->g;
  for (int ax = 0; ax < g.ndim; ax++) {
    const int tlo = c->cfg.bc[2 * ax], thi = c->cfg.bc[2 * ax + 1];
    const bool mpi_face = (tlo == PION_BC_MPI || thi == PION_BC_MPI);
    if (tlo != PION_BC_MPI || thi != PION_BC_MPI) {  // at least one physical face: both in ONE launch
      BCArgs b;
      fill_bc_args(c, b, 2 * ax, tlo, A0, A1, simtime, c->bc_refval[2 * ax]);
      BCRef r2;
      for (int v = 0; v < PION_MAXVAR; v++) r2.v[v] = c->bc_refval[2 * ax + 1][v];
      k_bc_axis<<<nblocks(2 * face_cells(g, 2 * ax), 128), 128, 0, st>>>(b, thi, r2);
      c->launches++;
    }
    if (mpi_face) {
      if (halo_exchange_axis(c, ax, A0, st)) return 1;
      if (A1 && halo_exchange_axis(c, ax, A1, st)) return 1;
    }
  }
  for (int i = 0; i < c->cfg.n_internal_bc; i++) {
    if (c->cfg.internal_bc[i] == PION_BC_DMACH2) {
      BCArgs b;
      fill_bc_args(c, b, 2, PION_BC_DMACH2, A0, A1, simtime, c->bc_refval[6 + i]);
      k_bc_dmach2<<<nblocks((long)g.NG[0] * g.nb[1], 128), 128, 0, st>>>(b);
      c->launches++;
    }
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

// TimeUpdateInternalBCs + TimeUpdateExternalBCs, as every caller in the reference issues them back to back
static int update_bcs_arrays(pion_gpu_ctx* c, double* A0, double* A1, double simtime, cudaStream_t st = nullptr) {
  if (!st) st = c->stream;
  if (update_internal_bcs(c, st)) return 1;
  return update_external_bcs(c, A0, A1, simtime, st);
}

static int ensure_ph_valid(pion_gpu_ctx* c) {
  if (!c->ph_valid) {  // make Ph a true copy before it is used as a separate array again
    CUDA_OK(cudaMemcpyAsync(c->Ph, c->P, c->arr_elems * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->ph_valid = true;
  }
  return 0;
}

extern "C" int pion_gpu_time_update_bcs(pion_gpu_ctx* c, double simtime, int cstep, int maxstep) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (ensure_ph_valid(c)) return 1;
  return update_bcs_arrays(c, c->Ph, (cstep == maxstep) ? c->P : nullptr, simtime);
}
extern "C" int pion_gpu_time_update_internal_bcs(pion_gpu_ctx* c, double simtime, int cstep, int maxstep) {
  (void)simtime; (void)cstep; (void)maxstep;
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (ensure_ph_valid(c)) return 1;
  if (update_internal_bcs(c, c->stream)) return 1;
  CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int pion_gpu_time_update_external_bcs(pion_gpu_ctx* c, double simtime, int cstep, int maxstep) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (ensure_ph_valid(c)) return 1;
  return update_external_bcs(c, c->Ph, (cstep == maxstep) ? c->P : nullptr, simtime, c->stream);
}

// one device->host read of `n` doubles starting at element `idx` of variable planes of A
static int fetch_cell(pion_gpu_ctx* c, const double* A, long cidx_, double* out) {
  for (int v = 0; v < c->nvar; v++)
    CUDA_OK(cudaMemcpyAsync(out + v, A + (size_t)v * c->g.vs + cidx_, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

// sim_init::Init after ReadData (sim_init.cpp:215-262)
extern "C" int pion_gpu_init_after_upload(pion_gpu_ctx* c) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const GridD& g = c->g;
  const long ncell = (long)g.NG[0] * g.NG[1] * g.NG[2];
  // Ph = P on the grid; psi = 0 at step 0 for GLM (:215-241)
  const int zero_var = (c->cfg.eqntype == PION_EQGLM && c->timestep == 0) ? 8 : -1;
  if (zero_var >= 0) {
    k_zero_var_interior<<<nblocks(ncell, 256), 256, 0, c->stream>>>(g, c->P, zero_var);
    c->launches++;
  }
  k_copy_interior<<<nblocks(ncell, 256), 256, 0, c->stream>>>(g, c->P, c->Ph, c->nvar, -1);
  c->launches++;
  c->ph_valid = true;
  // assign_boundary_data (assign_update_bcs.cpp:28-120), face by face in list order;
  // reference values of INFLOW / FIXED / REFLECTING / DMACH are fixed here.
  for (int face = 0; face < 2 * g.ndim; face++) {
    const int type = c->cfg.bc[face];
    const int ax = face >> 1, pos = face & 1;
    double* rv = c->bc_refval[face];
    for (int v = 0; v < PION_MAXVAR; v++) rv[v] = 0.0;
    if (type == PION_BC_MPI) continue;
    if (type == PION_BC_REFLECTING) {  // reflecting_boundaries.cpp:36-75
      for (int v = 0; v < c->nvar; v++) rv[v] = 1.0;
      rv[2 + ax] = -1.0;
      if (c->cfg.eqntype != PION_EQEUL) rv[5 + ax] = -1.0;
    } else if (type == PION_BC_DMACH) {  // double_Mach_ref_boundaries.cpp:33-38
      rv[0] = 1.4; rv[1] = 1.0; rv[2] = rv[3] = rv[4] = 0.0;
      for (int v = c->nbase_; v < c->nvar; v++) rv[v] = -1.0;
    } else if (type == PION_BC_INFLOW || type == PION_BC_FIXED) {
      // INFLOW: P of the source of the LAST ghost cell in the list (inflow_boundaries.cpp:36-52);
      // FIXED: P of the source of the FIRST ghost cell (fixed_boundaries.cpp:45-63).
      int lo[3], hi[3];
      for (int q = 0; q < 3; q++) { lo[q] = 0; hi[q] = g.NGa[q]; }
      if (ax == 0) { for (int q = 1; q < 3; q++) { lo[q] = g.nb[q]; hi[q] = g.NGa[q] - g.nb[q]; } }
      else if (ax == 1) { lo[2] = g.nb[2]; hi[2] = g.NGa[2] - g.nb[2]; }
      int ijk[3];
      for (int q = 0; q < 3; q++) ijk[q] = (type == PION_BC_INFLOW) ? hi[q] - 1 : lo[q];
      ijk[ax] = pos ? g.NGa[ax] - g.nb[ax] - 1 : g.nb[ax];  // the edge cell
      if (fetch_cell(c, c->P, gidx(g, ijk[0], ijk[1], ijk[2]), rv)) return 1;
    }
    BCArgs b;
    // BC_assign_ONEWAY_OUT is BC_assign_OUTFLOW: no velocity clamp at assign time
    // (oneway_out_boundaries.cpp:24-32)
    fill_bc_args(c, b, face, (type == PION_BC_ONEWAY_OUT) ? PION_BC_OUTFLOW : type, c->P, c->Ph, c->simtime, rv);
    k_bc_face<<<nblocks(face_cells(g, face), 128), 128, 0, c->stream>>>(b);
    c->launches++;
  }
  for (int i = 0; i < c->cfg.n_internal_bc; i++) {
    double* rv = c->bc_refval[6 + i];
    for (int v = 0; v < PION_MAXVAR; v++) rv[v] = 0.0;
    if (c->cfg.internal_bc[i] == PION_BC_DMACH2) {  // double_Mach_ref_boundaries.cpp:104-110
      rv[0] = 8.0; rv[1] = 116.5; rv[2] = 7.14470958; rv[3] = -4.125; rv[4] = 0.0;
      for (int v = c->nbase_; v < c->nvar; v++) rv[v] = 1.0;
    } else if (c->cfg.internal_bc[i] != PION_BC_STWIND) {
      set_error("unsupported internal boundary");
      return 1;
    }
  }
  // first TimeUpdateInternal/ExternalBCs (sim_init.cpp:259-262), cstep==maxstep
  if (update_bcs_arrays(c, c->Ph, c->P, c->simtime)) return 1;
  if (c->cfg.op_criterion == 1) {  // sim_init.cpp:270-280
    c->next_optime = c->simtime + c->cfg.opfreq_time;
    double tmp = ((c->simtime / c->cfg.opfreq_time) - floor(c->simtime / c->cfg.opfreq_time)) * c->cfg.opfreq_time;
    c->next_optime -= tmp;
  }
  c->next_dt_valid = false;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---------------------------------------------------------------------------
// time step
// ---------------------------------------------------------------------------
static const unsigned long long DT_INIT_BITS = 0x54B249AD2594C37DULL;  // bits of 1.0e100

static int launch_calc_dt(pion_gpu_ctx* c) {
  const GridD& g = c->g;
  const long ncell = (long)g.NG[0] * g.NG[1] * g.NG[2];
  CUDA_OK(cudaMemcpyAsync(c->d_dtmin, &DT_INIT_BITS, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
  const int blocks = nblocks(ncell, 256, 148 * 8);
  switch (c->cfg.eqntype) {
    case PION_EQEUL: k_calc_dt<EQ_EULER><<<blocks, 256, 0, c->stream>>>(g, c->pp, c->P, c->mask, c->cfg.cfl, c->d_dtmin); break;
    case PION_EQMHD: k_calc_dt<EQ_MHD><<<blocks, 256, 0, c->stream>>>(g, c->pp, c->P, c->mask, c->cfg.cfl, c->d_dtmin); break;
    default: k_calc_dt<EQ_GLM><<<blocks, 256, 0, c->stream>>>(g, c->pp, c->P, c->mask, c->cfg.cfl, c->d_dtmin); break;
  }
  c->launches++;
  CUDA_OK(cudaGetLastError());
  c->next_dt_valid = true;
  return 0;
}

// Queues the kernels that leave the LOCAL minima in d_dtmin[0] (t_dyn) and d_dtmin[1] (t_mp, 1e99 without a
// microphysics limit) -- no host synchronisation.  A launch failure is returned, not thrown: the multi-rank
// caller must still enter the collective.
static const unsigned long long MP_INIT_BITS = 0x547D42AEA2879F2EULL;  // bits of 1.0e99
static int queue_local_dt(pion_gpu_ctx* c) {
  if (!c->next_dt_valid && launch_calc_dt(c)) return 1;
  // calc_microphysics_dt without MP / without a limit returns 1e99 (calc_timestep.cpp:348-357)
  CUDA_OK(cudaMemcpyAsync(c->d_dtmin + 1, &MP_INIT_BITS, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
  const int lim = c->cfg.mp_timestep_limit;
  if (c->cfg.cooling && lim >= 1 && lim <= 3) {  // 4 = recombination time only: none for mp_only_cooling
    CoolArgs a;
    a.g = c->g; a.cp = c->cool; a.P = c->ph_valid ? c->Ph : c->P;  // timescales(c->Ph)
    a.dE = nullptr; a.dU = nullptr; a.mask = c->mask; a.dt = 0.0; a.gamma = c->pp.gamma; a.counters = c->d_counters;
    a.dtmin = c->d_dtmin + 1; a.mp_timestep_limit = lim;
    const long ncell = (long)c->g.NG[0] * c->g.NG[1] * c->g.NG[2];
    const size_t smem = cool_smem_bytes(c->cool);
    // (the kernel reads rho and p only: one instantiation per cooling function)
    PION_COOL_DISPATCH(c->cool.mode, k_mp_dt, EQ_EULER, nblocks(ncell, 256, 148 * 8), 256, smem, c->stream, a)
    c->launches++;
    CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// One read-back of two doubles from `src` (device) through the pinned mirror.
static int read_two(pion_gpu_ctx* c, const void* src, double* a, double* b) {
  CUDA_OK(cudaMemcpyAsync(c->h_pinned, src, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  memcpy(a, c->h_pinned, sizeof(double));
  memcpy(b, c->h_pinned + 1, sizeof(double));
  return 0;
}

// first step with stellar winds: limit dt by the wind speed (calc_timestep.cpp:318-323); the source list is
// global, so every rank applies the same limit
static double wind_dt_limit(const pion_gpu_ctx* c, double d) {
  if (c->timestep == 0)
    for (int v = 0; v < c->cfg.n_wind; v++) d = fmin(d, 0.1 * c->cfg.cfl * c->g.dx / (c->cfg.wind[v].vinf * 1.0e5));
  return d;
}

extern "C" int pion_gpu_calc_dt(pion_gpu_ctx* c, double* t_dyn, double* t_mp) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (queue_local_dt(c)) return 1;
  double d, m;
  if (read_two(c, c->d_dtmin, &d, &m)) return 1;
  if (t_dyn) *t_dyn = wind_dt_limit(c, d);
  if (t_mp) {
    *t_mp = m;
    if (!(m > 0.0)) { set_error("get_mp_timescales_no_radiation() returned error"); return 1; }
  }
  return 0;
}

extern "C" int pion_gpu_calculate_timestep(pion_gpu_ctx* c, double* dt_out) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  double t_dyn, t_mp;
  if (c->comm) {
    // sim_control_MPI.cpp:503-504: global MIN of t_dyn and t_mp.  The local minima never leave the device: one
    // 2-element ncclAllReduce(min) straight on the kernels' result words (positive doubles order like their bit
    // patterns, so the atomicMin words ARE the doubles), then ONE 16-byte read-back.  A rank whose local
    // kernels failed still enters the collective -- with -1, which wins the MIN -- so that its peers fail with
    // it instead of blocking in the all-reduce for ever.
    const int local_err = queue_local_dt(c);
    if (local_err) {
      static const double bad[2] = {-1.0, -1.0};
      cudaMemcpyAsync(c->d_dtmin, bad, sizeof(bad), cudaMemcpyHostToDevice, c->stream);
    }
    NCCL_OK(ncclAllReduce(c->d_dtmin, c->d_red, 2, ncclDouble, ncclMin, c->comm, c->stream));
    if (read_two(c, c->d_red, &t_dyn, &t_mp)) return 1;
    if (local_err) return 1;
    if (t_dyn < 0.0) { set_error("calculate_timestep failed on another rank"); return 1; }
    t_dyn = wind_dt_limit(c, t_dyn);
  } else {
    if (pion_gpu_calc_dt(c, &t_dyn, &t_mp)) return 1;
  }
  if (!(t_mp > 0.0)) { set_error("get_mp_timescales_no_radiation() returned error"); return 1; }
  if (!(t_dyn > 0.0)) { set_error("CellTimeStep function returned failing value"); return 1; }
  c->dt = fmin(t_dyn, t_mp);
  // Set_GLM_Speeds(td, dx, cr): c_h = CFL*dx/t_dyn, c_r = 0.25/dx (calc_timestep.cpp:121-131)
  if (c->cfg.eqntype == PION_EQGLM) {
    c->chyp = c->cfg.cfl * c->g.dx / t_dyn;
    c->cr = 0.25 / c->g.dx;
  }
  // timestep_checking_and_limiting (:219-262)
  if (c->dt < c->cfg.min_timestep) { set_error("Timestep too short! dt=" + std::to_string(c->dt) + "  min-step=" + std::to_string(c->cfg.min_timestep)); return 1; }
  c->dt = fmin(c->dt, 1.3 * c->last_dt);
  if (c->cfg.op_criterion == 1) {
    c->dt = fmin(c->dt, c->next_optime - c->simtime);
    if (c->dt <= 0.0) { set_error("Went past output time without outputting!"); return 1; }
  }
  c->dt = fmin(c->dt, c->cfg.finishtime - c->simtime);
  if (c->dt <= 0.0) { set_error("Negative timestep!"); return 1; }
  c->FV_dt = c->dt;
  if (dt_out) *dt_out = c->dt;
  return 0;
}

// constants::equalD (constants.cpp:48-69)
static bool host_equalD(double a, double b) {
  if (a == b) return true;
  if (fabs(a) + fabs(b) < 1.0e-100) return true;
  return (fabs(a - b) / (fabs(a) + fabs(b) + 1.0e-100)) < 1.0e-12;
}

// The output-criterion part of sim_init::output_data (sim_init.cpp:711-744): is the current step one that the
// caller should save?  With op_criterion == 1 an output time that has been reached is consumed here
// (next_optime += opfreq_time), exactly where the reference does it -- without this the dt limiter of
// calculate_timestep would see next_optime - simtime == 0 on the following step.
extern "C" int pion_gpu_output_due(pion_gpu_ctx* c, int opfreq, int* due) {
  int d = 1;
  const bool maxtime = c->simtime >= c->cfg.finishtime;
  if (c->timestep == 0) {
  } else if (c->cfg.op_criterion == 0) {
    if (opfreq == 0 && !maxtime) d = 0;
    else if (!maxtime && opfreq != 0 && (c->timestep % opfreq) != 0) d = 0;
  } else if (c->cfg.op_criterion == 1) {
    if (!host_equalD(c->simtime, c->next_optime) && !maxtime) d = 0;
    else c->next_optime += c->cfg.opfreq_time;
  } else {
    set_error("op_criterion must be 0 or 1");
    return 1;
  }
  if (due) *due = d;
  return 0;
}

extern "C" int pion_gpu_set_dt(pion_gpu_ctx* c, double dt) {
  c->dt = dt;
  c->FV_dt = dt;
  return 0;
}
extern "C" int pion_gpu_set_glm_speeds(pion_gpu_ctx* c, double t_dyn, double dx, double cr) {
  c->chyp = c->cfg.cfl * dx / t_dyn;  // solver_eqn_mhd_adi.cpp:916
  c->cr = cr;
  return 0;
}
extern "C" int pion_gpu_set_time(pion_gpu_ctx* c, double simtime, double last_dt, int timestep) {
  c->simtime = simtime;
  c->last_dt = last_dt;
  c->timestep = timestep;
  return 0;
}
extern "C" int pion_gpu_get_time(pion_gpu_ctx* c, double* simtime, double* dt, double* last_dt, int* timestep) {
  if (simtime) *simtime = c->simtime;
  if (dt) *dt = c->dt;
  if (last_dt) *last_dt = c->last_dt;
  if (timestep) *timestep = c->timestep;
  return 0;
}

// ---------------------------------------------------------------------------
// dynamics
// ---------------------------------------------------------------------------
static int launch_preprocess(pion_gpu_ctx* c, const double* S, int order) {
  const GridD& g = c->g;
  if (c->hll) {  // solver_eqn_base.cpp:398-412
    if (g.ndim == 3 && g.coord == PION_COORD_CRT) {
      const int ex = g.NGa[0] - 2, ey = g.NGa[1] - 2, ez = g.NGa[2] - 2;
      int kchunk = 64;
      const int bx = (ex + 31) / 32, by = (ey + 7) / 8;
      while (kchunk > 8 && (long)bx * by * ((ez + kchunk - 1) / kchunk) < 148L * 8) kchunk >>= 1;
      k_hlld_flags_3d<<<dim3(bx, by, (ez + kchunk - 1) / kchunk), 256, 0, c->stream>>>(g, S, c->hll, kchunk);
      if (c->hllf) {  // face form for the TMA sweep kernel
        const long nf = (g.sy / 4) * (long)(g.NGa[1] - 1) * (g.NGa[2] - 1);
        k_hll_face_flags<<<nblocks(nf, 256, 148 * 64), 256, 0, c->stream>>>(g, c->hll, c->hllf);
        c->launches++;
      }
    } else {
      long n = (long)(g.NGa[0] - 2) * ((g.ndim > 1) ? g.NGa[1] - 2 : 1) * ((g.ndim > 2) ? g.NGa[2] - 2 : 1);
      k_hlld_flags<<<nblocks(n, 256), 256, 0, c->stream>>>(g, S, c->hll);
    }
    c->launches++;
  }
  if (c->eta) {  // solver_eqn_base.cpp:423-573
    long n = (long)g.NGa[0] * g.NGa[1] * g.NGa[2];
    const double tiny2 = PION_VERY_TINY_VALUE * g.dx * g.dx;
    switch (c->cfg.eqntype) {
      case PION_EQEUL: k_hcorr_eta<EQ_EULER><<<nblocks(n, 128), 128, 0, c->stream>>>(g, S, c->eta, order, c->pp.gamma, tiny2); break;
      case PION_EQMHD: k_hcorr_eta<EQ_MHD><<<nblocks(n, 128), 128, 0, c->stream>>>(g, S, c->eta, order, c->pp.gamma, tiny2); break;
      default: k_hcorr_eta<EQ_GLM><<<nblocks(n, 128), 128, 0, c->stream>>>(g, S, c->eta, order, c->pp.gamma, tiny2); break;
    }
    c->launches++;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

struct StageBox { int tx0, tx1, ty0, ty1, k_lo, k_hi; };

// a timing event from the context's pool (created on first use, recycled by pion_gpu_stage_timing)
static int timing_event(pion_gpu_ctx* c, cudaEvent_t* e) {
  if (!c->tev_pool.empty()) {
    *e = c->tev_pool.back();
    c->tev_pool.pop_back();
    return 0;
  }
  CUDA_OK(cudaEventCreate(e));
  return 0;
}

// cells per tile of the sweep kernel that a fused stage of this context runs (launch_sweep_any's choice)
static void stage_tile_cells(const pion_gpu_ctx* c, int order, int* cx, int* cy) {
  sweep_tile_cells(c->cfg.eqntype, cx, cy);
  if (c->have_tmap && !c->eta) {
    int cw, rh, nb;
    sweep_tma_box(c->cfg.eqntype, order, c->ntr, &cw, &rh, &nb, cx, cy);
  }
}

// one stage = one launch over the whole grid (box == nullptr), or one launch per box when the stage is
// split into boundary shell + interior (only the first launch of a stage resets the dt minimum)
static int launch_stage(pion_gpu_ctx* c, const double* S, const double* Pb, double* out, double* dU, double dt, int order,
                        bool fused, bool want_dt, const StageBox* box = nullptr, int nbox = 0, bool first_box = true,
                        cudaStream_t st = nullptr) {
  if (!st) st = c->stream;
  StageArgs a;
  a.g = c->g;
  a.pp = c->pp;
  a.pp.chyp = c->chyp;
  a.pp.lf_c = c->g.dx / dt;  // get_LaxFriedrichs_flux: dx / FV_dt
  a.pp.lf_ndim = (double)c->g.ndim;
  a.S = S;
  a.Pb = Pb;
  a.out = out;
  a.dU = dU;
  a.mp_dE = (fused && c->cfg.cooling) ? c->mp_dE_stage : nullptr;
  a.hll = c->hll;
  a.hllf = c->hllf;
  a.eta = c->eta;
  a.mask = c->mask;
  a.dt = dt;
  a.idx = 1.0 / c->g.dx;
  a.dtdx = dt * a.idx;
  a.hdtdx = 0.5 * dt * a.idx;
  a.tiny2 = PION_VERY_TINY_VALUE * c->g.dx * c->g.dx;
  a.glm_damp = exp(-dt * c->chyp * c->cr);  // eqns_mhd_adiabatic.cpp:650-660 with FV_dt
  a.cfl = c->cfg.cfl;
  a.dtmin = want_dt ? c->d_dtmin : nullptr;
  a.counters = c->d_counters;
  a.order = order;
  a.ntr = c->ntr;
  a.fused = fused ? 1 : 0;
  const int fkj = (c->cfg.artviscosity == 1 || c->cfg.artviscosity == 4) ? 1 : 0;
  a.fkj = fkj;
  const int oi = (order == 2) ? 1 : 0;
  a.tmap = !c->have_tmap ? nullptr : (S == c->P) ? (const void*)&c->tmapP[oi] : (S == c->Ph) ? (const void*)&c->tmapPh[oi] : nullptr;
  {
    int cx, cy;
    stage_tile_cells(c, order, &cx, &cy);
    a.tx0 = 0; a.tx1 = (c->g.NG[0] + cx - 1) / cx;
    a.ty0 = 0; a.ty1 = (c->g.NG[1] + cy - 1) / cy;
    a.k_lo = 0; a.k_hi = c->g.NG[2];
    a.nbox = 0;
    if (box && nbox == 1) { a.tx0 = box->tx0; a.tx1 = box->tx1; a.ty0 = box->ty0; a.ty1 = box->ty1; a.k_lo = box->k_lo; a.k_hi = box->k_hi; }
    if (box && nbox > 1) {  // several boxes, ONE launch (the sweep kernels decode their box from a table)
      a.nbox = nbox;
      for (int q = 0; q < nbox; q++) {
        a.box[q][0] = box[q].tx0; a.box[q][1] = box[q].tx1; a.box[q][2] = box[q].ty0; a.box[q][3] = box[q].ty1;
        a.box[q][4] = box[q].k_lo; a.box[q][5] = box[q].k_hi;
      }
    }
  }
  if (want_dt && first_box)
    CUDA_OK(cudaMemcpyAsync(c->d_dtmin, &DT_INIT_BITS, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = c->timing && !c->timing_suspended;
  if (timed) {
    if (timing_event(c, &e0) || timing_event(c, &e1)) return 1;
    CUDA_OK(cudaEventRecord(e0, st));
  }
  // fused 2-D/3-D stages run the flux-once sweep kernel; 1-D grids and the unfused seam
  // call (calc_dynamics_dU) run the per-cell gather kernel
  // (Lax-Friedrichs and the linear / exact / hybrid Riemann solvers are only instantiated for the gather kernel)
  const bool sweep = fused && c->g.ndim >= 2 && c->g.coord == PION_COORD_CRT && !c->force_gather && c->cfg.solver >= PION_FLUX_ROE;
  switch (c->cfg.eqntype) {
    case PION_EQEUL: c->last_stage_kernel = (sweep ? launch_sweep_euler : launch_stage_euler)(c->cfg.solver, fkj, a, st); break;
    case PION_EQMHD: c->last_stage_kernel = (sweep ? launch_sweep_mhd : launch_stage_mhd)(c->cfg.solver, fkj, a, st); break;
    default: c->last_stage_kernel = (sweep ? launch_sweep_glm : launch_stage_glm)(c->cfg.solver, fkj, a, st); break;
  }
  if (timed) {
    CUDA_OK(cudaEventRecord(e1, st));
    c->tev.push_back(e0);
    c->tev.push_back(e1);
  }
  c->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// calc_noRT_microphysics_dU (time_integrator.cpp:438-489): always integrates from P
static int launch_cooling(pion_gpu_ctx* c, double dt, double* dE, double* dU, double dt2 = 0.0, double* dE2 = nullptr) {
  CoolArgs a;
  a.g = c->g; a.cp = c->cool; a.P = c->P; a.dE = dE; a.dU = dU; a.mask = c->mask; a.dt = dt; a.gamma = c->pp.gamma;
  a.dE2 = dE2; a.dt2 = dt2;
  a.counters = c->d_counters; a.dtmin = nullptr; a.mp_timestep_limit = 0;
  const long ncell = (long)c->g.NG[0] * c->g.NG[1] * c->g.NG[2];
  const size_t smem = cool_smem_bytes(c->cool);
  const int blocks = nblocks(ncell, 128, 148 * 32);
  switch (c->cfg.eqntype) {
    case PION_EQEUL: PION_COOL_DISPATCH(c->cool.mode, k_cooling_dU, EQ_EULER, blocks, 128, smem, c->stream, a) break;
    case PION_EQMHD: PION_COOL_DISPATCH(c->cool.mode, k_cooling_dU, EQ_MHD, blocks, 128, smem, c->stream, a) break;
    default: PION_COOL_DISPATCH(c->cool.mode, k_cooling_dU, EQ_GLM, blocks, 128, smem, c->stream, a) break;
  }
  c->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int pion_gpu_calc_microphysics_dU(pion_gpu_ctx* c, double dt) {
  if (!c->cfg.cooling) return 0;  // time_integrator.cpp:264: no MP -> nothing to do
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (ensure_dU(c)) return 1;
  return launch_cooling(c, dt, nullptr, c->dU);
}

extern "C" int pion_gpu_calc_dynamics_dU(pion_gpu_ctx* c, double dt, int step) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (ensure_dU(c)) return 1;
  if (!c->ph_valid) {
    CUDA_OK(cudaMemcpyAsync(c->Ph, c->P, c->arr_elems * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->ph_valid = true;
  }
  c->FV_dt = dt;
  const int order = (step == 1) ? 1 : 2;
  if (launch_preprocess(c, c->Ph, order)) return 1;
  return launch_stage(c, c->Ph, c->P, nullptr, c->dU, dt, order, false, false);
}

extern "C" int pion_gpu_grid_update_state_vector(pion_gpu_ctx* c, double dt, int step, int ooa) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (ensure_dU(c)) return 1;
  if (!c->ph_valid) {
    CUDA_OK(cudaMemcpyAsync(c->Ph, c->P, c->arr_elems * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->ph_valid = true;
  }
  c->FV_dt = dt;
  UpdateArgs a;
  a.g = c->g;
  a.pp = c->pp;
  a.P = c->P;
  a.Ph = c->Ph;
  a.dU = c->dU;
  a.mask = c->mask;
  a.counters = c->d_counters;
  a.glm_damp = exp(-dt * c->chyp * c->cr);
  a.ntr = c->ntr;
  a.full = (step == ooa) ? 1 : 0;
  const long n = (long)c->g.NGa[0] * c->g.NGa[1] * c->g.NGa[2];
  switch (c->cfg.eqntype) {
    case PION_EQEUL: k_update_state<EQ_EULER><<<nblocks(n, 256), 256, 0, c->stream>>>(a); break;
    case PION_EQMHD: k_update_state<EQ_MHD><<<nblocks(n, 256), 256, 0, c->stream>>>(a); break;
    default: k_update_state<EQ_GLM><<<nblocks(n, 256), 256, 0, c->stream>>>(a); break;
  }
  c->launches++;
  if (a.full) c->next_dt_valid = false;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// One fused stage followed by its boundary update.  With a communicator the stage is split: the tiles next
// to the six faces (ONE launch over a table of six boxes) run first on the high-priority comm stream, then
// the boundary update -- stellar-wind cells, ghost-fill kernels and the NCCL halo exchange, axis by axis --
// runs there WHILE the interior tiles run on the compute stream; the next stage waits for both.
// A stellar-wind internal boundary does not prevent the split: its cells are outside the domain (mask), hold
// the same constant reference state in P and Ph at all times, and k_wind_set only rewrites those constants.
static int stage_and_bcs_impl(pion_gpu_ctx* c, const double* S, const double* Pb, double* out, double dt, int order, bool want_dt,
                              double* bcA0, double* bcA1) {
  int cx, cy;
  stage_tile_cells(c, order, &cx, &cy);
  const GridD& g = c->g;
  const int ntx = (g.NG[0] + cx - 1) / cx, nty = (g.NG[1] + cy - 1) / cy, NZ = g.NG[2];
  // shell thickness in tiles: the last tile may hold fewer than the 2 cells the halo slab needs
  const int sxh = (g.NG[0] - (ntx - 1) * cx < 2) ? 2 : 1, syh = (g.NG[1] - (nty - 1) * cy < 2) ? 2 : 1;
  const int zs = 8;
  int nmpi = 0, nint_other = 0;
  for (int f = 0; f < 6; f++) nmpi += (c->cfg.bc[f] == PION_BC_MPI);
  for (int i = 0; i < c->cfg.n_internal_bc; i++) nint_other += (c->cfg.internal_bc[i] != PION_BC_STWIND);
  const bool want = c->force_overlap || nmpi >= 1;
  const bool overlap = want && c->comm && nint_other == 0 && g.ndim == 3 && g.coord == PION_COORD_CRT && !c->force_gather &&
                       !c->no_overlap && ntx >= 2 + sxh && nty >= 2 + syh && NZ >= 3 * zs;
  c->last_stage_split = overlap;
  if (!overlap) {
    if (launch_stage(c, S, Pb, out, nullptr, dt, order, true, want_dt)) return 1;
    return update_bcs_arrays(c, bcA0, bcA1, c->simtime);
  }
  const StageBox shell[6] = {
      {0, ntx, 0, nty, 0, zs},                          // z low
      {0, ntx, 0, nty, NZ - zs, NZ},                    // z high
      {0, ntx, 0, 1, zs, NZ - zs},                      // y low
      {0, ntx, nty - syh, nty, zs, NZ - zs},            // y high
      {0, 1, 1, nty - syh, zs, NZ - zs},                // x low
      {ntx - sxh, ntx, 1, nty - syh, zs, NZ - zs},      // x high
  };
  const StageBox interior = {1, ntx - sxh, 1, nty - syh, zs, NZ - zs};
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (c->timing) {  // the launches of a split stage count as ONE stage launch for the roofline leg
    if (timing_event(c, &e0) || timing_event(c, &e1)) return 1;
    CUDA_OK(cudaEventRecord(e0, c->stream));
    c->timing_suspended = true;
  }
  if (want_dt) CUDA_OK(cudaMemcpyAsync(c->d_dtmin, &DT_INIT_BITS, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
  // comm stream (high priority): shell tiles, then ghost fill + halo exchange; compute stream: interior
  CUDA_OK(cudaEventRecord(c->ev_a, c->stream));
  CUDA_OK(cudaStreamWaitEvent(c->comm_stream, c->ev_a, 0));
  if (launch_stage(c, S, Pb, out, nullptr, dt, order, true, want_dt, shell, 6, false, c->comm_stream)) return 1;
  if (launch_stage(c, S, Pb, out, nullptr, dt, order, true, want_dt, &interior, 1, false, c->stream)) return 1;
  if (update_bcs_arrays(c, bcA0, bcA1, c->simtime, c->comm_stream)) return 1;
  CUDA_OK(cudaEventRecord(c->ev_b, c->comm_stream));
  CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
  if (c->timing) {  // begin .. join: the whole split stage incl. the exchange it overlaps
    CUDA_OK(cudaEventRecord(e1, c->stream));
    c->tev.push_back(e0);
    c->tev.push_back(e1);
  }
  return 0;
}
static int stage_and_bcs(pion_gpu_ctx* c, const double* S, const double* Pb, double* out, double dt, int order, bool want_dt,
                         double* bcA0, double* bcA1) {
  const int err = stage_and_bcs_impl(c, S, Pb, out, dt, order, want_dt, bcA0, bcA1);
  c->timing_suspended = false;  // also on the error paths
  return err;
}

// time_integrator::advance_time (time_integrator.cpp:72-142), fused fast path.
//   predictor : stencil from P (== Ph at the start of a step), writes Ph
//   corrector : stencil from Ph, base state P, writes P in place and reduces the
//               next step's CFL dt; Ph's interior is then stale until the next predictor.
// HBM traffic per cell-update: read P, write Ph, read Ph + P, write P = 5*nvar*8 B.
extern "C" int pion_gpu_advance_time(pion_gpu_ctx* c, double* dt_done) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const double dt = c->dt;
  if (c->cfg.tmOOA == 1) {
    // first_order_update(dt, OA1) + BCs (OA1, OA1): full step in one stage
    c->FV_dt = dt;
    // the single stage reads P's stencil and must not write P in place: go through Ph
    if (c->cfg.cooling && launch_cooling(c, dt, c->mp_dE, nullptr)) return 1;
    c->mp_dE_stage = c->mp_dE;
    if (launch_preprocess(c, c->P, 1)) return 1;
    if (launch_stage(c, c->P, c->P, c->Ph, nullptr, dt, 1, true, true)) return 1;
    // P = Ph on the interior, then boundaries of both
    const long ncell = (long)c->g.NG[0] * c->g.NG[1] * c->g.NG[2];
    k_copy_interior<<<nblocks(ncell, 256), 256, 0, c->stream>>>(c->g, c->Ph, c->P, c->nvar, -1);
    c->launches++;
    c->ph_valid = true;
    if (update_bcs_arrays(c, c->Ph, c->P, c->simtime)) return 1;
    c->next_dt_valid = true;
  } else {
    // first_order_update(0.5 dt, OA2): Setdt(0.5dt), dynamics OA1, update Ph
    c->FV_dt = 0.5 * dt;
    // calc_microphysics_dU(0.5dt) of the predictor AND calc_microphysics_dU(dt) of the corrector: both integrate from
    // the start-of-step P (time_integrator.cpp:472), so one launch produces both source terms
    if (c->cfg.cooling && launch_cooling(c, 0.5 * dt, c->mp_dE, nullptr, dt, c->mp_dE2)) return 1;
    c->mp_dE_stage = c->mp_dE;
    if (launch_preprocess(c, c->P, 1)) return 1;
    // ... then boundaries of Ph (cstep=OA1 != maxstep=OA2), simtime = start of step
    if (stage_and_bcs(c, c->P, c->P, c->Ph, 0.5 * dt, 1, false, c->Ph, nullptr)) return 1;
    // second_order_update(dt, OA2): Setdt(dt), dynamics OA2 from Ph, update P
    c->FV_dt = dt;
    c->mp_dE_stage = c->mp_dE2;
    if (launch_preprocess(c, c->Ph, 2)) return 1;
    // ... then boundaries of P (and Ph == P): only P is kept current
    if (stage_and_bcs(c, c->Ph, c->P, c->P, dt, 2, true, c->P, nullptr)) return 1;
    c->ph_valid = false;
    c->next_dt_valid = true;
  }
  c->simtime += dt;
  c->last_dt = dt;
  c->timestep++;
  if (dt_done) *dt_done = dt;
  return 0;
}

extern "C" int pion_gpu_run(pion_gpu_ctx* c, int nsteps, double* dts) {
  for (int i = 0; i < nsteps; i++) {
    double dt;
    if (pion_gpu_calculate_timestep(c, &dt)) return 1;
    if (pion_gpu_advance_time(c, nullptr)) return 1;
    if (dts) dts[i] = dt;
    // Time_Int calls output_data after every step (sim_control.cpp:252); only its next_optime bookkeeping matters here
    if (c->cfg.op_criterion == 1 && pion_gpu_output_due(c, 0, nullptr)) return 1;
  }
  return 0;
}

extern "C" int pion_gpu_counters(pion_gpu_ctx* c, long long* out3) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  long long h[3];
  CUDA_OK(cudaMemcpyAsync(h, c->d_counters, 3 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  out3[0] = h[0];
  out3[1] = h[1];
  out3[2] = c->launches;
  c->mp_failures = h[2];
  return 0;
}
extern "C" int pion_gpu_mp_failures(pion_gpu_ctx* c, long long* out) {
  long long t[3];
  if (pion_gpu_counters(c, t)) return 1;
  *out = c->mp_failures;
  return 0;
}
extern "C" int pion_gpu_riemann_failures(pion_gpu_ctx* c, long long* out) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  long long h = 0;
  CUDA_OK(cudaMemcpyAsync(&h, c->d_counters + 3, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  *out = h;
  return 0;
}
extern "C" int pion_gpu_sync(pion_gpu_ctx* c) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" void* pion_gpu_stream(pion_gpu_ctx* c) { return (void*)c->stream; }

extern "C" int pion_gpu_stage_timing(pion_gpu_ctx* c, int enable, double* total_ms, long long* nlaunch) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  for (size_t i = 0; i + 1 < c->tev.size(); i += 2) {
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, c->tev[i], c->tev[i + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (nlaunch) *nlaunch = (long long)(c->tev.size() / 2);
  for (auto e : c->tev) c->tev_pool.push_back(e);
  c->tev.clear();
  c->timing = enable != 0;
  return 0;
}

// What this context actually launches: the stage-kernel variant of its most recent stage, the options in
// effect (environment switches included) and the build flags of the library -- so that a bench line or a test
// log states what ran instead of what was meant to run.
extern "C" int pion_gpu_describe(pion_gpu_ctx* c, char* buf, int n) {
  if (!buf || n <= 0) return 1;
  int nmpi = 0;
  for (int f = 0; f < 6; f++) nmpi += (c->cfg.bc[f] == PION_BC_MPI);
  snprintf(buf, (size_t)n,
           "stage_kernel=%s; tma_tensor_maps=%d; ranks=%d; exchanged_faces=%d; halo_overlap=%s; split_stage=%d; force_gather=%d; "
           "build=%s%s%s",
           c->last_stage_kernel, c->have_tmap ? 1 : 0, c->cfg.nproc > 0 ? c->cfg.nproc : 1, nmpi,
           c->no_overlap ? "off(PION_B200_NO_OVERLAP)" : c->force_overlap ? "forced(PION_B200_OVERLAP)" : "auto",
           c->last_stage_split ? 1 : 0, c->force_gather ? 1 : 0,
#ifdef PION_STRICT
           "PION_STRICT",
#else
           "default",
#endif
#ifdef PION_BUILD_TAG
           " " PION_BUILD_TAG,
#else
           "",
#endif
           "");
  return 0;
}

// ---------------------------------------------------------------------------
// multi-GPU: NCCL halo exchange + decomposition
// ---------------------------------------------------------------------------
extern "C" int pion_gpu_nccl_unique_id(char* out128) {
  ncclUniqueId id;
  NCCL_OK(ncclGetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out128, &id, 128);
  return 0;
}

extern "C" int pion_gpu_nccl_init(pion_gpu_ctx* c, const char* unique_id128) {
  CUDA_OK(cudaSetDevice(c->cfg.device));
  ncclUniqueId id;
  memcpy(&id, unique_id128, 128);
  NCCL_OK(ncclCommInitRank(&c->comm, c->cfg.nproc, id, c->cfg.rank));
  CUDA_OK(cudaMalloc(&c->d_red, 2 * sizeof(double)));
  for (int f = 0; f < 2 * c->g.ndim; f++) {
    if (c->cfg.bc[f] != PION_BC_MPI) continue;
    c->halo_elems[f] = (size_t)face_cells(c->g, f) * c->nvar;
    CUDA_OK(cudaMalloc(&c->sendbuf[f], c->halo_elems[f] * sizeof(double)));
    CUDA_OK(cudaMalloc(&c->recvbuf[f], c->halo_elems[f] * sizeof(double)));
  }
  return 0;
}

// BC_update_BCMPI for both faces of one axis (MCMD_boundaries.cpp:122-236):
// pack -> grouped ncclSend/ncclRecv -> unpack, all on the compute stream.
static int halo_exchange_axis(pion_gpu_ctx* c, int ax, double* A, cudaStream_t st) {
  if (!c->comm) { set_error("BCMPI face but no NCCL communicator (call pion_gpu_nccl_init)"); return 1; }
  const GridD& g = c->g;
  const int f0 = 2 * ax, f1 = 2 * ax + 1;
  const bool m0 = c->cfg.bc[f0] == PION_BC_MPI, m1 = c->cfg.bc[f1] == PION_BC_MPI;
  const size_t nmax = (m0 && m1) ? (c->halo_elems[f0] > c->halo_elems[f1] ? c->halo_elems[f0] : c->halo_elems[f1])
                                 : (m0 ? c->halo_elems[f0] : c->halo_elems[f1]);
  HaloArgs h;
  h.g = g; h.A = A; h.face = f0; h.nvar = c->nvar;
  // pack the slabs of both exchanged faces in one launch
  h.buf = m0 ? c->sendbuf[f0] : nullptr; h.pack = 1;
  k_halo_axis<<<nblocks((long)nmax, 256), 256, 0, st>>>(h, m1 ? c->sendbuf[f1] : nullptr);
  c->launches++;
  // Sends go out in face order (N, P) and receives are posted in the opposite order
  // (P, N): when both neighbours of this axis are the SAME rank (2 ranks, periodic) NCCL
  // matches operations per peer in order, and the peer's N-side slab must land in our P
  // ghost layers.  (The reference gets the same pairing from its even/odd send ordering,
  // assign_update_bcs_MPI.cpp:104-124.)
  NCCL_OK(ncclGroupStart());
  for (int s = 0; s < 2; s++) {
    const int f = 2 * ax + s;
    if (c->cfg.bc[f] != PION_BC_MPI) continue;
    NCCL_OK(ncclSend(c->sendbuf[f], c->halo_elems[f], ncclDouble, c->cfg.ngbprocs[f], c->comm, st));
  }
  for (int s = 1; s >= 0; s--) {
    const int f = 2 * ax + s;
    if (c->cfg.bc[f] != PION_BC_MPI) continue;
    NCCL_OK(ncclRecv(c->recvbuf[f], c->halo_elems[f], ncclDouble, c->cfg.ngbprocs[f], c->comm, st));
  }
  NCCL_OK(ncclGroupEnd());
  h.buf = m0 ? c->recvbuf[f0] : nullptr; h.pack = 0;
  k_halo_axis<<<nblocks((long)nmax, 256), 256, 0, st>>>(h, m1 ? c->recvbuf[f1] : nullptr);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// MCMDcontrol::decomposeDomain (MCMD_control.cpp:62-221): halve the longest local
// axis until nproc sub-blocks (ties -> lowest axis); rank = nx*ny*iz + nx*iy + ix;
// pointToNeighbours (:316-420): neighbour ranks, periodic wrap.
extern "C" int pion_gpu_decompose_domain(pion_gpu_config* cfg, int rank, int nproc) {
  if (nproc < 1 || (nproc & (nproc - 1))) { set_error("nproc must be a power of two (MCMD_control.cpp:98-104)"); return 1; }
  int nx[3] = {1, 1, 1};
  double range[3], locrange[3];
  int locNG[3];
  for (int a = 0; a < 3; a++) {
    range[a] = (a < cfg->ndim) ? cfg->xmax[a] - cfg->xmin[a] : 0.0;
    locrange[a] = range[a];
    locNG[a] = (a < cfg->ndim) ? cfg->NG[a] : 1;
  }
  int npcounter = 1;
  while (npcounter < nproc) {
    int dsplit = 0;
    double maxrange = 0.;
    for (int a = 0; a < cfg->ndim; a++) {
      if (locrange[a] > maxrange * (1.0 + 1.0e-12)) { maxrange = locrange[a]; dsplit = a; }  // ties keep the lowest axis
    }
    locrange[dsplit] /= 2.;
    if (locNG[dsplit] % 2) { set_error("grid not divisible by the decomposition"); return 1; }
    locNG[dsplit] /= 2;
    nx[dsplit] *= 2;
    npcounter *= 2;
  }
  int ix[3];
  ix[0] = rank % nx[0];
  ix[1] = (rank / nx[0]) % nx[1];
  ix[2] = rank / (nx[0] * nx[1]);
  cfg->rank = rank;
  cfg->nproc = nproc;
  double gxmin[3], dxg = range[0] / cfg->NG[0];
  (void)dxg;
  for (int a = 0; a < 3; a++) gxmin[a] = cfg->xmin[a];
  for (int a = 0; a < cfg->ndim; a++) {
    cfg->sim_xmin[a] = gxmin[a];
    cfg->NG[a] = locNG[a];
    cfg->xmin[a] = gxmin[a] + ix[a] * locrange[a];
    cfg->xmax[a] = gxmin[a] + (ix[a] + 1) * locrange[a];
    const int stride = (a == 0) ? 1 : (a == 1) ? nx[0] : nx[0] * nx[1];
    const int lo = 2 * a, hi = 2 * a + 1;
    const int bc_lo = cfg->bc[lo], bc_hi = cfg->bc[hi];
    cfg->ngbprocs[lo] = cfg->ngbprocs[hi] = -1;
    if (ix[a] > 0) { cfg->ngbprocs[lo] = rank - stride; cfg->bc[lo] = PION_BC_MPI; }
    else if (bc_lo == PION_BC_PERIODIC && nx[a] > 1) { cfg->ngbprocs[lo] = rank + (nx[a] - 1) * stride; cfg->bc[lo] = PION_BC_MPI; }
    if (ix[a] < nx[a] - 1) { cfg->ngbprocs[hi] = rank + stride; cfg->bc[hi] = PION_BC_MPI; }
    else if (bc_hi == PION_BC_PERIODIC && nx[a] > 1) { cfg->ngbprocs[hi] = rank - (nx[a] - 1) * stride; cfg->bc[hi] = PION_BC_MPI; }
  }
  return 0;
}
