/* pion_b200.h -- C ABI of libpion_b200.so, the B200 (sm_100a) implementation of
 * PION's finite-volume hydro/MHD dynamics update.
 *
 * PION has no FFI: its extension seam is C++ virtual inheritance (SURVEY.md
 * 8b).  A GPU cannot sit behind the per-cell virtuals of FV_solver_base, so the
 * drop-in boundary is the GRID-LEVEL seam -- the methods of time_integrator /
 * calc_timestep / assign_update_bcs that sim_control calls once per step and
 * that sim_control_pllel / sim_control_NG already override in the reference.
 * Each entry point below names the reference method it replaces (paths relative
 * to /root/reference/source).  INTEGRATION.md shows the `sim_control_gpu`
 * subclass a PION maintainer adds to bind them.
 *
 * Conventions kept from the reference: every call returns an int error count,
 * 0 = success (the caller turns non-zero into rep.error(...), tools/reporting.h
 * :63-85); FP64 throughout (#define pion_flt double, defines/
 * functionality_flags.h:31); primitive order {RO,PG,VX,VY,VZ,BX,BY,BZ,SI,
 * tracers...} and conserved order {RHO,ERG,MMX,MMY,MMZ,BBX,BBY,BBZ,PSI,...}
 * (constants.h:256-281); B in code units (NEW_B_NORM).  One host thread per
 * context; the library owns all device memory behind the opaque handle.
 *
 * Host <-> device state exchange is structure-of-arrays: double
 * [nvar][NZ+2g][NY+2g][NX+2g], x fastest, ghost depth g = 2 (second order) or 1,
 * unused dimensions have extent 1.  That is exactly the reference's cell-id
 * order (grid/uniform_grid.cpp:449-451) transposed to variable-major.
 */
#ifndef PION_B200_H
#define PION_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PION_GPU_MAXVAR 16
#define PION_GPU_MAXTR 4

/* integer codes are the reference's (constants.h:166-246, boundaries/boundaries.h:32-52) */
enum { PION_EQEUL = 1, PION_EQMHD = 2, PION_EQGLM = 3 };
enum { PION_COORD_CRT = 1, PION_COORD_CYL = 2, PION_COORD_SPH = 3 };
enum { PION_FLUX_LF = 0, PION_FLUX_RSLINEAR = 1, PION_FLUX_RSEXACT = 2, PION_FLUX_RSHYBRID = 3, PION_FLUX_ROE = 4, PION_FLUX_ROE_PV = 5, PION_FLUX_FVS = 6, PION_FLUX_HLLD = 7, PION_FLUX_HLL = 8 };  /* 2, 3, 5, 6: Euler only (1 with MHD = the linear MHD solver of riemannMHD.cpp); 0 forces first order */
enum { PION_AV_NONE = 0, PION_AV_FKJ98 = 1, PION_AV_HCORR = 3, PION_AV_HCORR_FKJ98 = 4 };
enum {
  PION_BC_PERIODIC = 1, PION_BC_OUTFLOW = 2, PION_BC_INFLOW = 3, PION_BC_REFLECTING = 4, PION_BC_FIXED = 5,
  PION_BC_DMACH = 8, PION_BC_DMACH2 = 9, PION_BC_MPI = 10, PION_BC_ONEWAY_OUT = 13, PION_BC_STWIND = 14
};
enum { PION_STATE_P = 0, PION_STATE_PH = 1, PION_STATE_DU = 2 };

/* One constant stellar-wind source: the arguments of stellar_wind::add_source
 * (grid/stellar_wind_BC.cpp:125-140) in parameter-file units (SWP.params[i],
 * sim_params.h:340-380; WINDTYPE_CONSTANT only). */
typedef struct pion_gpu_wind_source {
  double dpos[3];    /* position, cm */
  double radius;     /* boundary radius, cm */
  double mdot;       /* Msun/yr */
  double vinf, vrot; /* km/s */
  double temp;       /* wind temperature, K */
  double rstar;      /* stellar radius, cm */
  double bsrf;       /* surface split-monopole field, Gauss */
  double tr[PION_GPU_MAXTR]; /* tracer values */
} pion_gpu_wind_source;

/* Mirrors the subset of class SimParams (sim_params.h:200-285) the path reads. */
typedef struct pion_gpu_config {
  int device;          /* CUDA device ordinal */
  int ndim;            /* SimPM.ndim */
  int NG[3];           /* LOCAL interior cells per axis (SimPM.NG / MCMD LocalNG) */
  int nvar, ntracer;   /* SimPM.nvar, SimPM.ntracer (tracers are the last ntracer variables) */
  int eqntype;         /* SimPM.eqntype */
  int coord_sys;       /* SimPM.coord_sys: Cartesian 1-3D, cylindrical (z,R) 2-D, spherical 1-D Euler */
  int solver;          /* SimPM.solverType: 0 Lax-Friedrichs, 1 / 2 / 3 linear / exact / hybrid Riemann solver (Euler; MHD: 1 only), 4 Roe-CV, 5 Roe-PV, 6 FVS, 7 HLLD, 8 HLL */
  int artviscosity;    /* SimPM.artviscosity: 0,1,3,4 */
  int spOOA, tmOOA;    /* SimPM.spOOA / tmOOA: (1,1) or (2,2) */
  double gamma, cfl, etav;
  double xmin[3], xmax[3];   /* LOCAL domain; dx = (xmax[0]-xmin[0])/NG[0] */
  double sim_xmin[3];        /* GLOBAL SimPM.Xmin (cell positions for DMR boundaries) */
  int bc[6];           /* per LOCAL face XN,XP,YN,YP,ZN,ZP; PION_BC_MPI for a face shared with a neighbour rank */
  int n_internal_bc;
  int internal_bc[4];
  double refvec[PION_GPU_MAXVAR]; /* SimPM.RefVec */
  double starttime, finishtime;
  int op_criterion;    /* SimPM.op_criterion, 1: dt limited by next_optime */
  double opfreq_time;
  /* microphysics: mp_only_cooling (EP.cooling && !EP.chemistry) */
  int cooling;         /* EP.cooling (mp_only_cooling.cpp:42-48): 0 none, 2 KI02, 4 SD93_CIE, 5 SD93_PLUS_HEATING,
                          6 WSS09_CIE_PLUS_HEATING, 7 WSS09_CIE_ONLY_COOLING, 8 WSS09_CIE_LINE_HEAT_COOL */
  int mp_timestep_limit;
  double min_temperature, max_temperature;
  int n_table;
  const double *table_T, *table_rrhp, *table_C_rrh, *table_C_ffhe, *table_C_fbdn, *table_C_cie;
  /* decomposition (decomposition/MCMD_control.cpp:62-221) */
  int rank, nproc;
  int ngbprocs[6];     /* neighbour rank per face, -1 = none (MCMDcontrol::ngbprocs) */
  /* internal boundary PION_BC_STWIND ("stellar-wind"): BC_assign_STWIND / BC_update_STWIND
   * (boundaries/stellar_wind_boundaries.cpp:29-341) for constant sources */
  int n_wind;
  pion_gpu_wind_source wind[2];
  double min_timestep; /* SimPM.min_timestep (sim_params.h:227): calculate_timestep fails if dt falls below it */
  /* EP.cooling 4..7: the knots (log10 T, log10 Lambda) of cooling_function_SD93CIE's cooling-curve spline --
   * Tarray / Larray after setup_SD93_cie() [4, 5] or setup_WSS09_CIE() [6, 7] -- and its power-law slopes outside
   * the table (microphysics/cooling_SD93_cie.cpp:87-200,555-704); n_table / table_* are for EP.cooling 8 only */
  int n_spline;
  const double *spline_logT, *spline_logL;
  double spline_min_slope, spline_max_slope;
} pion_gpu_config;

typedef struct pion_gpu_ctx pion_gpu_ctx;

/* setup_fixed_grid::setup_grid + set_equations + setup_microphysics
 * (grid/setup_fixed_grid.cpp:160-246,254-470,1067-1190): allocates the device
 * SoA grid.  Returns NULL on failure (reason via pion_gpu_last_error). */
pion_gpu_ctx *pion_gpu_create(const pion_gpu_config *cfg);
void pion_gpu_destroy(pion_gpu_ctx *ctx);
const char *pion_gpu_last_error(void);

/* host -> device / device -> host copies of P, Ph or dU (dataio->ReadData /
 * OutputData side of the seam, sim_init.cpp:213, :671-760) */
/* Both directions move one variable at a time as a flat PCIe copy through a device staging buffer and re-pitch it on the
 * device (55 GB/s measured on B200 with page-locked host memory: cudaHostAlloc / cudaHostRegister the buffer; pageable
 * memory works but is copied synchronously by the driver). */
int pion_gpu_upload(pion_gpu_ctx *ctx, int which, const double *soa);
int pion_gpu_download(pion_gpu_ctx *ctx, int which, double *soa);

/* sim_init::Init after ReadData (sim_init.cpp:215-262): Ph=P, psi=0 for GLM at
 * step 0, assign_boundary_data (boundaries/assign_update_bcs.cpp:28-120) and
 * the first TimeUpdateInternal/ExternalBCs. */
int pion_gpu_init_after_upload(pion_gpu_ctx *ctx);

/* calc_timestep::calc_dynamics_dt / calc_microphysics_dt
 * (sim_control/calc_timestep.cpp:271-333, :342-463).  LOCAL minima. */
int pion_gpu_calc_dt(pion_gpu_ctx *ctx, double *t_dyn, double *t_mp);
/* calc_timestep::calculate_timestep (calc_timestep.cpp:68-153) incl.
 * Set_GLM_Speeds and timestep_checking_and_limiting (:219-262); with nproc>1
 * the minima are reduced over ranks (sim_control_MPI.cpp:503-504). */
int pion_gpu_calculate_timestep(pion_gpu_ctx *ctx, double *dt);
/* FV_solver_base::Setdt + Set_GLM_Speeds(spatial_solvers/solver_eqn_mhd_adi.cpp:906) */
int pion_gpu_set_dt(pion_gpu_ctx *ctx, double dt);
int pion_gpu_set_glm_speeds(pion_gpu_ctx *ctx, double t_dyn, double dx, double cr);
int pion_gpu_set_time(pion_gpu_ctx *ctx, double simtime, double last_dt, int timestep);
int pion_gpu_get_time(pion_gpu_ctx *ctx, double *simtime, double *dt, double *last_dt, int *timestep);

/* time_integrator::calc_microphysics_dU (time_integrator.cpp:253, :438): per-cell
 * mp_only_cooling::TimeUpdateMP (adaptive RK5 Cash-Karp) from P, dU[ERG] += ... */
int pion_gpu_calc_microphysics_dU(pion_gpu_ctx *ctx, double dt);
/* time_integrator::calc_dynamics_dU (time_integrator.cpp:498): preprocess_data +
 * set_dynamics_dU; accumulates into the device dU array. `step` is OA1 / OA2. */
int pion_gpu_calc_dynamics_dU(pion_gpu_ctx *ctx, double dt, int step);
/* time_integrator::grid_update_state_vector (time_integrator.cpp:881) */
int pion_gpu_grid_update_state_vector(pion_gpu_ctx *ctx, double dt, int step, int ooa);
/* assign_update_bcs::TimeUpdateInternalBCs + TimeUpdateExternalBCs
 * (boundaries/assign_update_bcs.cpp:134-246); with nproc>1 the BCMPI faces are
 * NCCL halo exchanges (boundaries/MCMD_boundaries.cpp:57-236). */
int pion_gpu_time_update_bcs(pion_gpu_ctx *ctx, double simtime, int cstep, int maxstep);
/* ... and the two halves on their own: TimeUpdateInternalBCs (:134-181; of the internal boundaries only
 * STWIND is updated there) and TimeUpdateExternalBCs (:191-246; the faces in BC_bd order, then DMACH2). */
int pion_gpu_time_update_internal_bcs(pion_gpu_ctx *ctx, double simtime, int cstep, int maxstep);
int pion_gpu_time_update_external_bcs(pion_gpu_ctx *ctx, double simtime, int cstep, int maxstep);
/* time_integrator::advance_time (time_integrator.cpp:72-142): the fused fast
 * path (predictor, BCs, corrector, BCs, next-step CFL reduction); returns dt. */
int pion_gpu_advance_time(pion_gpu_ctx *ctx, double *dt_done);
/* nsteps x { calculate_timestep; advance_time } = body of sim_control::Time_Int
 * (sim_control.cpp:220-266); dts[nsteps] (optional) receives each dt. */
int pion_gpu_run(pion_gpu_ctx *ctx, int nsteps, double *dts);

/* The output-criterion part of sim_init::output_data (sim_init.cpp:711-744): *due = 1 if the current step is
 * one the caller should save (op_criterion 0: every `opfreq` steps; 1: simtime has reached next_optime, which
 * is then advanced by opfreq_time exactly where the reference does it -- calculate_timestep limits dt by it).
 * sim_control::Time_Int calls output_data after every step (sim_control.cpp:252); pion_gpu_run does this
 * bookkeeping itself. */
int pion_gpu_output_due(pion_gpu_ctx *ctx, int opfreq, int *due);

/* error / diagnostic counters accumulated on the device:
 * [0] negative-density events (fatal in the reference), [1] negative-pressure
 * fix-ups, [2] kernels launched so far */
int pion_gpu_counters(pion_gpu_ctx *ctx, long long *out3);
/* cells whose cooling integration failed so far (mp_only_cooling.cpp:203-207: fatal in the
 * reference; the caller should turn a non-zero count into rep.error) */
int pion_gpu_mp_failures(pion_gpu_ctx *ctx, long long *out);
/* interfaces at which the linear / exact / hybrid Riemann solver returned an error so far (riemann_Euler::
 * JMs_riemann_solve, riemann.cpp:245-463: the error propagates out of InterCellFlux and is fatal in the reference) */
int pion_gpu_riemann_failures(pion_gpu_ctx *ctx, long long *out);
/* block until all queued device work of this context has finished */
int pion_gpu_sync(pion_gpu_ctx *ctx);
/* stream the context launches on (cudaStream_t as void*), for CUDA-event timing */
void *pion_gpu_stream(pion_gpu_ctx *ctx);

/* bench support: with enable!=0 every later stage-kernel launch is bracketed by CUDA
 * events on the context's stream; the call returns the summed duration and count of
 * the launches recorded since the previous call and resets the record. */
int pion_gpu_stage_timing(pion_gpu_ctx *ctx, int enable, double *total_ms, long long *nlaunch);

/* one line saying what this context launches: the stage-kernel variant of its most recent stage
 * (template arguments included), tensor maps, halo-overlap mode, environment switches in effect, build flags */
int pion_gpu_describe(pion_gpu_ctx *ctx, char *buf, int n);

/* multi-GPU: NCCL communicator for the BCMPI faces and the dt all-reduce
 * (replaces comms/comm_mpi.cpp).  `unique_id` is the 128-byte ncclUniqueId
 * produced by pion_gpu_nccl_unique_id on rank 0 and broadcast by the host. */
int pion_gpu_nccl_unique_id(char *out128);
int pion_gpu_nccl_init(pion_gpu_ctx *ctx, const char *unique_id128);

/* decomposition helper = MCMDcontrol::decomposeDomain + pointToNeighbours
 * (decomposition/MCMD_control.cpp:62-221, :316-420): fills local NG, xmin, xmax,
 * bc[] (PION_BC_MPI on shared faces) and ngbprocs[] of `cfg` for `rank` of
 * `nproc` from the GLOBAL values already in cfg. */
int pion_gpu_decompose_domain(pion_gpu_config *cfg, int rank, int nproc);

#ifdef __cplusplus
}
#endif
#endif
