// pion_b200/csrc/cooling.cuh -- per-cell radiative cooling source term
// (microphysics without chemistry: mp_only_cooling, every cooling function its Edot dispatches to).
//
// Reference path restated (paths relative to /root/reference/source):
//   sim_control/time_integrator.cpp:438-489   calc_noRT_microphysics_dU: for every isdomain
//                                             cell  dU += PtoU(TimeUpdateMP(P, dt)) - PtoU(P)
//   microphysics/mp_only_cooling.cpp:167-221  TimeUpdateMP (integrates from the UNCLAMPED
//                                             Eint0; T clamp applied to the result)
//   microphysics/mp_only_cooling.cpp:383-420  Edot: dispatch on EP_cooling (2, 4, 5, 6, 7, 8)
//   microphysics/mp_only_cooling.cpp:427-468  Edot_SD93CIE_cool / _heat_cool, Edot_WSS09CIE_cool / _heat_cool
//   microphysics/cooling_SD93_cie.cpp:666-704 cooling_rate_SD93CIE: natural cubic spline in (log10 T, log10 Lambda),
//                                             power laws outside the table
//   microphysics/cooling.cpp:325-399          CoolingFn::CoolingRate, WhichFunction 2 (KI02)
//   microphysics/mp_only_cooling.cpp:470-521  Edot_WSS09CIE_heat_cool_metallines: binary
//                                             search in the 200-point T table + linear interp.
//   microphysics/mp_only_cooling.cpp:333-358  timescales (cooling time)
//   microphysics/integrator.cpp:285-371       Step_RK5CK (first-order shortcut if |k1|dt/E<1e-6)
//   microphysics/integrator.cpp:401-530       Stepper_RKCK (bisection on the error estimate)
//   microphysics/integrator.cpp:540-606       Int_Adaptive_RKCK (<= 25 sub-steps)
//   sim_control/calc_timestep.cpp:342-463     calc_microphysics_dt
//
// One thread per cell; the 11 table columns (T, 5 rates, 5 slopes) sit in shared memory
// (17.6 KB) so the data-dependent binary search never leaves the SM.  The integrator keeps
// the reference's exact control flow (trip counts vary per cell: divergence is inherent);
// arithmetic here is plain IEEE FP64 (divisions included) -- this kernel moves 24 B per cell
// and is nowhere near any roofline that matters for the step.
#pragma once
#include "stage_kernel.cuh"

namespace pion {

struct CoolParams {
  // device tables.  EP_cooling 8: [11][nT] = T, rrhp, C_rrh, C_ffhe, C_fbdn, C_cie, then the 5 slopes;
  // EP_cooling 4..7: [3][nT] = spline knots x = log10 T, y = log10 Lambda, c = y''/2 (natural spline); 2: none
  const double* tables;
  int nT;
  int mode;  // EP_cooling
  double inv_Mu2, inv_Mu2_elec_H, Mu_tot_over_kB, MinT, MaxT;
  double Mu, Mu_elec, Mu_ion, smin, smax;  // mean masses; spline slopes below / above the table
  // EP_cooling 8: the T column is log-uniform (mp_only_cooling.cpp:117-160 builds it that way; checked on the host):
  // the interval of a temperature is GUESSED from a float log2 and then corrected against the table, instead of the
  // eight dependent shared-memory loads of the reference's binary search -- same interval for every input
  int guess_ok;
  float l2T0, inv_l2step;
};

// columns of the device table per cooling function
__host__ __device__ inline int cool_ncol(int mode) { return mode == 8 ? 11 : (mode >= 4 && mode <= 7) ? 3 : 0; }
// dynamic shared memory of the cooling kernels: the table columns
inline size_t cool_smem_bytes(const CoolParams& cp) { return (size_t)cool_ncol(cp.mode) * cp.nT * sizeof(double); }

struct CoolArgs {
  GridD g;
  CoolParams cp;
  const double* P;            // state the source term is integrated from (start-of-step P)
  double* dE;                 // fused path: energy source per cell (one plane), overwritten
  double* dE2;                // fused second-order step: the source of the SECOND interval dt2 (the corrector's), or null
  double dt2;
  double* dU;                 // seam path: full dU array, energy plane accumulated (+=)
  const unsigned char* mask;  // isdomain
  double dt, gamma;
  long long* counters;        // [2] integration failures (fatal in the reference)
  unsigned long long* dtmin;  // k_mp_dt: ordered-bits min of the cooling time
  int mp_timestep_limit;
};

// The table columns in shared memory: ONE base pointer, column q starts at tab + q nT (eleven column pointers held
// across the integrator cost 22 registers: with them the kernel did not fit the 64 registers that 32 warps per SM allow)
struct CoolTab {
  const double* tab;
  int nT;
  __device__ __forceinline__ const double* col(int q) const { return tab + q * nT; }
  double inv_Mu2, inv_Mu2_elec_H;
  double Mu, Mu_elec, Mu_ion, smin, smax;
  int guess_ok;
  float l2T0, inv_l2step;
};

__device__ __forceinline__ CoolTab cool_tables_to_smem(const CoolParams& cp, double* s) {
  for (int t = threadIdx.x; t < cool_ncol(cp.mode) * cp.nT; t += blockDim.x) s[t] = cp.tables[t];
  __syncthreads();
  CoolTab ct;
  ct.tab = s;  // EP_cooling 8 columns: 0 T, 1 rrhp, 2 C_rrh, 3 C_ffhe, 4 C_fbdn, 5 C_cie, 6..10 their slopes
  ct.nT = cp.nT;
  ct.inv_Mu2 = cp.inv_Mu2;
  ct.inv_Mu2_elec_H = cp.inv_Mu2_elec_H;
  ct.Mu = cp.Mu; ct.Mu_elec = cp.Mu_elec; ct.Mu_ion = cp.Mu_ion; ct.smin = cp.smin; ct.smax = cp.smax;
  ct.guess_ok = cp.guess_ok; ct.l2T0 = cp.l2T0; ct.inv_l2step = cp.inv_l2step;
  return ct;
}

// cooling_function_SD93CIE::cooling_rate_SD93CIE (cooling_SD93_cie.cpp:666-704): Lambda(T) from the natural cubic
// spline through (log10 T, log10 Lambda) -- GSL cspline evaluation: bisection for the interval, then the cubic in
// (x - x_i) with b_i, d_i formed from the second-derivative coefficients (tools/interpolate.cpp:59-118) -- and the
// power laws beyond the ends of the table.  Columns (CoolTab): T = x, rrhp = y, Crrh = c.
__device__ __forceinline__ double cool_rate_SD93CIE(const CoolTab& t, double T) {
  if (T < 0.0 || !isfinite(T)) return HUGE_VAL;
  const double* x = t.col(0);
  const double* y = t.col(1);
  const double* c = t.col(2);
  const int n = t.nT;
  double rate;
  T = log10(T);
  const double MinTemp = x[0], MaxTemp = x[n - 1];
  if (T > MaxTemp) rate = y[n - 1] + t.smax * (T - MaxTemp);
  else if (T < MinTemp) rate = y[0] + t.smin * (T - MinTemp);
  else {
    int lo = 0, hi = n - 1;
    while (hi > lo + 1) {
      const int mid = (lo + hi) >> 1;
      if (x[mid] > T) hi = mid;
      else lo = mid;
    }
    const double dx = x[lo + 1] - x[lo], dy = y[lo + 1] - y[lo];
    const double c_i = c[lo], c_ip1 = c[lo + 1];
    const double b_i = dy / dx - dx * (c_ip1 + 2.0 * c_i) / 3.0;
    const double d_i = (c_ip1 - c_i) / (3.0 * dx);
    const double delx = T - x[lo];
    rate = y[lo] + delx * (b_i + delx * (c_i + delx * d_i));
  }
  return exp(2.3025850929940459 * rate);  // pconst.ln10() (constants.h:44)
}

__device__ __forceinline__ double cool_Edot_metallines(const CoolTab& t, double rho, double T);

// mp_only_cooling::Edot (mp_only_cooling.cpp:383-420): the dispatch on EP_cooling is a TEMPLATE parameter -- the
// integrator inlines Edot seven times, and with every cooling function's code (exp / log10 bodies) in each copy the
// kernel needed 126 registers whichever function ran
template <int MODE>
__device__ __forceinline__ double cool_Edot(const CoolTab& t, double rho, double T) {
  switch (MODE) {
    case 2: {  // KI02: -CoolingFn::CoolingRate(T, 0, rho/Mu, 0, 0), WhichFunction 2, MinTemp 5 K (cooling.cpp:325-399)
      const double nH = rho / t.Mu;
      if (T <= 0.0 || isnan(T) || isinf(T)) return -0.0;
      double rate = 0.0;
      if (T > 5.0) rate += nH * nH * (2.0e-19 * exp(-1.184e5 / (T + 1.0e3)) + 2.8e-28 * sqrt(T) * exp(-92.0 / T));
      rate -= nH * 2.0e-26;
      return -rate;
    }
    case 4:  // Edot_SD93CIE_cool (:427-433)
      return -(rho * rho / t.Mu_elec / t.Mu_ion) * cool_rate_SD93CIE(t, T);
    case 5:  // Edot_SD93CIE_heat_cool (:444-451)
      return (rho * rho) * (2.733e-21 * exp(-0.782991 * log(T)) / t.Mu_elec / t.Mu - cool_rate_SD93CIE(t, T) / t.Mu_elec / t.Mu_ion);
    case 7:  // Edot_WSS09CIE_cool (:461-467)
      return 2e-26 * rho / t.Mu - (rho * rho / t.Mu / t.Mu) * cool_rate_SD93CIE(t, T);
    case 6:  // Edot_WSS09CIE_heat_cool
      return (rho * rho) * (2.733e-21 * exp(-0.782991 * log(T)) / t.Mu_elec / t.Mu - cool_rate_SD93CIE(t, T) / t.Mu / t.Mu);
    default:
      return cool_Edot_metallines(t, rho, T);
  }
}

// mp_only_cooling::Edot_WSS09CIE_heat_cool_metallines (mp_only_cooling.cpp:470-521)
__device__ __forceinline__ double cool_Edot_metallines(const CoolTab& t, double rho, double T) {
  // The reference's bisection ends at iT = the last entry below T, clamped to [0, nT-2] (0 for T <= T[0], NaN or
  // negative T; nT-2 for T > T[nT-1]).  Same index here: guess, then walk until both neighbours agree.
  int iT;
  if (t.guess_ok) {
    int gss = (int)((__log2f((float)T) - t.l2T0) * t.inv_l2step);  // (int) of NaN is 0, of +-inf saturates
    gss = max(0, min(gss, t.nT - 2));
    const double* tT = t.tab;
    while (gss > 0 && !(tT[gss] < T)) gss--;
    while (gss < t.nT - 2 && tT[gss + 1] < T) gss++;
    iT = gss;
  } else {
    int ihi = t.nT - 1, ilo = 0;
    do {
      const int imid = ilo + ((ihi - ilo) >> 1);  // ilo + floor((ihi-ilo)/2.0)
      if (t.tab[imid] < T) ilo = imid;
      else ihi = imid;
    } while (ihi - ilo > 1);
    iT = ilo;
  }
  const double* e = t.tab + iT;  // entry iT of column q: e[q nT]
  const int n = t.nT;
  const double dT = T - e[0];
  const double rho2 = rho * rho;
  double rate = -(e[4 * n] + dT * e[9 * n]) * rho2 * t.inv_Mu2_elec_H;   // C_fbdn
  rate = fmin(rate, -(e[5 * n] + dT * e[10 * n]) * rho2 * t.inv_Mu2);    // C_cie
  rate -= (e[2 * n] + dT * e[7 * n]) * rho2 * t.inv_Mu2_elec_H;          // C_rrh
  rate -= (e[3 * n] + dT * e[8 * n]) * rho2 * t.inv_Mu2_elec_H;          // C_ffhe
  rate += 8.01e-12 * (e[n] + dT * e[6 * n]) * rho2 * t.inv_Mu2_elec_H;   // rrhp
  return rate;
}

struct CoolCell {
  double rho, gm1, Mu_tot_over_kB;
};
// mp_only_cooling::dPdt (:227-236)
template <int MODE>
__device__ __forceinline__ double cool_dPdt(const CoolTab& t, const CoolCell& c, double E) {
  return cool_Edot<MODE>(t, c.rho, E * c.gm1 * c.Mu_tot_over_kB / c.rho);
}

// Integrator_Base::Step_RK5CK for one variable (integrator.cpp:285-371).  k1 = dPdt(p0) is handed in: it only depends
// on p0, which the bisection loop of the stepper does not change, so it is evaluated once per stepper call instead of once
// per trial step -- and once per CELL for the first sub-step of the two intervals (dt/2, dt) a second-order step
// integrates from the same P (k_cooling_dU2).  Same value, same rounding: bit-identical results.
template <int MODE>
__device__ __forceinline__ void cool_step_rk5ck(const CoolTab& t, const CoolCell& c, double p0, double k1, double dt, double& pf, double& dp) {
  const double b21 = 0.2, b31 = 3. / 40., b32 = 9. / 40., b41 = 0.3, b42 = -0.9, b43 = 1.2, b51 = -11. / 54., b52 = 2.5,
               b53 = -70. / 27., b54 = 35. / 27., b61 = 1631. / 55296., b62 = 175. / 512., b63 = 575. / 13824.,
               b64 = 44275. / 110592., b65 = 253. / 4096., c1 = 37. / 378., c3 = 250. / 621., c4 = 125. / 594.,
               c6 = 512. / 1771.;
  const double dc1 = c1 - 2825. / 27648., dc3 = c3 - 18575. / 48384., dc4 = c4 - 13525. / 55296., dc5 = -277. / 14336.,
               dc6 = c6 - 0.25;
  double ptemp = 0.0;
  ptemp += fabs(k1) * dt / (p0 + 1.0e-100);
  if (ptemp < 1.e-6) {
    pf = p0 + k1 * dt;
    dp = k1 * dt;
    return;
  }
  k1 *= dt;
  ptemp = p0 + b21 * k1;
  double k2 = cool_dPdt<MODE>(t, c, ptemp) * dt;
  ptemp = p0 + b31 * k1 + b32 * k2;
  double k3 = cool_dPdt<MODE>(t, c, ptemp) * dt;
  ptemp = p0 + b41 * k1 + b42 * k2 + b43 * k3;
  double k4 = cool_dPdt<MODE>(t, c, ptemp) * dt;
  ptemp = p0 + b51 * k1 + b52 * k2 + b53 * k3 + b54 * k4;
  double k5 = cool_dPdt<MODE>(t, c, ptemp) * dt;
  ptemp = p0 + b61 * k1 + b62 * k2 + b63 * k3 + b64 * k4 + b65 * k5;
  double k6 = cool_dPdt<MODE>(t, c, ptemp) * dt;
  pf = p0 + c1 * k1 + c3 * k3 + c4 * k4 + c6 * k6;
  dp = dc1 * k1 + dc3 * k3 + dc4 * k4 + dc5 * k5 + dc6 * k6;
}

// Integrator_Base::Stepper_RKCK, BISECTION_STEPPER variant (integrator.cpp:401-530)
template <int MODE>
__device__ __forceinline__ int cool_stepper(const CoolTab& t, const CoolCell& c, double p0, double k1, double t0, double htry,
                                            double errtol, double& p1, double& hdid, double& hnext) {
  int rval = 0, ct = 0;
  double h = htry, maxerr, err = 0.0, ptemp = 0.0;
  if (h < 0) return 1;
  do {
    cool_step_rk5ck<MODE>(t, c, p0, k1, h, ptemp, err);
    maxerr = 0;
    if (!isfinite(err) || !isfinite(ptemp) || ptemp < 0.0) {
      maxerr = fmax(maxerr, 1000.0);
    } else {
      err /= fabs(ptemp) + 1.e-100;
      err = fabs(err / errtol);
      maxerr = fmax(maxerr, err);
    }
    if (maxerr > 1.) h /= 2.0;
    if (t0 + h == t0) return -2;
    ct++;
  } while (maxerr > 1.0 && ct < 50);
  if (maxerr > 1.0) rval += ct + (int)(fabs(maxerr));
  hnext = h * 2.0;
  hdid = h;
  p1 = ptemp;
  if (isnan(p1) || isinf(p1)) { p1 = -1.e100; rval++; }
  return rval;
}

// Integrator_Base::Int_Adaptive_RKCK (integrator.cpp:540-606); k1_0 = dPdt(p0) of the first sub-step
template <int MODE>
__device__ __forceinline__ int cool_integrate(const CoolTab& t, const CoolCell& c, double p0, double k1_0, double dt, double& pf) {
  double tt = 0.0, p1 = p0, p2 = 0.0, k1 = k1_0;
  const double tf = 0.0 + dt;
  double h = dt, hdid = 0.0, hnext = 0.0;
  int err = 0, ct = 0;
  for (;;) {
    err += cool_stepper<MODE>(t, c, p1, k1, tt, h, 1.0e-2, p2, hdid, hnext);
    tt += hdid;
    h = fmin(hnext, tf - tt);
    ct++;
    p1 = p2;
    if (!(tt < tf && (err == 0) && (ct < 25))) break;
    k1 = cool_dPdt<MODE>(t, c, p1);
  }
  pf = p1;
  return err;
}
template <int MODE>
__device__ __forceinline__ int cool_integrate(const CoolTab& t, const CoolCell& c, double p0, double dt, double& pf) {
  return cool_integrate<MODE>(t, c, p0, cool_dPdt<MODE>(t, c, p0), dt, pf);
}

// dE = (PtoU(p') - PtoU(P)).erg for the integrated internal energy Eint (calc_noRT_microphysics_dU, time_integrator.cpp:
// 438-489, with the temperature clamp of TimeUpdateMP, mp_only_cooling.cpp:208-216)
template <int EQ>
__device__ __forceinline__ double cool_dE_of(const CoolArgs& a, const Prim& p, const Cons& ui, double Eint) {
  Prim po = p;
  po.pg = Eint * (a.gamma - 1);
  const double Tf = po.pg * a.cp.Mu_tot_over_kB / po.ro;
  if (Tf > a.cp.MaxT) po.pg *= a.cp.MaxT / Tf;
  else if (Tf < a.cp.MinT) po.pg *= a.cp.MinT / Tf;
  Cons uf;
  PtoU<EQ>(po, uf, a.gamma - 1.0);
  return uf.erg - ui.erg;
}

// calc_noRT_microphysics_dU: one thread per interior cell.  A second-order step calls it twice from the same P -- with dt/2
// before the predictor and with dt before the corrector (time_integrator.cpp:150-250) -- so the fused path integrates both
// intervals in ONE launch (dE for dt, dE2 for dt2): one read of P, one table set-up, and the first Edot evaluation shared.
#ifndef PION_COOL_MINB
#define PION_COOL_MINB 8
#endif
template <int EQ, int MODE>
__global__ void __launch_bounds__(128, PION_COOL_MINB) k_cooling_dU(const __grid_constant__ CoolArgs a) {
  extern __shared__ double s_tab[];
  const CoolTab t = cool_tables_to_smem(a.cp, s_tab);
  const GridD& g = a.g;
  const long ncell = (long)g.NG[0] * g.NG[1] * g.NG[2];
  int fails = 0;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += (long)gridDim.x * blockDim.x) {
    const int i = (int)(q % g.NG[0]), j = (int)((q / g.NG[0]) % g.NG[1]), k = (int)(q / ((long)g.NG[0] * g.NG[1]));
    const long c = gidx(g, i + g.nb[0], j + g.nb[1], k + g.nb[2]);
    double dE = 0.0, dE2 = 0.0;
    if (!a.mask || a.mask[c]) {
      // the integration only needs rho and p; the rest of the state (kinetic / magnetic energy of PtoU) is read AFTER
      // it, so that it does not occupy registers across the integrator (the memory clobber keeps the loads there)
      const double ro = __ldg(a.P + c), pg = __ldg(a.P + g.vs + c);
      CoolCell cc;
      cc.rho = ro;
      cc.gm1 = a.gamma - 1.0;
      cc.Mu_tot_over_kB = a.cp.Mu_tot_over_kB;
      const double Eint0 = pg / (a.gamma - 1.0);
      const double k1 = cool_dPdt<MODE>(t, cc, Eint0);
      double Eint, Eint2 = 0.0;
      fails += (cool_integrate<MODE>(t, cc, Eint0, k1, a.dt, Eint) != 0);
      if (a.dE2) fails += (cool_integrate<MODE>(t, cc, Eint0, k1, a.dt2, Eint2) != 0);  // same P, same k1: the corrector's interval
      asm volatile("" ::: "memory");
      // dU += PtoU(p') - PtoU(P): every component but the energy cancels exactly
      const Prim p = load_prim<EQ>(a.P, c, g.vs, 0, 1, 2);
      Cons ui;
      PtoU<EQ>(p, ui, a.gamma - 1.0);
      dE = cool_dE_of<EQ>(a, p, ui, Eint);
      if (a.dE2) dE2 = cool_dE_of<EQ>(a, p, ui, Eint2);
    }
    if (a.dE) a.dE[c] = dE;
    if (a.dE2) a.dE2[c] = dE2;
    if (a.dU) a.dU[g.vs + c] += dE;
  }
  if (fails && a.counters) atomicAdd((unsigned long long*)&a.counters[2], (unsigned long long)fails);
}

// calc_microphysics_dt / get_mp_timescales_no_radiation + mp_only_cooling::timescales
template <int EQ, int MODE>
__global__ void __launch_bounds__(256) k_mp_dt(const __grid_constant__ CoolArgs a) {
  extern __shared__ double s_tab[];
  const CoolTab t = cool_tables_to_smem(a.cp, s_tab);
  const GridD& g = a.g;
  const long ncell = (long)g.NG[0] * g.NG[1] * g.NG[2];
  double my = 1.0e99;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < ncell; q += (long)gridDim.x * blockDim.x) {
    const int i = (int)(q % g.NG[0]), j = (int)((q / g.NG[0]) % g.NG[1]), k = (int)(q / ((long)g.NG[0] * g.NG[1]));
    const long c = gidx(g, i + g.nb[0], j + g.nb[1], k + g.nb[2]);
    if (a.mask && !a.mask[c]) continue;  // isbd cells (internal boundaries) are skipped (:435)
    const double ro = __ldg(a.P + c), pg = __ldg(a.P + g.vs + c);
    const double Eint = pg / (a.gamma - 1.0);
    const double T = pg * a.cp.Mu_tot_over_kB / ro;
    if (T >= 1.1 * a.cp.MinT) {
      const double rate = fmax(fabs(cool_Edot<MODE>(t, ro, T)), fabs(cool_Edot<MODE>(t, ro, fmax(a.cp.MinT, 0.5 * T))));
      my = fmin(my, Eint / rate);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my = fmin(my, __shfl_xor_sync(0xffffffffu, my, o));
  __shared__ double s[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s[w] = my;
  __syncthreads();
  if (w == 0) {
    my = (lane < (int)(blockDim.x >> 5)) ? s[lane] : 1.0e99;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my = fmin(my, __shfl_xor_sync(0xffffffffu, my, o));
    if (lane == 0) atomicMin(a.dtmin, dbl_ordered_bits(my));
  }
}

}  // namespace pion
