"""The DEVICE arithmetic checked on the CPU: pion_b200/csrc/physics.cuh + fastmath.cuh (the product's Riemann solvers, one-sided
HLLD, reciprocal-multiply forms, Newton-refined MUFU seeds) are compiled for the host through tests/host_physics/shim.h and
`intercell_flux<EQ, SOLVER, AV>` is compared with the oracle's InterCellFlux interface by interface -- equation sets x solvers x
viscosities x sweep axes, smooth, strong-gradient, supersonic and cold (plasma beta 1e-7) states.  No GPU needed; the GPU parity
tests remain the proof for the kernels around this header.  TEST INFRASTRUCTURE: the host build exists only inside this test."""
import ctypes as C
import math
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from cases import case_3d
from harness import OracleSim

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "pion_b200" / "csrc"
HERE = Path(__file__).resolve().parent / "host_physics"
EQN = {"euler": 1, "i-mhd": 2, "glm-mhd": 3}
CHYP = 1.3


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    """g++ build of the product's physics header: the asm seeds and the predicated-DFMA limiter are the only lines replaced."""
    d = tmp_path_factory.mktemp("host_physics")
    fm = (CSRC / "fastmath.cuh").read_text()
    fm = fm.replace("#include <cuda_runtime.h>", '#include "shim.h"')
    fm = fm.replace('asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));', "r = pion_rcp_approx(x);")
    fm = fm.replace('asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));', "y = pion_rsqrt_approx(x);")
    assert "asm(" not in fm
    ph = (CSRC / "physics.cuh").read_text()
    ph = ph.replace("#include <cuda_runtime.h>", '#include "shim.h"')
    a = ph.index('  asm("{\\n\\t.reg .pred p;')
    b = ph.index("__double2hiint(b)));", a) + len("__double2hiint(b)));")
    ph = ph[:a] + "  if ((__double2hiint(a) ^ __double2hiint(b)) >= 0) e = fma(m, h, e);" + ph[b:]
    assert "asm(" not in ph
    (d / "fastmath.cuh").write_text(fm)
    (d / "physics.cuh").write_text(ph)
    for f in ("shim.h", "host_flux.cpp"):
        shutil.copy(HERE / f, d / f)
    so = d / "libhost_physics.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=fast", "-I", str(d), str(d / "host_flux.cpp"), "-o", str(so)],
                   check=True)
    lib = C.CDLL(str(so))
    lib.host_intercell_flux.restype = C.c_int
    lib.host_intercell_flux.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3 + [C.c_int, C.c_double, C.c_int, C.c_void_p]
    return lib


def solver_refvec(prob):
    """What the host side of the library puts into PhysParams::rs_refvec (pion_b200.cu; eqns_*::SetAvgState)."""
    rv = list(prob.refvec[:8])
    g = prob.gamma
    if prob.eqn == "euler":
        c = 0.1 * math.sqrt(g * rv[1] / rv[0])
        return [rv[0], rv[1], c, c, c]

    def cfast(v):
        ch = math.sqrt(g * v[1] / v[0])
        t1 = ch * ch + (v[5] * v[5] + v[6] * v[6] + v[7] * v[7]) / v[0]
        t2 = max(5e-16, t1 * t1 - 4.0 * ch * ch * v[5] * v[5] / v[0])
        return math.sqrt((t1 + math.sqrt(t2)) / 2.0)

    def rot(v, th):
        ct, st = math.cos(th), math.sin(th)
        v[2], v[3] = v[2] * ct - v[3] * st, v[2] * st + v[3] * ct
        v[5], v[6] = v[5] * ct - v[6] * st, v[5] * st + v[6] * ct
    ang = rv[6] * rv[6] + rv[5] * rv[5]
    if ang > 10 * 5e-16:
        ang = math.pi / 2.0 - math.asin(rv[6] / math.sqrt(ang))
        if rv[5] < 0:
            ang = -ang
        rot(rv, ang)
        vel = cfast(rv)
        rot(rv, -ang)
    else:
        vel = cfast(rv)
    return [prob.refvec[0], prob.refvec[1], 0.1 * vel, math.sqrt(rv[5] ** 2 + rv[6] ** 2 + rv[7] ** 2), 0.0]


def interfaces(kind, n, nphys, rng):
    """n random left / right grid-frame primitive states [ro, pg, vx, vy, vz, (bx, by, bz, (psi))]."""
    L = np.zeros((n, 9))
    R = np.zeros((n, 9))
    for S in (L, R):
        S[:, 0] = rng.uniform(0.5, 1.5, n)
        S[:, 1] = rng.uniform(0.5, 1.5, n)
        S[:, 2:5] = rng.uniform(-0.5, 0.5, (n, 3))
        S[:, 5:8] = rng.uniform(-0.7, 0.7, (n, 3))
        S[:, 8] = rng.uniform(-0.1, 0.1, n)
    if kind == "smooth":
        R[:] = L + 1e-3 * (R - L)
    elif kind == "strong":
        R[:, 1] *= 100.0
        R[:, 0] *= 5.0
    elif kind == "supersonic":
        L[:, 2:5] *= 6.0
        R[:, 2:5] *= 6.0
    elif kind == "cold":
        L[:, 2:5] *= 6.0
        R[:, 2:5] *= 6.0
        L[:, 1] *= 1e-7
        R[:, 1] *= 1e-7
    if nphys < 9:
        L[:, nphys:] = 0.0
        R[:, nphys:] = 0.0
    return L, R


def frame(S, ax):
    """grid frame -> solver frame of axis ax."""
    a1, a2 = (ax + 1) % 3, (ax + 2) % 3
    return np.array([S[0], S[1], S[2 + ax], S[2 + a1], S[2 + a2], S[5 + ax], S[5 + a1], S[5 + a2], S[8]])


def unframe_flux(F, ax):
    """solver-frame flux -> grid-frame conserved order [rho, erg, mx, my, mz, bx, by, bz, psi]."""
    a1, a2 = (ax + 1) % 3, (ax + 2) % 3
    out = np.zeros(9)
    out[0], out[1], out[8] = F[0], F[1], F[8]
    out[2 + ax], out[2 + a1], out[2 + a2] = F[2], F[3], F[4]
    out[5 + ax], out[5 + a1], out[5 + a2] = F[5], F[6], F[7]
    return out


CASES = [("euler", s) for s in (1, 2, 3, 4, 5, 6, 8)] + [(e, s) for e in ("i-mhd", "glm-mhd") for s in (1, 4, 7, 8)]


@pytest.mark.parametrize("eqn,solver", CASES)
@pytest.mark.parametrize("av", [0, 1, 3, 4])
def test_host_build_of_device_flux_matches_oracle(hostlib, eqn, solver, av):
    prob = case_3d(eqn, solver, av)
    nphys = {"euler": 5, "i-mhd": 8, "glm-mhd": 9}[eqn]
    o = OracleSim(prob)
    etav = 0.0 if av == 0 else 0.1 if av == 3 else prob.etav  # ics/get_sim_info.cpp:452-468, as pion_gpu_create applies it
    if eqn == "glm-mhd":
        o.set_glm_speeds(prob.cfl * prob.dx / CHYP, prob.dx, 0.25 / prob.dx)
    par = np.array([prob.gamma, etav, CHYP if eqn == "glm-mhd" else 0.0, prob.refvec[0]] + solver_refvec(prob))
    rng = np.random.default_rng(2024 + 31 * solver + EQN[eqn])
    flux = np.zeros(9)
    worst = 0.0
    kinds = ["smooth", "random", "strong"]
    if solver not in (1, 5):
        kinds.append("supersonic")   # the linearised solvers produce NaNs in the reference beyond |v| ~ 1.5 c
    if solver in (7, 8):
        kinds.append("cold")
    try:
        for kind in kinds:
            L, R = interfaces(kind, 400, nphys, rng)
            for q in range(L.shape[0]):
                ax = q % 3
                hll = (q % 5 == 0) and solver == 7
                hc = 0.05 * (q % 4) if av in (3, 4) else 0.0
                Fo = np.zeros(9)
                Fo[:nphys] = o.intercell_flux(ax, L[q, :nphys], R[q, :nphys], divv_l=-1.0 if hll else 0.0, gradp_l=10.0 if hll else 0.0,
                                              hc_etamax=hc)[:nphys]
                l, r = frame(L[q], ax), frame(R[q], ax)
                rc = hostlib.host_intercell_flux(EQN[eqn], solver, av, l.ctypes.data, r.ctypes.data, par.ctypes.data, int(hll), hc, ax,
                                                 flux.ctypes.data)
                assert rc == 0, (kind, q, rc)
                Fh = unframe_flux(flux, ax)
                assert np.all(np.isfinite(Fo)) and np.all(np.isfinite(Fh)), (kind, q, Fo, Fh)
                err = np.max(np.abs(Fh - Fo)) / max(np.max(np.abs(Fo)), 1e-300)
                worst = max(worst, err)
                assert err < 2e-12, (kind, q, ax, err, Fo, Fh)
    finally:
        o.close()
    print(f"{eqn} solver {solver} av {av}: worst relative flux difference {worst:.2e}")


def test_hlld_unused_side_negative_star_density_regression(hostlib):
    """Interfaces from the negative-pressure-reset probe (plasma beta 1e-7): the contact speed lies beyond the fast wave of the
    right side, that side's star density is negative and its square root a NaN.  The reference selects the LEFT single-star state
    and never forms the ** states; the one-sided device form used to multiply the ** jumps by c1 = 0 (0 x NaN) and returned a
    non-finite flux.  Fixed in mhd_HLLD; these vectors pin it."""
    prob = case_3d("i-mhd", 7, 1)
    vectors = [
        ([8.02167195e-01, 8.74318822e-08, 1.47116692e-01, -6.65557044e-01, -6.65526018e-01, 7.60160121e-01, -1.00847968e-01, -1.21214418e-01],
         [7.77773319e-01, 8.51360648e-08, 2.68643888e-01, -3.17182320e-01, -1.19822917e+00, -2.35631974e-01, -8.02817697e-02, 1.60948595e-01]),
        ([1.09670003e+00, 9.79794704e-08, -1.14105813e-01, -4.08981731e-01, -7.43479941e-01, 6.73442335e-01, 3.78896631e-03, 1.34951565e-01],
         [1.06576731e+00, 9.80919396e-08, 1.49729881e-01, -2.69129035e-01, -1.30139325e+00, -3.30327084e-01, 5.89291423e-02, -3.16928819e-02]),
    ]
    o = OracleSim(prob)
    par = np.array([prob.gamma, prob.etav, 0.0, prob.refvec[0]] + solver_refvec(prob))
    flux = np.zeros(9)
    try:
        for Lg, Rg in vectors:
            L, R = np.array(Lg + [0.0]), np.array(Rg + [0.0])
            Fo = np.zeros(9)
            Fo[:8] = o.intercell_flux(0, L[:8], R[:8])[:8]
            rc = hostlib.host_intercell_flux(2, 7, 1, L.ctypes.data, R.ctypes.data, par.ctypes.data, 0, 0.0, 0, flux.ctypes.data)
            assert rc == 0 and np.all(np.isfinite(flux)), flux
            assert np.max(np.abs(flux - Fo)) / np.max(np.abs(Fo)) < 2e-12, (flux, Fo)
    finally:
        o.close()
