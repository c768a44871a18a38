"""CPU: pins the plain-C oracle (oracle/pion_oracle.c) to the UNMODIFIED reference
translation units compiled into oracle/_ref/libpion_ref.so.  The bar is BIT-EXACT
equality of the full padded state (ghost cells included) and of every dt."""
import numpy as np
import pytest

from cases import AVS, EQ_SOLVERS, case_1d, case_2d, case_3d
from harness import OracleSim, RefSim, have_ref, hot_sphere_state, random_state

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")


def run_pair(prob, nsteps=3, seed=1, amp=0.5, state=None):
    r, o = RefSim(prob), OracleSim(prob)
    try:
        assert r.shape() == o.shape() == prob.padded_shape()
        P = state(prob) if state else random_state(prob, seed, amp=amp)
        for s in (r, o):
            s.set_state(P)
            assert s.init_after_state() == 0
        assert np.array_equal(r.get_state(0), o.get_state(0))
        assert np.array_equal(r.get_state(1), o.get_state(1))
        dr, do = r.run(nsteps), o.run(nsteps)
        Pr, Po = r.get_state(0), o.get_state(0)
        assert np.array_equal(dr, do), (dr, do)
        assert np.max(np.abs(Pr - P)) > 1e-6, "state did not evolve"
        assert np.array_equal(Pr, Po), float(np.max(np.abs(Pr - Po)))
        assert o.error_counts()[0] == 0
    finally:
        r.close()
        o.close()


@pytest.mark.parametrize("eqn,solver", EQ_SOLVERS)
@pytest.mark.parametrize("av", AVS)
def test_2d_periodic_bit_exact(eqn, solver, av):
    run_pair(case_2d(eqn, solver, av))


@pytest.mark.parametrize("eqn,solver,av,ntr", [("glm-mhd", 7, 1, 0), ("glm-mhd", 4, 4, 2), ("euler", 4, 3, 1),
                                               ("i-mhd", 7, 0, 1), ("euler", 8, 1, 0)])
def test_3d_periodic_bit_exact(eqn, solver, av, ntr):
    run_pair(case_3d(eqn, solver, av, ntracer=ntr))


@pytest.mark.parametrize("bcs", ["outflow", "reflect-outflow", "mixed1", "mixed2"])
@pytest.mark.parametrize("eqn,solver", [("glm-mhd", 7), ("euler", 8), ("i-mhd", 4)])
def test_boundary_types_bit_exact(bcs, eqn, solver):
    run_pair(case_3d(eqn, solver, 1, bcs=bcs, NG=(10, 8, 6)))
    run_pair(case_2d(eqn, solver, 4, bcs=bcs, ntracer=1, NG=(10, 8, 1)))


@pytest.mark.parametrize("eqn", ["i-mhd", "glm-mhd"])
def test_hlld_to_hll_switch_bit_exact(eqn):
    """x100 pressure ellipsoid: thousands of cells trip the HLLD -> HLL switch (the smooth random state trips none)."""
    run_pair(case_3d(eqn, 7, 1, bcs="reflect-outflow", NG=(20, 16, 12)), state=hot_sphere_state)
    run_pair(case_2d(eqn, 7, 0, bcs="outflow", NG=(48, 40, 1)), state=hot_sphere_state)


@pytest.mark.parametrize("eqn,solver,av", [("euler", 4, 3), ("euler", 8, 1), ("euler", 6, 1), ("euler", 1, 1), ("euler", 2, 0), ("euler", 3, 1),
                                           ("euler", 0, 1), ("i-mhd", 0, 1), ("i-mhd", 4, 1), ("i-mhd", 8, 0),
                                           ("glm-mhd", 4, 4), ("glm-mhd", 8, 1)])
def test_strong_gradients_bit_exact(eqn, solver, av):
    """The x100 pressure ellipsoid for the other solvers (Roe entropy fix / H-correction, HLL, FVS), second and
    first order, next to reflecting walls."""
    run_pair(case_3d(eqn, solver, av, bcs="reflect-outflow", NG=(14, 12, 10)), nsteps=2, state=hot_sphere_state)
    run_pair(case_3d(eqn, solver, av, bcs="mixed2", NG=(9, 7, 5), ooa=1), nsteps=2, state=hot_sphere_state)


@pytest.mark.parametrize("solver", [4, 5, 6, 8, 1, 2, 3])
def test_euler_supersonic_branches_bit_exact(solver):
    """|v| up to 3 (1.5 for the linearised Roe-PV solver: beyond that the reference itself produces NaNs) with c ~ 1: the one-sided (supersonic) branches of FVS, Roe-PV, HLL and the Roe-CV entropy fix."""
    amp = 1.5 if solver in (5, 1) else 3.0  # the linearised solvers produce NaNs in the reference beyond that
    run_pair(case_2d("euler", solver, 1, bcs="outflow"), nsteps=2, amp=amp)
    run_pair(case_3d("euler", solver, 0, ntracer=1), nsteps=2, amp=amp)


def test_exact_riemann_solver_rarefaction_and_cavitation_branches_bit_exact():
    """Strongly diverging flow: (u_R - u_L) beyond the two-rarefaction and the cavitation thresholds of
    riemann_Euler::JMs_riemann_solve (riemann.cpp:322-420), 1-D so that every interface is one of those."""
    import dataclasses
    for solver in (2, 3):
        prob = dataclasses.replace(case_1d("euler", solver, 0, bcs=("outflow", "outflow")), cfl=0.2)

        def diverging(p, amp_v):
            P = random_state(p, 5, amp=0.2)
            x = np.arange(P.shape[3]) - P.shape[3] / 2
            P[2] = amp_v * np.sign(x + 0.5)[None, None, :]  # a jump of 2 amp_v across ONE interface
            return P
        run_pair(prob, nsteps=3, state=lambda p: diverging(p, 3.2))   # two rarefactions: 5.6 < du < 7.7 (c ~ 1.3)
        run_pair(prob, nsteps=2, state=lambda p: diverging(p, 6.0))   # cavitation: du > 3 (c_l + c_r)


def _mhd_linear_state(kind):
    def f(p):
        P = random_state(p, 9, amp=0.3)
        if kind == "tiny":  # nearly uniform: the same-state shortcut of riemann_MHD::JMs_riemann_solve (riemannMHD.cpp:227-268)
            m = P.mean(axis=(1, 2, 3), keepdims=True)
            P = m + (P - m) * 1e-7
            P[1, :, P.shape[2] // 2, P.shape[3] // 2] *= 1.5
        elif kind == "v0":   # stationary contacts: get_pstar averages the left- and right-going solutions (:849-960)
            P[2:5] = 0.0
        elif kind == "bt0":  # no tangential field: beta_y = beta_z = 1/sqrt 2 (:645-655)
            P[6:8] = 0.0
        elif kind == "bx0":  # no normal field along x and y: c_a = 0, c_s = c_a / 2 (:690-700)
            P[5] = 0.0
            P[6] = 0.0
        elif kind == "b0":   # hydrodynamic limit
            P[5:8] = 0.0
        return P
    return f


@pytest.mark.parametrize("eqn", ["i-mhd", "glm-mhd"])
@pytest.mark.parametrize("av", [0, 1])
def test_mhd_linear_riemann_solver_bit_exact(eqn, av):
    """solverType 1 with the MHD equations: riemann_MHD (riemannMHD.cpp), second and first order, strong gradients."""
    run_pair(case_2d(eqn, 1, av, bcs="outflow"), nsteps=3)
    run_pair(case_3d(eqn, 1, av, bcs="mixed1", ntracer=1), nsteps=2)
    run_pair(case_3d(eqn, 1, av, bcs="reflect-outflow", NG=(18, 14, 10)), nsteps=2, state=hot_sphere_state)
    run_pair(case_3d(eqn, 1, av, bcs="mixed2", NG=(9, 7, 5), ooa=1), nsteps=2, state=hot_sphere_state)
    run_pair(case_2d(eqn, 1, av, bcs="outflow"), nsteps=2, amp=2.0)


@pytest.mark.parametrize("eqn", ["i-mhd", "glm-mhd"])
@pytest.mark.parametrize("kind", ["tiny", "v0", "bt0", "bx0", "b0"])
def test_mhd_linear_riemann_solver_degenerate_branches_bit_exact(eqn, kind):
    run_pair(case_2d(eqn, 1, 1, bcs="outflow"), nsteps=3, state=_mhd_linear_state(kind))


def _stone_mhd_blastwave(bfield, eqn, solver):
    """test_problems/MHD_Blastwave2D/params_MHD_blastwave2D_UG_B*_n256.txt at 32 x 48 (Stone's MHD blast wave, periodic)."""
    from harness import Problem
    return Problem(ndim=2, NG=(32, 48, 1), eqn=eqn, solver=solver, artviscosity=1, etav=0.1, gamma=1.6666666666666666666667, cfl=0.3,
                   xmin=(-0.5, -0.75, 0.0), xmax=(0.5, 0.75, 1.0), ics="BlastWave",
                   extra={"BWradius": 0.1, "BWpressure": 0.1, "BWdensity": 1.0, "BWmagfieldX": bfield, "BWmagfieldY": bfield, "BWmagfieldZ": 0.0,
                          "BW_energy": 0.471238898038, "BW_nzones": 3.2, "BW_blast_dens": 1.0, "BW_interface": 1.0e50, "BW_amb2_RO": 0.0,
                          "BW_amb2_PG": 0.0, "BW_amb2_VX": 0.0, "BW_amb2_VY": 0.0, "BW_amb2_VZ": 0.0, "InitIons": "LEAVE"})


@pytest.mark.parametrize("bfield", [0.25066282746310002, 2.5066282746310002, 25.066282746310002])
@pytest.mark.parametrize("eqn,solver", [("glm-mhd", 7), ("i-mhd", 4), ("glm-mhd", 1)])
def test_reference_mhd_blastwave_problem_bit_exact(bfield, eqn, solver):
    """The reference's MHD blast wave (plasma beta from 1.6 down to 1.6e-4) from its own IC class: 15 steps, oracle == reference."""
    prob = _stone_mhd_blastwave(bfield, eqn, solver)
    r = RefSim(prob, run_ics=True)
    P0 = r.get_state(0)
    r.close()
    assert np.max(np.abs(P0[5])) > 0.0
    run_pair(prob, nsteps=15, state=lambda p: P0)


@pytest.mark.parametrize("eqn,solver,av", [("glm-mhd", 4, 1), ("glm-mhd", 7, 1), ("i-mhd", 8, 1), ("glm-mhd", 4, 4)])
def test_reference_mhd_axisymmetric_blastwave_problem_bit_exact(eqn, solver, av):
    """test_problems/blastwave_axi2d/params_MHDaxi2dBW_HalfPlane_NR128.txt at 32 x 16 cells: cylindrical (z, R), axial field with
    plasma beta 1, reflecting axis -- the geometric source terms and area-weighted divergence of the MHD equations."""
    from harness import Problem
    prob = Problem(ndim=2, NG=(32, 16, 1), eqn=eqn, solver=solver, artviscosity=av, etav=0.15, gamma=1.666666666666666666666, cfl=0.2,
                   coords="cylindrical", xmin=(-30.86e18, 0.0, 0.0), xmax=(30.86e18, 30.86e18, 1.0),
                   bcs=("outflow", "outflow", "reflecting", "outflow", "periodic", "periodic"), ics="BlastWave",
                   refvec=(2.34e-22, 1.38e-11, 1e6, 1e6, 1e6, 5e-6, 5e-6, 5e-6, 5e-6) + (1.0,) * 7,
                   extra={"BWpressure": 1.38e-11, "BWdensity": 2.34e-22, "BWmagfieldX": 5.25357e-06, "BWmagfieldY": 0.0, "BWmagfieldZ": 0.0,
                          "BW_energy": 1.0e51, "BW_nzones": 2, "BW_blast_dens": 2.34e-22, "BW_interface": 1.0e50, "BW_amb2_RO": 0.0,
                          "BW_amb2_PG": 0.0, "BW_amb2_VX": 0.0, "BW_amb2_VY": 0.0, "BW_amb2_VZ": 0.0, "InitIons": "LEAVE"})
    r = RefSim(prob, run_ics=True)
    P0 = r.get_state(0)
    r.close()
    assert np.max(np.abs(P0[5])) > 0.0
    run_pair(prob, nsteps=15, state=lambda p: P0)


@pytest.mark.parametrize("solver,av", [(4, 1), (8, 1), (4, 4), (6, 1)])
def test_reference_oblique_shock_problem_bit_exact(solver, av):
    """test_problems/ObliqueShock/params_oblique_shock_M25.txt at 50 x 25: a Mach-25 shock at 2 degrees to the grid from the
    reference's ShockTube IC class (custom pre / post-shock states, cgs units), outflow + FIXED boundaries."""
    from harness import Problem
    prob = Problem(ndim=2, NG=(50, 25, 1), eqn="euler", solver=solver, artviscosity=av, etav=0.15, gamma=1.666666666666666666, cfl=0.4,
                   xmax=(1.0e17, 0.5e17, 1.0), bcs=("outflow", "fixed", "outflow", "outflow", "periodic", "periodic"), ics="ShockTube",
                   refvec=(1.0e-22, 1.0e-12, 1.0e6, 1.0e6, 1.0e6) + (1.0,) * 11,
                   extra={"STnumber": -7, "STangleXY": 2.0, "STangleXZ": 0.0, "STshockpos": 4.0e16,
                          "STpostvecRO": 3.9808917197e-22, "STpostvecPG": 7.8100000000e-10, "STpostvecVX": -8.2074451397e05,
                          "STpostvecVY": 0.0, "STpostvecVZ": 0.0, "STprevecRO": 1.0e-22, "STprevecPG": 1.0e-12,
                          "STprevecVX": -3.237486122e6, "STprevecVY": 0.0, "STprevecVZ": 0.0})
    r = RefSim(prob, run_ics=True)
    P0 = r.get_state(0)
    r.close()
    assert P0[0].max() > 3.0e-22 and P0[0].min() < 1.1e-22
    run_pair(prob, nsteps=15, state=lambda p: P0)


@pytest.mark.parametrize("eqn,solver,av", [("euler", 8, 1), ("glm-mhd", 7, 1), ("i-mhd", 4, 0)])
def test_two_tracers_bit_exact(eqn, solver, av):
    run_pair(case_3d(eqn, solver, av, bcs="mixed1", ntracer=2, NG=(12, 10, 8)))


@pytest.mark.parametrize("eqn,solver", [("euler", 8), ("euler", 4), ("i-mhd", 7), ("i-mhd", 4)])
def test_negative_pressure_floor_path_bit_exact(eqn, solver):
    """Cold, highly supersonic random flow (p ~ 1e-7 rho v^2): the conservative update leaves dozens to hundreds of cells with a
    negative pressure, which UtoP resets (SET_NEGATIVE_PRESSURE_TO_FIXED_TEMPERATURE, eqns_hydro_adiabatic.cpp:117-205,
    eqns_mhd_adiabatic.cpp:110-224) -- oracle == reference on that path, and the path is really taken."""
    def cold(p):
        P = random_state(p, 3, amp=3.0)
        P[1] *= 1e-7
        return P
    prob = case_2d(eqn, solver, 1, bcs="outflow")
    run_pair(prob, nsteps=3, state=cold)
    o = OracleSim(prob)
    o.set_state(cold(prob))
    o.init_after_state()
    o.run(3)
    counts = o.error_counts()
    o.close()
    assert counts[0] == 0 and counts[1] > 10, counts


def test_1d_and_first_order():
    run_pair(case_1d("i-mhd", 7, 1))
    run_pair(case_1d("euler", 8, 1, bcs=("reflecting", "inflow")))
    run_pair(case_3d("glm-mhd", 7, 1, ooa=1))


def test_per_call_seam_bit_exact():
    """The grid-level seam one call at a time: dU after calc_dynamics_dU, Ph after the
    half-step update, ghost cells after TimeUpdateBCs, the HLLD switch scalars."""
    prob = case_3d("glm-mhd", 7, 1, bcs="reflect-outflow")
    r, o = RefSim(prob), OracleSim(prob)
    P = random_state(prob, 3)
    for s in (r, o):
        s.set_state(P)
        s.init_after_state()
    dt = r.dynamics_dt()
    assert dt == o.dynamics_dt()
    for s in (r, o):
        s.set_glm_speeds(dt, prob.dx, 0.25 / prob.dx)
        s.set_dt(0.5 * dt)
        s.dynamics_dU(0.5 * dt, 1)
    assert np.array_equal(r.get_state(2), o.get_state(2))
    assert np.array_equal(r.get_extra(0), o.get_extra(0))  # divV
    assert np.array_equal(r.get_extra(1), o.get_extra(1))  # |grad p|/p
    for s in (r, o):
        s.update_state(0.5 * dt, 1, 2)
        s.update_bcs(1, 2)
    assert np.array_equal(r.get_state(1), o.get_state(1))
    for s in (r, o):
        s.set_dt(dt)
        s.dynamics_dU(dt, 2)
    assert np.array_equal(r.get_state(2), o.get_state(2))
    r.close()
    o.close()


# ------------------------------------------------------------------------------------------
# cooling source term (mp_only_cooling, EP_cooling 8): tables taken from the reference's own
# rate functions, then the whole step -- adaptive RKCK integration included -- is bit-exact.
from cases import case_cooling  # noqa: E402
from harness import cooling_state  # noqa: E402


@pytest.mark.parametrize("eqn,solver,ndim,NG,ntr,lim,rho0", [
    ("euler", 8, 3, (12, 10, 8), 1, 1, 2.0e-24),
    ("euler", 8, 3, (12, 10, 8), 1, 1, 2.0e-21),   # cooling time < CFL step: MP limit + bisection
    ("glm-mhd", 7, 3, (12, 10, 8), 0, 2, 2.0e-22),
    ("euler", 4, 2, (16, 12, 1), 1, 0, 2.0e-21),
    ("i-mhd", 8, 2, (16, 12, 1), 2, 4, 2.0e-22),
])
def test_cooling_bit_exact(eqn, solver, ndim, NG, ntr, lim, rho0):
    prob = case_cooling(eqn, solver, ndim=ndim, NG=NG, ntracer=ntr, mp_limit=lim)
    r = RefSim(prob)
    tab = r.cooling_tables()
    o = OracleSim(prob, tables=tab)
    try:
        P = cooling_state(prob, seed=11, rho0=rho0)
        for s in (r, o):
            s.set_state(P)
            assert s.init_after_state() == 0
        assert r.microphysics_dt() == o.microphysics_dt()
        dr, do = r.run(3), o.run(3)
        assert np.array_equal(dr, do), (dr, do)
        assert np.array_equal(r.get_state(0), o.get_state(0))
        # the seam call on its own: dU after calc_microphysics_dU
        assert r.microphysics_dU(0.5 * dr[-1]) == 0 and o.microphysics_dU(0.5 * dr[-1]) == 0
        dUr, dUo = r.get_state(2), o.get_state(2)
        assert np.max(np.abs(dUr[1])) > 0
        assert np.array_equal(dUr, dUo)
    finally:
        r.close()
        o.close()


# the other cooling functions of mp_only_cooling::Edot (mp_only_cooling.cpp:383-420): KI02 (2), SD93_CIE (4),
# SD93_PLUS_HEATING (5), WSS09_CIE_PLUS_HEATING (6), WSS09_CIE_ONLY_COOLING (7).  (DMcC, 3, has no case in Edot:
# the reference aborts with "bad cooling flag".)  The spline knots come from the reference's own MP object.
COOLING_FLAG_CASES = [
    # flag, eqn, solver, rho0, Tlo, Thi, Tmin, Tmax
    (2, "euler", 8, 2.0e-22, 20.0, 2.0e4, 10.0, 1.0e5),
    (4, "euler", 8, 2.0e-23, 2.0e4, 5.0e7, 1.0e4, 1.0e8),
    (5, "glm-mhd", 7, 2.0e-22, 6.0e3, 5.0e7, 5.0e3, 1.0e8),
    (6, "euler", 4, 2.0e-22, 6.0e3, 5.0e7, 5.0e3, 1.0e8),
    (7, "i-mhd", 8, 2.0e-23, 2.0e4, 5.0e7, 1.0e4, 1.0e8),
]


def cooling_flag_problem(flag, eqn, solver, Tmin, Tmax, NG=(12, 10, 8)):
    import dataclasses
    return dataclasses.replace(case_cooling(eqn, solver, ndim=3, NG=NG, ntracer=1 if eqn == "euler" else 0, mp_limit=1),
                               cooling=flag, min_temperature=Tmin, max_temperature=Tmax)


@pytest.mark.parametrize("flag,eqn,solver,rho0,Tlo,Thi,Tmin,Tmax", COOLING_FLAG_CASES)
def test_cooling_functions_bit_exact(flag, eqn, solver, rho0, Tlo, Thi, Tmin, Tmax):
    prob = cooling_flag_problem(flag, eqn, solver, Tmin, Tmax)
    r = RefSim(prob)
    tab = r.cooling_spline() if flag != 2 else None
    o = OracleSim(prob, tables=tab)
    try:
        P = cooling_state(prob, seed=23 + flag, rho0=rho0, Tlo=Tlo, Thi=Thi)
        for s in (r, o):
            s.set_state(P)
            assert s.init_after_state() == 0
        assert r.microphysics_dt() == o.microphysics_dt()
        dr, do = r.run(3), o.run(3)
        assert np.array_equal(dr, do), (dr, do)
        assert np.array_equal(r.get_state(0), o.get_state(0))
        assert r.microphysics_dU(0.5 * dr[-1]) == 0 and o.microphysics_dU(0.5 * dr[-1]) == 0
        dUr, dUo = r.get_state(2), o.get_state(2)
        assert np.max(np.abs(dUr[1])) > 0  # the source term does something on this state
        assert np.array_equal(dUr, dUo)
    finally:
        r.close()
        o.close()


def test_committed_cooling_splines_match_reference():
    from harness import load_cooling_spline
    for flag in (4, 5, 6, 7):
        r = RefSim(cooling_flag_problem(flag, "euler", 8, 5.0e3, 1.0e8))
        sp, gold = r.cooling_spline(), load_cooling_spline(flag)
        r.close()
        for k in sp:
            assert np.array_equal(sp[k], gold[k]), (flag, k)


def test_committed_cooling_tables_match_reference():
    from harness import TABLE_KEYS, load_cooling_tables
    r = RefSim(case_cooling())
    tab, gold = r.cooling_tables(), load_cooling_tables()
    r.close()
    for k in TABLE_KEYS:
        assert np.array_equal(tab[k], gold[k]), k


# ------------------------------------------------------------------------------------------
# stellar-wind internal boundary (constant source), with and without cooling
from cases import case_wind, wind_ambient_state  # noqa: E402


@pytest.mark.parametrize("eqn,solver,ndim,NG,cooling,vrot", [
    ("euler", 8, 3, (16, 16, 16), True, 0.0),
    ("glm-mhd", 7, 3, (12, 12, 12), False, 20.0),
    ("i-mhd", 8, 2, (24, 24, 1), True, 20.0),
    ("euler", 4, 2, (24, 24, 1), False, 0.0),
])
def test_stellar_wind_bit_exact(eqn, solver, ndim, NG, cooling, vrot):
    prob = case_wind(eqn, solver, ndim=ndim, NG=NG, cooling=cooling, vrot=vrot)
    r = RefSim(prob)
    o = OracleSim(prob, tables=r.cooling_tables() if cooling else None)
    try:
        P = wind_ambient_state(prob)
        for s in (r, o):
            s.set_state(P)
            assert s.init_after_state() == 0
        assert np.array_equal(r.get_flags(), o.get_flags())  # isbd / isdomain / timestep of the wind cells
        assert np.array_equal(r.get_state(0), o.get_state(0))
        assert r.dynamics_dt() == o.dynamics_dt()  # first-step wind limit (calc_timestep.cpp:318-323)
        dr, do = r.run(5), o.run(5)
        assert np.array_equal(dr, do)
        assert np.array_equal(r.get_state(0), o.get_state(0))
        assert np.array_equal(r.get_state(1), o.get_state(1))
    finally:
        r.close()
        o.close()


# ------------------------------------------------------------------------------------------
# curvilinear grids: 2-D axisymmetric (z,R) and 1-D spherical geometric source terms (a12)
from cases import case_cyl, case_sph  # noqa: E402

GEOM_CASES = [
    case_cyl("euler", 8, 1), case_cyl("euler", 4, 4, ntracer=1), case_cyl("i-mhd", 8, 1), case_cyl("i-mhd", 7, 1),
    case_cyl("glm-mhd", 7, 1), case_cyl("glm-mhd", 4, 3, ntracer=1), case_cyl("glm-mhd", 7, 1, ooa=1),
    case_cyl("euler", 8, 0, bcs=("periodic", "periodic", "reflecting", "reflecting")),
    case_sph(8, 1), case_sph(4, 1, rmin=1.0, bcs=("inflow", "outflow"), ntracer=1), case_sph(8, 0, ooa=1),
]


@pytest.mark.parametrize("prob", GEOM_CASES, ids=lambda p: f"{p.coords[:3]}-{p.eqn}-s{p.solver}-av{p.artviscosity}-oa{p.ooa}")
def test_curvilinear_bit_exact(prob):
    run_pair(prob)
