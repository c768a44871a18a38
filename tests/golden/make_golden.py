"""Generates the golden vectors in this directory FROM THE COMPILED REFERENCE
(oracle/_ref/libpion_ref.so = unmodified PION translation units).  Run in the build
container (needs /root/reference at build time):

    python tests/golden/make_golden.py

Each fixture holds the seeded input state, the per-step dt sequence and the full padded
state after N steps; tests/test_golden.py replays them through the plain-C oracle
(bit-exact) and, on the GPU box, through the CUDA path (tolerance)."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from cases import case_1d, case_2d, case_3d, case_cooling, case_cyl, case_sph, case_wind, wind_ambient_state  # noqa: E402
from harness import COOLING_SPLINES, COOLING_TABLES, TABLE_KEYS, RefSim, cooling_state, hot_sphere_state, random_state  # noqa: E402

CASES = {
    "glm_hlld_fkj_3d_periodic": (case_3d("glm-mhd", 7, 1, NG=(12, 10, 8)), 4),
    "glm_hlld_fkj_3d_outflow": (case_3d("glm-mhd", 7, 1, bcs="outflow", NG=(12, 10, 8)), 4),
    "glm_roe_hcorr_2d_periodic": (case_2d("glm-mhd", 4, 4), 4),
    "imhd_hll_2d_reflect": (case_2d("i-mhd", 8, 1, bcs="reflect-outflow"), 4),
    "euler_roe_fkj_3d_mixed": (case_3d("euler", 4, 1, bcs="mixed1", ntracer=1, NG=(10, 8, 6)), 4),
    "euler_hll_1d_reflect_inflow": (case_1d("euler", 8, 1, bcs=("reflecting", "inflow")), 6),
    # curvilinear grids: 2-D axisymmetric (z,R) and 1-D spherical
    "cyl_glm_hlld_fkj_2d": (case_cyl("glm-mhd", 7, 1), 4),
    "cyl_euler_roe_hcorr_2d_tracer": (case_cyl("euler", 4, 4, ntracer=1), 4),
    "sph_euler_hll_1d": (case_sph(8, 1), 6),
    # Euler flux options FVS (6, van Leer) and Roe-PV (5)
    "euler_fvs_fkj_3d_tracer": (case_3d("euler", 6, 1, bcs="reflect-outflow", ntracer=1, NG=(12, 10, 8)), 4),
    "euler_roepv_fkj_2d_outflow": (case_2d("euler", 5, 1, bcs="outflow"), 4),
}

# round 2: Lax-Friedrichs (0), the Euler linear / exact / hybrid Riemann solvers (1-3), the MHD linear Riemann solver (1),
# and two tracers on 3-D multi-tile grids (extra tile variables of the TMA sweep kernel)
CASES.update({
    "euler_lf_2d_outflow": (case_2d("euler", 0, 1, bcs="outflow"), 4),
    "imhd_lf_3d_mixed": (case_3d("i-mhd", 0, 0, bcs="mixed1", NG=(10, 8, 6)), 4),
    "euler_rslinear_fkj_2d_outflow": (case_2d("euler", 1, 1, bcs="outflow"), 4),
    "euler_rsexact_3d_tracer": (case_3d("euler", 2, 0, bcs="reflect-outflow", ntracer=1, NG=(10, 8, 6)), 3),
    "euler_rshybrid_fkj_2d_reflect": (case_2d("euler", 3, 1, bcs="reflect-outflow"), 4),
    "imhd_rslinear_fkj_3d_mixed": (case_3d("i-mhd", 1, 1, bcs="mixed1", NG=(10, 8, 6)), 4),
    "glm_rslinear_2d_outflow": (case_2d("glm-mhd", 1, 0, bcs="outflow"), 4),
    "glm_hlld_fkj_3d_two_tracers_tiles": (case_3d("glm-mhd", 7, 1, bcs="mixed1", ntracer=2, NG=(36, 14, 10)), 3),
    "euler_hll_fkj_3d_two_tracers_tiles": (case_3d("euler", 8, 1, bcs="reflect-outflow", ntracer=2, NG=(36, 14, 10)), 3),
})

# x100 pressure ellipsoid (harness.hot_sphere_state): thousands of cells trip the HLLD -> HLL switch, next to
# reflecting walls where the ideal-MHD HLLD contact speed is an exact zero; the second case spans 2 x 2 tiles and
# 2 z chunks of the TMA sweep kernel
HOT = {
    "imhd_hlld_3d_hot_sphere_reflect": (case_3d("i-mhd", 7, 0, bcs="reflect-outflow", NG=(20, 16, 12)), 3),
    "glm_hlld_fkj_3d_hot_sphere_tiles": (case_3d("glm-mhd", 7, 1, bcs="reflect-outflow", NG=(36, 14, 10)), 2),
}
CASES.update(HOT)


# The reference's own test problems, initial conditions produced by the reference's IC
# classes (ics/basic_tests.cpp:553-812, ics/blast_wave.cpp:626-695) at reduced resolution.
from harness import Problem  # noqa: E402

TEST_PROBLEMS = {
    # test_problems/double_Mach_reflection/params_DMR_n065.txt
    "tp_DMR_n065_roe": (Problem(ndim=2, NG=(65, 20, 1), eqn="euler", solver=4, artviscosity=1, etav=0.1, gamma=1.4, cfl=0.4,
                                xmax=(3.25, 1.0, 1.0), bcs=("inflow", "outflow", "reflecting", "DMR", "periodic", "periodic"),
                                internal_bcs=("DMR2",), ics="DoubleMachRef", finishtime=0.2, op_criterion=1, opfreq_time=0.05,
                                extra={"DMRmach": 10.0, "DMRtheta": 60}), 12),
    "tp_DMR_n065_hll": (Problem(ndim=2, NG=(65, 20, 1), eqn="euler", solver=8, artviscosity=1, etav=0.1, gamma=1.4, cfl=0.4,
                                xmax=(3.25, 1.0, 1.0), bcs=("inflow", "outflow", "reflecting", "DMR", "periodic", "periodic"),
                                internal_bcs=("DMR2",), ics="DoubleMachRef", finishtime=0.2, op_criterion=1, opfreq_time=0.05,
                                extra={"DMRmach": 10.0, "DMRtheta": 60}), 12),
    # test_problems/FieldLoop/params_FieldLoop200.txt at 64x32, HLLD (7) and the shipped Roe (4)
    "tp_FieldLoop_64x32_hlld": (Problem(ndim=2, NG=(64, 32, 1), eqn="glm-mhd", solver=7, artviscosity=1, etav=0.1,
                                        gamma=1.666666666666666666666, cfl=0.4, xmin=(-1.0, -0.5, 0.0), xmax=(1.0, 0.5, 1.0),
                                        ics="FieldLoop", finishtime=2.0, op_criterion=1, opfreq_time=1.0), 12),
    "tp_FieldLoop_64x32_roe": (Problem(ndim=2, NG=(64, 32, 1), eqn="glm-mhd", solver=4, artviscosity=1, etav=0.1,
                                       gamma=1.666666666666666666666, cfl=0.4, xmin=(-1.0, -0.5, 0.0), xmax=(1.0, 0.5, 1.0),
                                       ics="FieldLoop", finishtime=2.0, op_criterion=1, opfreq_time=1.0), 12),
    # test_problems/blastwave_crt3d/params_BWcrt3D_Octant_NR016.txt
    "tp_BWcrt3D_octant_n016": (Problem(ndim=3, NG=(16, 16, 16), eqn="euler", solver=4, artviscosity=1, etav=0.1,
                                       gamma=1.666666666666666666666, cfl=0.3, xmax=(30.86e18,) * 3,
                                       bcs=("reflecting", "outflow") * 3, ics="BlastWave", finishtime=1.58e12, op_criterion=1,
                                       opfreq_time=6.32e10,
                                       extra={"BWpressure": 1.38e-11, "BWdensity": 2.34e-22, "BWmagfieldX": 0.0, "BWmagfieldY": 0.0,
                                              "BWmagfieldZ": 0.0, "BW_energy": 1.0e51, "BW_nzones": 2, "BW_blast_dens": 2.34e-22,
                                              "BW_interface": 1.0e50, "BW_amb2_RO": 0.0, "BW_amb2_PG": 0.0, "BW_amb2_VX": 0.0,
                                              "BW_amb2_VY": 0.0, "BW_amb2_VZ": 0.0, "InitIons": "LEAVE"}), 12),
}


# The reference's shock-tube known-answer INPUTS (ics/shock_tube.cpp:473-810): Toro's tests 1-5 (Euler), Brio & Wu, Falle's fast /
# slow shocks and Ryu & Jones 1a (ideal MHD), 1-D, 128 cells, 40 steps, states written by the reference's own IC class
def _shock_tube(num, eqn, solver, gamma, av=1):
    return Problem(ndim=1, NG=(128, 1, 1), eqn=eqn, solver=solver, artviscosity=av, etav=0.1, gamma=gamma, cfl=0.4, xmax=(1.0, 1.0, 1.0),
                   bcs=("outflow", "outflow") + ("periodic",) * 4, ics="ShockTube",
                   extra={"STnumber": num, "STshockpos": 0.5, "STangleXY": 0, "STangleXZ": 0})


SHOCK_TUBES = {
    "st_toro1_euler_hll": (_shock_tube(1, "euler", 8, 1.4), 40),
    "st_toro2_euler_exact": (_shock_tube(2, "euler", 2, 1.4, av=0), 40),
    "st_toro3_euler_roe": (_shock_tube(3, "euler", 4, 1.4), 40),
    "st_toro4_euler_hybrid": (_shock_tube(4, "euler", 3, 1.4), 40),
    "st_toro5_euler_fvs": (_shock_tube(5, "euler", 6, 1.4), 40),
    "st_briowu_imhd_hlld": (_shock_tube(7, "i-mhd", 7, 2.0), 40),
    "st_briowu_imhd_linear": (_shock_tube(7, "i-mhd", 1, 2.0), 40),
    "st_falle_fs_imhd_roe": (_shock_tube(9, "i-mhd", 4, 5.0 / 3.0), 40),
    "st_falle_ss_imhd_hll": (_shock_tube(10, "i-mhd", 8, 5.0 / 3.0), 40),
    "st_ryujones1a_imhd_hlld": (_shock_tube(15, "i-mhd", 7, 5.0 / 3.0, av=0), 40),
}
# more of the reference's own test problems (ics/basic_tests.cpp): Liska & Wendroff implosion (test_problems/LiskaWendroffImplosion,
# reflecting walls), Orszag-Tang vortex (:680-735) and Stone's Kelvin-Helmholtz set-up (:814-870; the IC class forces gamma = 1.4
# and seeds its noise with srand(975)), at 48 x 48
MORE_TEST_PROBLEMS = {
    "tp_LWI_n048_roe": (Problem(ndim=2, NG=(48, 48, 1), eqn="euler", solver=4, artviscosity=0, etav=0.15, gamma=1.4, cfl=0.3,
                                xmax=(0.3, 0.3, 1.0), bcs=("reflecting",) * 4 + ("periodic",) * 2, ics="LiskaWendroffImplosion"), 20),
    "tp_OrszagTang_n048_glm_hlld": (Problem(ndim=2, NG=(48, 48, 1), eqn="glm-mhd", solver=7, artviscosity=1, etav=0.1, gamma=5.0 / 3.0, cfl=0.4,
                                            xmax=(1.0, 1.0, 1.0), ics="OrszagTang", extra={"OTVbeta": 3.333333333, "OTVmach": 1.0}), 20),
    "tp_KHStone_n048_euler_hll": (Problem(ndim=2, NG=(48, 48, 1), eqn="euler", solver=8, artviscosity=1, etav=0.1, gamma=1.4, cfl=0.4,
                                          xmin=(-0.5, -0.5, 0.0), xmax=(0.5, 0.5, 1.0), ics="KelvinHelmholzStone"), 20),
    "tp_KHStone_n048_imhd_hlld": (Problem(ndim=2, NG=(48, 48, 1), eqn="i-mhd", solver=7, artviscosity=1, etav=0.1, gamma=1.4, cfl=0.4,
                                          xmin=(-0.5, -0.5, 0.0), xmax=(0.5, 0.5, 1.0), ics="KelvinHelmholzStone"), 20),
}
# the reference's curvilinear blast waves: test_problems/blastwave_axi2d/params_axi2dBW_HalfPlane_NR016.txt (cylindrical (z,R),
# Roe-CV) and test_problems/blastwave_sph1d/params_sphBW_n128.txt (1-D spherical, the shipped HYBRID Riemann solver, and HLL)
_BW = {"BWpressure": 1.38e-11, "BWdensity": 2.34e-22, "BWmagfieldX": 0.0, "BWmagfieldY": 0.0, "BWmagfieldZ": 0.0, "BW_energy": 1.0e51,
       "BW_nzones": 2, "BW_blast_dens": 2.34e-22, "BW_interface": 1.0e50, "BW_amb2_RO": 0.0, "BW_amb2_PG": 0.0, "BW_amb2_VX": 0.0,
       "BW_amb2_VY": 0.0, "BW_amb2_VZ": 0.0, "InitIons": "LEAVE"}
_BWREF = (2.34e-22, 1.38e-11, 1e6, 1e6, 1e6) + (1.0,) * 11
MORE_TEST_PROBLEMS.update({
    "tp_BWaxi2D_halfplane_n016_roe": (Problem(ndim=2, NG=(32, 16, 1), eqn="euler", solver=4, artviscosity=1, etav=0.1, gamma=1.666666666666666666666,
                                              cfl=0.3, coords="cylindrical", xmin=(-30.86e18, 0.0, 0.0), xmax=(30.86e18, 30.86e18, 1.0),
                                              bcs=("outflow", "outflow", "reflecting", "outflow", "periodic", "periodic"), ics="BlastWave",
                                              extra=_BW, refvec=_BWREF), 20),
    "tp_BWsph1D_n128_hybrid": (Problem(ndim=1, NG=(128, 1, 1), eqn="euler", solver=3, artviscosity=1, etav=0.1, gamma=1.666666666666666666666, cfl=0.3,
                                       coords="spherical", xmin=(0.0, 0.0, 0.0), xmax=(30.86e18, 1.0, 1.0),
                                       bcs=("reflecting", "outflow") + ("periodic",) * 4, ics="BlastWave", extra=_BW, refvec=_BWREF), 20),
    "tp_BWsph1D_n128_hll": (Problem(ndim=1, NG=(128, 1, 1), eqn="euler", solver=8, artviscosity=1, etav=0.1, gamma=1.666666666666666666666, cfl=0.3,
                                    coords="spherical", xmin=(0.0, 0.0, 0.0), xmax=(30.86e18, 1.0, 1.0),
                                    bcs=("reflecting", "outflow") + ("periodic",) * 4, ics="BlastWave", extra=_BW, refvec=_BWREF), 20),
})
TEST_PROBLEMS.update(MORE_TEST_PROBLEMS)
TEST_PROBLEMS.update(SHOCK_TUBES)
CASES.update(TEST_PROBLEMS)

# Per-cell cooling source term (mp_only_cooling, EP_cooling 8) on cgs states; `dense` puts the
# cooling time near the CFL step (adaptive sub-stepping + MP_timestep_limit active).
COOLING = {
    "cool_euler_hll_3d_tracer": (case_cooling("euler", 8, ntracer=1), 4, 2.0e-24),
    "cool_euler_hll_3d_dense": (case_cooling("euler", 8, ntracer=1), 4, 2.0e-21),
    "cool_glm_hlld_3d_dense": (case_cooling("glm-mhd", 7, ntracer=0, bcs="periodic"), 3, 2.0e-22),
}



def _cool(flag, eqn, solver, Tmin, Tmax, ntracer):
    import dataclasses
    return dataclasses.replace(case_cooling(eqn, solver, ntracer=ntracer), cooling=flag, min_temperature=Tmin, max_temperature=Tmax)


# the other cooling functions of mp_only_cooling::Edot: (problem, steps, rho0, T range of the seeded state)
COOLING.update({
    "cool_ki02_euler_hll_3d": (_cool(2, "euler", 8, 10.0, 1.0e5, 1), 4, 2.0e-22, 20.0, 2.0e4),
    "cool_sd93_euler_roe_3d": (_cool(4, "euler", 4, 1.0e4, 1.0e8, 0), 4, 2.0e-23, 2.0e4, 5.0e7),
    "cool_sd93heat_glm_hlld_3d": (_cool(5, "glm-mhd", 7, 5.0e3, 1.0e8, 0), 3, 2.0e-22, 6.0e3, 5.0e7),
    "cool_wss09heat_euler_hll_3d": (_cool(6, "euler", 8, 5.0e3, 1.0e8, 1), 4, 2.0e-22, 6.0e3, 5.0e7),
    "cool_wss09_imhd_hll_3d": (_cool(7, "i-mhd", 8, 1.0e4, 1.0e8, 0), 3, 2.0e-23, 2.0e4, 5.0e7),
})
CASES.update({k: v[:2] for k, v in COOLING.items()})


def make_cooling_splines():
    """Knots of the reference's cooling-curve splines (SD93-CIE for EP_cooling 2..5, WSS09-CIE for 6, 7)."""
    out = {}
    for pre, flag in (("sd93", 4), ("wss09", 6)):
        r = RefSim(_cool(flag, "euler", 8, 5.0e3, 1.0e8, 0))
        sp = r.cooling_spline()
        r.close()
        out[pre + "_logT"], out[pre + "_logL"], out[pre + "_slopes"] = sp["spline_logT"], sp["spline_logL"], sp["spline_slopes"]
    np.savez_compressed(COOLING_SPLINES, **out)

# Stellar-wind internal boundary + cooling: the Wind3D configuration at 16^3 / 24^2
WIND = {
    "wind3d_euler_hll_cool_n016": (case_wind("euler", 8), 14),
    "wind3d_glm_hlld_nocool_n012": (case_wind("glm-mhd", 7, NG=(12, 12, 12), cooling=False, vrot=20.0), 6),
    "wind3d_euler_roe_nocool_n012": (case_wind("euler", 4, NG=(12, 12, 12), cooling=False), 6),
}
CASES.update(WIND)


def main():
    only = set(sys.argv[1:])  # optional: regenerate just these fixtures
    if not COOLING_SPLINES.exists() or "cooling_splines" in only:
        make_cooling_splines()
    for name, (prob, nsteps) in CASES.items():
        if only and name not in only:
            continue
        if name in HOT:
            P0 = hot_sphere_state(prob, seed=2024)
            r = RefSim(prob)
            r.set_state(P0)
        elif name in WIND:
            r = RefSim(prob)
            P0 = wind_ambient_state(prob)
            r.set_state(P0)
        elif name in COOLING:
            r = RefSim(prob)
            if prob.cooling == 8 and (not COOLING_TABLES.exists() or name == next(iter(COOLING))):
                tab = r.cooling_tables()
                np.savez_compressed(COOLING_TABLES, **tab)
            P0 = cooling_state(prob, seed=2024, rho0=COOLING[name][2], **(dict(Tlo=COOLING[name][3], Thi=COOLING[name][4]) if len(COOLING[name]) > 3 else {}))
            r.set_state(P0)
        elif name in TEST_PROBLEMS:
            r = RefSim(prob, run_ics=True)
            P0 = r.get_state(0)
        else:
            P0 = random_state(prob, seed=2024)
            r = RefSim(prob)
            r.set_state(P0)
        r.init_after_state()
        Pinit = r.get_state(0)
        dts = r.run(nsteps)
        P = r.get_state(0)
        r.close()
        np.savez_compressed(HERE / f"{name}.npz", P0=P0, Pinit=Pinit, dts=dts, P=P, nsteps=nsteps)
        print(name, P.shape, dts)


if __name__ == "__main__":
    main()
