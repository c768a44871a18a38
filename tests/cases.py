"""Shared parity-case definitions (equation set x solver x viscosity x boundaries)."""
from harness import Problem

EQ_SOLVERS = [("euler", 8), ("euler", 4), ("euler", 5), ("euler", 6), ("euler", 0), ("euler", 1), ("euler", 2), ("euler", 3), ("i-mhd", 8), ("i-mhd", 7), ("i-mhd", 4), ("i-mhd", 0), ("i-mhd", 1),
              ("glm-mhd", 8), ("glm-mhd", 7), ("glm-mhd", 4), ("glm-mhd", 0), ("glm-mhd", 1)]
AVS = [0, 1, 3, 4]
DX = 0.0625  # cells must be cubic (uniform_grid.cpp:866-874); a power of two keeps xmax exact

BCSETS = {
    "periodic": ("periodic",) * 6,
    "outflow": ("outflow",) * 6,
    "reflect-outflow": ("reflecting", "outflow") * 3,
    "mixed1": ("fixed", "one-way-outflow", "reflecting", "inflow", "periodic", "periodic"),
    "mixed2": ("one-way-outflow", "fixed", "one-way-outflow", "one-way-outflow", "inflow", "inflow"),
}


def case_2d(eqn, solver, av, bcs="periodic", ntracer=0, NG=(24, 16, 1)):
    return Problem(ndim=2, NG=NG, eqn=eqn, solver=solver, artviscosity=av, xmax=(NG[0] * DX, NG[1] * DX, 1.0),
                   bcs=BCSETS[bcs], ntracer=ntracer)


def case_3d(eqn, solver, av, bcs="periodic", ntracer=0, NG=(12, 10, 8), ooa=2):
    return Problem(ndim=3, NG=NG, eqn=eqn, solver=solver, artviscosity=av, xmax=(NG[0] * DX, NG[1] * DX, NG[2] * DX),
                   bcs=BCSETS[bcs], ntracer=ntracer, ooa=ooa)


def case_1d(eqn, solver, av, bcs=("outflow", "outflow")):
    return Problem(ndim=1, NG=(64, 1, 1), eqn=eqn, solver=solver, artviscosity=av, bcs=tuple(bcs) + ("periodic",) * 4)


def case_cooling(eqn="euler", solver=8, ndim=3, NG=(12, 10, 8), ntracer=1, mp_limit=1, bcs="reflect-outflow", ooa=2):
    """Wind3D-style units: cgs, dx = 1e17 cm, EP_cooling 8, T in [5e3, 1e8] K."""
    dx = 1.0e17
    return Problem(ndim=ndim, NG=NG, eqn=eqn, solver=solver, artviscosity=1, etav=0.1, cfl=0.3,
                   xmax=(NG[0] * dx, NG[1] * dx, NG[2] * dx if ndim > 2 else 1.0), bcs=BCSETS[bcs], ntracer=ntracer, ooa=ooa,
                   cooling=8, mp_timestep_limit=mp_limit, min_temperature=5.0e3, max_temperature=1.0e8,
                   refvec=(2.0e-24, 3.0e-10, 1.0e6, 1.0e6, 1.0e6, 1.0e-6, 1.0e-6, 1.0e-6, 1.0e-6) + (1.0,) * 7)


def case_wind(eqn="euler", solver=8, ndim=3, NG=(16, 16, 16), cooling=True, vrot=0.0):
    """Wind3D-style octant (SURVEY 8d config 5 at test size): constant stellar wind at the origin corner,
    reflecting N faces / one-way-outflow P faces, 1 tracer, EP_cooling 8 with MP_timestep_limit 1."""
    import dataclasses
    base = case_cooling(eqn, solver, ndim=ndim, NG=NG, ntracer=1, mp_limit=1)
    wind = dict(pos=(0.0, 0.0, 0.0), radius=4.3 * base.dx, mdot=1.0e-7, vinf=1500.0, vrot=vrot, temp=3.0e4, rstar=6.96e11,
                bsrf=10.0, tr=(1.0, 0.0, 0.0, 0.0))
    kw = dict(internal_bcs=("stellar-wind",), winds=(wind,), bcs=("reflecting", "one-way-outflow") * 3)
    if not cooling:
        kw.update(cooling=0, mp_timestep_limit=0, min_temperature=0.0, max_temperature=1.0e99)
    return dataclasses.replace(base, **kw)


def wind_ambient_state(prob):
    """Uniform Wind3D ambient medium (params_Wind3D_n0128_l2.txt: rho 2.124e-24, p 2.209e-12, tracer 0)."""
    import numpy as np
    P = np.zeros(prob.padded_shape())
    P[0] = 2.124e-24
    P[1] = 2.209e-12
    if prob.eqn != "euler":
        P[5] = 1.0e-6
    return P


def case_cyl(eqn, solver, av, NG=(24, 16, 1), bcs=("outflow", "outflow", "reflecting", "outflow"), ntracer=0, ooa=2):
    """2-D axisymmetric (z, R) grid, R from 0 with a reflecting axis (as the reference's cylindrical test problems)."""
    return Problem(ndim=2, NG=NG, eqn=eqn, solver=solver, artviscosity=av, xmax=(NG[0] * DX, NG[1] * DX, 1.0),
                   bcs=tuple(bcs) + ("periodic",) * 2, ntracer=ntracer, ooa=ooa, coords="cylindrical")


def case_sph(solver, av, N=64, rmin=0.0, bcs=("reflecting", "outflow"), ntracer=0, ooa=2):
    """1-D spherically symmetric Euler grid."""
    return Problem(ndim=1, NG=(N, 1, 1), eqn="euler", solver=solver, artviscosity=av, xmin=(rmin, 0.0, 0.0),
                   xmax=(rmin + N * DX / 4, 1.0, 1.0), bcs=tuple(bcs) + ("periodic",) * 4, ntracer=ntracer, ooa=ooa,
                   coords="spherical")
