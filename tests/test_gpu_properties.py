"""GPU: size-independent properties at BASELINE.json's full grid sizes (the oracle takes minutes
there, so parity at these sizes is checked through invariants the scheme guarantees):

* conservation: with periodic boundaries the finite-volume update conserves the sums of mass,
  momentum and total energy (Euler) / mass and the magnetic flux components (GLM-MHD: the Powell
  and GLM source terms are not conservative in momentum and energy) to rounding;
* symmetry: a state invariant under the reflection (x,y,vx,vy,Bx,By) <-> (y,x,vy,vx,By,Bx) stays
  invariant (every axis goes through the same code with the frame rotated);
* decoupling: a 2-D problem extruded along z evolves exactly like the 2-D problem in every plane;
* FieldLoop (config 2): max |div B| dx / |B| stays small and the magnetic energy decays slowly."""
import numpy as np
import pytest

from cases import DX
from harness import GpuSim, Problem, random_state

pytestmark = pytest.mark.gpu


def conserved_sums(P, prob):
    g = prob.gamma
    I = P[prob.interior()].astype(np.longdouble)
    rho, p, v = I[0], I[1], I[2:5]
    out = {"mass": rho.sum(), "mx": (rho * v[0]).sum(), "my": (rho * v[1]).sum(), "mz": (rho * v[2]).sum()}
    e = 0.5 * rho * (v ** 2).sum(0) + p / (g - 1)
    if prob.eqn != "euler":
        B = I[5:8]
        e = e + 0.5 * (B ** 2).sum(0)
        out.update(bx=B[0].sum(), by=B[1].sum(), bz=B[2].sum())
        if prob.eqn == "glm-mhd":
            e = e + 0.5 * I[8] ** 2
    out["energy"] = e.sum()
    return out


def test_euler_256cubed_conserves_mass_momentum_energy():
    N = 256  # config 3 size (blast wave 256^3 Euler)
    prob = Problem(ndim=3, NG=(N, N, N), eqn="euler", solver=4, artviscosity=1, etav=0.1, xmax=(N * DX,) * 3)
    g = GpuSim(prob)
    g.set_state(random_state(prob, seed=5))
    g.init_after_state()
    c0 = conserved_sums(g.get_state(0), prob)
    g.run(5)
    c1 = conserved_sums(g.get_state(0), prob)
    assert g.error_counts() == [0, 0]
    g.close()
    scale_m = float(c0["mass"]) * 0.5  # |rho v| ~ 0.5 rho
    assert abs(c1["mass"] - c0["mass"]) <= 1e-12 * abs(c0["mass"])
    assert abs(c1["energy"] - c0["energy"]) <= 1e-12 * abs(c0["energy"])
    for k in ("mx", "my", "mz"):
        assert abs(c1[k] - c0[k]) <= 1e-12 * scale_m, (k, c0[k], c1[k])


def test_glm_mhd_256cubed_conserves_mass_and_flux():
    N = 256  # half the linear size of config 4 (512^3 needs 40 GB of host arrays for the sums)
    prob = Problem(ndim=3, NG=(N, N, N), eqn="glm-mhd", solver=7, artviscosity=1, etav=0.15, cfl=0.2, xmax=(N * DX,) * 3)
    g = GpuSim(prob)
    g.set_state(random_state(prob, seed=6))
    g.init_after_state()
    c0 = conserved_sums(g.get_state(0), prob)
    g.run(4)
    c1 = conserved_sums(g.get_state(0), prob)
    assert g.error_counts() == [0, 0]
    g.close()
    assert abs(c1["mass"] - c0["mass"]) <= 1e-12 * abs(c0["mass"])
    bscale = float(c0["mass"]) * 0.5
    # d/dt sum(B) = -sum(div-B source) is not zero with Powell terms, but d/dt sum(B_n) from the GLM
    # flux c_h psi telescopes: check the flux-form part through the total being finite and small drift
    for k in ("bx", "by", "bz"):
        assert np.isfinite(float(c1[k])) and abs(c1[k] - c0[k]) <= 1e-2 * bscale


def test_xy_reflection_symmetry_512x512():
    N = 512
    prob = Problem(ndim=2, NG=(N, N, 1), eqn="glm-mhd", solver=7, artviscosity=1, xmax=(N * DX, N * DX, 1.0))
    P = random_state(prob, seed=8)
    T = lambda A: np.swapaxes(A, -1, -2)
    # symmetrise under (x,y) <-> (y,x): scalars symmetric, (vx,vy) and (Bx,By) swap
    S = P.copy()
    for v in (0, 1, 4, 7, 8):
        S[v] = 0.5 * (P[v] + T(P[v]))
    S[2], S[3] = 0.5 * (P[2] + T(P[3])), 0.5 * (P[3] + T(P[2]))
    S[5], S[6] = 0.5 * (P[5] + T(P[6])), 0.5 * (P[6] + T(P[5]))
    S[3] = T(S[2])
    S[6] = T(S[5])
    g = GpuSim(prob)
    g.set_state(S)
    g.init_after_state()
    g.run(6)
    Q = g.get_state(0)
    g.close()
    # v_z and B_z change sign under the reflection (pseudo-vector components of a 2-D MHD state do not;
    # here they are polar components of the mirrored frame): compare magnitudes pattern-wise
    for a, b in ((0, 0), (1, 1), (2, 3), (5, 6), (8, 8)):
        err = np.max(np.abs(Q[a] - T(Q[b]))) / max(np.max(np.abs(Q[a])), 1e-300)
        assert err < 1e-11, (a, b, err)


def test_extruded_2d_problem_matches_2d_run():
    n = 96
    p2 = Problem(ndim=2, NG=(n, n, 1), eqn="glm-mhd", solver=7, artviscosity=1, xmax=(n * DX, n * DX, 1.0))
    p3 = Problem(ndim=3, NG=(n, n, 64), eqn="glm-mhd", solver=7, artviscosity=1, xmax=(n * DX, n * DX, 64 * DX))
    P2 = random_state(p2, seed=9)
    P2[4] = 0.0  # no z velocity / field: the extruded problem stays z-independent
    P2[7] = 0.0
    P3 = np.repeat(P2, p3.padded_shape()[1], axis=1)
    g2, g3 = GpuSim(p2), GpuSim(p3)
    for g, P in ((g2, P2), (g3, P3)):
        g.set_state(P)
        g.init_after_state()
    d2, d3 = g2.run(5), g3.run(5)
    Q2, Q3 = g2.get_state(0), g3.get_state(0)
    g2.close()
    g3.close()
    assert np.allclose(d2, d3, rtol=1e-13, atol=0)
    for k in (2, 20, 40, 65):
        for v in range(9):
            err = np.max(np.abs(Q3[v, k] - Q2[v, 0])) / max(np.max(np.abs(Q2[v])), 1e-300)
            assert err < 1e-11 or np.max(np.abs(Q2[v])) < 1e-12, (k, v, err)


def test_fieldloop_512x256_divB_and_magnetic_energy():
    """Config 2: FieldLoop 512x256 GLM-MHD (params_FieldLoop200.txt scaled), ICs restated from
    ics/basic_tests.cpp:577-666: A_z = 1e-3 max(0, 0.3 - r), B = curl A by central differences."""
    nx, ny = 512, 256
    prob = Problem(ndim=2, NG=(nx, ny, 1), eqn="glm-mhd", solver=7, artviscosity=1, etav=0.1, gamma=5.0 / 3.0, cfl=0.4,
                   xmin=(-1.0, -0.5, 0.0), xmax=(1.0, 0.5, 1.0), finishtime=2.0)
    shp = prob.padded_shape()
    g_, dx = prob.nbc, prob.dx
    x = prob.xmin[0] + (np.arange(shp[3]) - g_ + 0.5) * dx
    y = prob.xmin[1] + (np.arange(shp[2]) - g_ + 0.5) * dx
    X, Y = np.meshgrid(x, y)
    Az = lambda xx, yy: 1.0e-3 * np.maximum(0.0, 0.3 - np.sqrt(xx * xx + yy * yy))
    P = np.zeros(shp)
    P[0], P[1] = 1.0, 1.0
    P[2], P[3] = 2.0, 1.0
    P[5, 0] = (Az(X, Y + dx) - Az(X, Y - dx)) / (2 * dx)
    P[6, 0] = -(Az(X + dx, Y) - Az(X - dx, Y)) / (2 * dx)
    g = GpuSim(prob)
    g.set_state(P)
    g.init_after_state()

    def diag(Q):
        I = Q[prob.interior()]
        bx, by = Q[5, 0], Q[6, 0]
        div = ((bx[2:-2, 3:-1] - bx[2:-2, 1:-3]) + (by[3:-1, 2:-2] - by[1:-3, 2:-2])) / (2 * dx)
        bmax = np.max(np.sqrt(I[5] ** 2 + I[6] ** 2))
        return np.max(np.abs(div)) * dx / bmax, float(np.sum(I[5] ** 2 + I[6] ** 2 + I[7] ** 2))

    d0, e0 = diag(g.get_state(0))
    g.run(200)
    d1, e1 = diag(g.get_state(0))
    assert g.error_counts() == [0, 0]
    g.close()
    assert d0 < 1e-10          # curl of a potential by the same central differences: div-free to rounding
    assert d1 < 0.1            # GLM cleaning keeps the normalised divergence bounded
    assert 0.7 * e0 < e1 <= e0 * (1 + 1e-9)  # magnetic energy decays slowly, never grows
