"""The state bench.py times (DTE3D_MHD-style: static medium, exact-zero velocities, uniform B_x, x200
pressure sphere) checked against the reference's CPU implementation, plus the run-length items the round-1
review asked for: the output-time limiter across output intervals and error bounds after 100 steps of the
reference's own shock problems."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from bench import divb_norm, dte_problem, dte_state  # noqa: E402  (the bench's own generators: the SAME state)
from cases import case_2d, case_3d  # noqa: E402
from harness import GpuSim, OracleSim, RefSim, have_ref, random_state, rel_err, ulp_response  # noqa: E402
from test_golden import load  # noqa: E402

L = 3.086e19
TOL = 5e-12


def dte_case(n):
    prob = dte_problem((n,) * 3, (-L,) * 3, (L,) * 3)
    return prob, dte_state(prob, (-L,) * 3, (L,) * 3)


# ---------------------------------------------------------------- CPU: oracle vs the compiled reference
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_oracle_reproduces_reference_on_the_bench_state():
    prob, P0 = dte_case(24)
    r, o = RefSim(prob), OracleSim(prob)
    for s in (r, o):
        s.set_state(P0)
        s.init_after_state()
    assert np.array_equal(r.run(4), o.run(4))
    assert np.array_equal(r.get_state(0), o.get_state(0))
    # the sphere's surface trips the HLLD -> HLL switch: the state really exercises both solvers
    r.close()
    o.close()


def _two_interval_problem():
    import dataclasses
    prob = case_2d("euler", 8, 1, bcs="outflow", NG=(16, 12, 1))
    return dataclasses.replace(prob, op_criterion=1, opfreq_time=0.05, finishtime=0.3)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_output_time_limiter_runs_across_output_intervals_cpu():
    """OutputCriterion 1: dt is clipped to land on next_optime, which output_data then advances
    (sim_init.cpp:733-742).  Twelve steps cross two output times; without the bookkeeping the run aborts with
    'Went past output time without outputting!' at the first one."""
    prob = _two_interval_problem()
    P0 = random_state(prob, 5)
    r, o = RefSim(prob), OracleSim(prob)
    for s in (r, o):
        s.set_state(P0)
        s.init_after_state()
    dr, do = r.run(12), o.run(12)
    assert np.array_equal(dr, do)
    t = np.cumsum(dr)
    assert any(abs(t - 0.05) < 1e-15) and any(abs(t - 0.10) < 1e-15), t  # landed exactly on two output times
    assert np.array_equal(r.get_state(0), o.get_state(0))
    r.close()
    o.close()


# ---------------------------------------------------------------- GPU
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_bench_state_conditioning_is_what_design_md_says():
    """The state bench.py times is a static, symmetric medium with a x200 pressure discontinuity: the reference's
    OWN answer moves by ~1e-11..1e-10 (relative, transverse velocity / field) after a handful of steps when its
    input moves by one ulp -- discrete decisions (HLLD fan region, HLLD -> HLL switch, minmod) acting on
    rounding noise.  This is why the GPU tests on this state are stated against that response, not against 5e-12."""
    prob, P0 = dte_case(32)
    resp, _ = ulp_response(prob, P0, 5, seeds=(1,), sim=RefSim)
    assert resp[:3].max() < 1e-13  # density, pressure, v_x: well conditioned
    assert 1e-13 < resp[3:5].max() < 1e-8  # transverse velocities: rounding noise amplified by ~1e5


@pytest.mark.gpu
@pytest.mark.parametrize("n,nsteps", [(64, 5), (64, 6), (96, 5)])
def test_gpu_matches_reference_on_the_bench_state(n, nsteps):
    """64^3 = 2 x 6 tiles x 1 chunk, 96^3 = 3 x 9 tiles x 3 chunks of the TMA sweep kernel; >= 5 steps.
    Bound per variable: 5e-12 + 5 x (the reference's own response to a one-ulp change of its input, see
    test_bench_state_conditioning_is_what_design_md_says); dt sequences agree to 1e-13."""
    prob, P0 = dte_case(n)
    resp, Pr = ulp_response(prob, P0, nsteps)  # oracle: bit-exact with the compiled reference on this state
    if have_ref():
        ref = RefSim(prob)
        ref.set_state(P0)
        ref.init_after_state()
        dr = ref.run(nsteps)
        assert np.array_equal(ref.get_state(0), Pr)
        ref.close()
    else:
        o = OracleSim(prob)
        o.set_state(P0)
        o.init_after_state()
        dr = o.run(nsteps)
        o.close()
    gpu = GpuSim(prob)
    try:
        gpu.set_state(P0)
        gpu.init_after_state()
        dg = gpu.run(nsteps)
        assert np.allclose(dr, dg, rtol=1e-13, atol=0), (dr, dg)
        Pg = gpu.get_state(0)
        assert np.max(np.abs(Pr[2:5])) > 0.0  # the sphere has started to expand
        err = rel_err(Pg, Pr, nphys=9)
        print(f"DTE {n}^3 x {nsteps} steps: GPU-vs-reference {err}, reference 1-ulp response {resp}")
        assert np.all(err <= TOL + 5.0 * resp), (err, resp)
        assert gpu.error_counts() == [0, 0]
        assert "k_stage_sweep_tma<EQ=3,SOLVER=7,FKJ=1" in gpu.ctx.describe(), gpu.ctx.describe()
        # div B (reported by bench.py): the GPU run's is the reference run's
        assert abs(divb_norm(Pg, prob) - divb_norm(Pr, prob)) < 1e-9
    finally:
        gpu.close()


@pytest.mark.gpu
def test_gpu_output_time_limiter_runs_across_output_intervals():
    prob = _two_interval_problem()
    P0 = random_state(prob, 5)
    o, g = OracleSim(prob), GpuSim(prob)
    for s in (o, g):
        s.set_state(P0)
        s.init_after_state()
    do, dg = o.run(12), g.run(12)
    assert np.allclose(do, dg, rtol=1e-12, atol=0), (do, dg)
    assert rel_err(g.get_state(0), o.get_state(0), nphys=5).max() < TOL
    o.close()
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,bound", [("tp_DMR_n065_hll", 1e-9), ("tp_DMR_n065_roe", 1e-9), ("tp_BWcrt3D_octant_n016", 1e-9),
                                        ("tp_FieldLoop_64x32_hlld", 1e-9)])
def test_error_bound_after_100_steps_of_the_reference_test_problems(name, bound):
    """north_star: 'about 1e-12 per step and a documented bound after N steps'.  The reference's own shock
    problems (initial conditions from its IC classes: the golden fixtures' P0), 100 steps, GPU against the
    oracle (bit-exact with the compiled reference on these problems, test_golden.py): bound 1e-9, measured
    values are printed (pytest -s) and quoted in DESIGN.md."""
    prob, _, z = load(name)
    o, g = OracleSim(prob), GpuSim(prob)
    for s in (o, g):
        s.set_state(z["P0"])
        s.init_after_state()
    do, dg = o.run(100), g.run(100)
    e_dt = float(np.max(np.abs(do - dg) / do))
    err = rel_err(g.get_state(0), o.get_state(0), nphys=prob.nvar - prob.ntracer)
    print(f"{name}: after 100 steps max rel err per variable {err}, dt {e_dt:.2e}")
    assert err.max() < bound, err
    assert e_dt < bound
    o.close()
    g.close()
