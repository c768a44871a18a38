#!/usr/bin/env python
"""Developer script: error of the CUDA path against the oracle over LONG runs of the reference's shock test problems
(BASELINE configs 1 and 2 at their named sizes): double Mach reflection 260x80 to t = 0.2 (Roe-CV and HLL) and the
advected field loop 128x64 GLM-MHD HLLD for 400 steps; plus a 3-D GLM-MHD blast (the bench state, 48^3, 100 steps).
Prints one JSON line per case: steps, max relative error per variable (scale = max |variable|), L1 relative error, divB.
The bounds asserted in tests/test_gpu_parity.py::test_shock_problem_error_bound come from this output."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
import bench_configs as bc  # noqa: E402
from harness import GpuSim, OracleSim, Problem, rel_err  # noqa: E402


def steps_to(sim, tfinal, chunk=50, nmax=5000):
    n = 0
    while n < nmax:
        t = sim.get_time()[0] if hasattr(sim, "get_time") else None
        if t is not None and t >= tfinal * (1 - 1e-14):
            break
        sim.run(chunk)
        n += chunk
    return n


def l1(a, b):
    return [float(np.sum(np.abs(a[v] - b[v])) / max(np.sum(np.abs(b[v])), 1e-300)) for v in range(a.shape[0])]


def case(name, prob, P0, nsteps):
    o, g = OracleSim(prob), GpuSim(prob)
    for s in (o, g):
        s.set_state(P0)
        s.init_after_state()
    out = []
    done = 0
    for n in nsteps:
        do, dg = o.run(n - done), g.run(n - done)
        done = n
        Po, Pg = o.get_state(0), g.get_state(0)
        nb = prob.nbc
        sl = (slice(None), slice(nb, -nb) if prob.ndim == 3 else slice(None), slice(nb, -nb), slice(nb, -nb))
        e = rel_err(Pg, Po)
        out.append({"case": name, "grid": list(prob.NG), "steps": n, "time": float(np.sum(do)) if done == n else None,
                    "max_rel_err_per_variable": [float(x) for x in e], "max_rel_err": float(e.max()),
                    "l1_rel_err_per_variable": l1(Pg[sl], Po[sl]), "dt_max_rel_err": float(np.max(np.abs(do - dg) / do)),
                    "negative_density": g.error_counts()[0]})
        print(json.dumps(out[-1]), flush=True)
    o.close(); g.close()
    return out


def main():
    cfgs = {c[0]: c for c in bc.configs(False)}
    import dataclasses
    # DMR 260x80 (config 1): ~600 steps to t = 0.2
    for key in ("1 DMR 2-D Euler 260x80 solver 4",):
        _, prob, ic, _, _ = cfgs[key]
        for solver in (4, 8):
            p = dataclasses.replace(prob, solver=solver, finishtime=1e30)
            case(f"DMR 260x80 solver {solver}", p, ic(p), [12, 100, 300, 600])
    _, prob, ic, _, _ = cfgs["2 FieldLoop 2-D GLM-MHD 512x256 solver 7"]
    p = dataclasses.replace(prob, NG=(128, 64, 1), finishtime=1e30)
    case("FieldLoop 128x64 GLM-MHD HLLD", p, ic(p), [12, 100, 400])
    _, prob, ic, _, _ = cfgs["4 DTE3D-style 3-D GLM-MHD 512^3 HLLD"]
    p = dataclasses.replace(prob, NG=(48, 48, 48))
    case("DTE3D-style 48^3 GLM-MHD HLLD", p, ic(p), [5, 30, 100])


if __name__ == "__main__":
    main()
