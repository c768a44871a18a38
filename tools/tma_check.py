"""Developer script: oracle vs GPU (TMA sweep) vs GPU (LDG sweep, PION_B200_NO_TMA=1) on 3-D multi-tile
cases; prints the per-variable error and where the worst cell sits."""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from harness import *
from cases import case_3d, EQ_SOLVERS

def run(prob, nsteps, no_tma):
    if no_tma: os.environ["PION_B200_NO_TMA"] = "1"
    else: os.environ.pop("PION_B200_NO_TMA", None)
    g = GpuSim(prob); P = random_state(prob, 7); g.set_state(P); g.init_after_state()
    g.run(nsteps); out = g.get_state(0); g.close(); return out

cases = [(e, s, av) for (e, s) in EQ_SOLVERS for av in (0, 1, 4)]
if len(sys.argv) > 1: cases = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))] * (int(sys.argv[4]) if len(sys.argv) > 4 else 1)
for eqn, solver, av in cases:
    prob = case_3d(eqn, solver, av, bcs="reflect-outflow", NG=(40, 26, 20))
    for nsteps in (1, 3):
        o = OracleSim(prob); P = random_state(prob, 7); o.set_state(P); o.init_after_state(); o.run(nsteps); Po = o.get_state(0); o.close()
        res = {}
        for name, no_tma in (("tma", False), ("ldg", True)):
            Pg = run(prob, nsteps, no_tma)
            e = rel_err(Pg, Po, nphys=prob.nvar - prob.ntracer)
            v = int(np.argmax(e)); w = np.unravel_index(np.argmax(np.abs(Pg[v] - Po[v])), Pg[v].shape)
            nbad = int(np.sum(np.abs(Pg[v] - Po[v]) > 1e-9 * np.max(np.abs(Po[v]))))
            res[name] = "%s err=%.2e var=%d at(k,j,i)=%s nbad=%d" % (name, e.max(), v, tuple(int(x) for x in w), nbad)
        print(eqn, solver, av, "steps", nsteps, "|", res["tma"], "|", res["ldg"], flush=True)
