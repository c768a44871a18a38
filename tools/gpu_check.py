"""Developer script: quick parity sweep GPU vs oracle (and vs the compiled
reference when present) over equation/solver/viscosity/BC combinations."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from harness import *

def cmp(prob, nsteps=3, seed=7):
    o = OracleSim(prob); g = GpuSim(prob)
    P = random_state(prob, seed)
    for s in (o, g):
        s.set_state(P); s.init_after_state()
    i0 = np.max(np.abs(o.get_state(0) - g.get_state(0)))
    do = o.run(nsteps); dg = g.run(nsteps)
    Po, Pg = o.get_state(0), g.get_state(0)
    e = rel_err(Pg, Po)
    ei = rel_err(Pg[prob.interior()], Po[prob.interior()])
    print(f"{prob.eqn:8s} nd={prob.ndim} s={prob.solver} av={prob.artviscosity} ntr={prob.ntracer} bc={prob.bcs[:2*prob.ndim]} "
          f"init={i0:.1e} dt={np.max(np.abs(do-dg)/do):.1e} err={e.max():.2e} int={ei.max():.2e} cnt={g.error_counts()}", flush=True)
    o.close(); g.close()
    return e.max()

worst = 0
for eqn, solvers in (("euler", (8, 4)), ("i-mhd", (8, 7, 4)), ("glm-mhd", (8, 7, 4))):
    for solver in solvers:
        for av in (0, 1, 3, 4):
            worst = max(worst, cmp(Problem(ndim=2, NG=(24, 16, 1), eqn=eqn, solver=solver, artviscosity=av, xmax=(1.5, 1.0, 1.0))))
            worst = max(worst, cmp(Problem(ndim=3, NG=(12, 10, 8), eqn=eqn, solver=solver, artviscosity=av, xmax=(1.2, 1.0, 0.8), ntracer=1)))
bcsets = [("outflow",) * 6, ("reflecting", "outflow") * 3, ("fixed", "one-way-outflow", "reflecting", "inflow", "periodic", "periodic"),
          ("one-way-outflow", "fixed", "one-way-outflow", "one-way-outflow", "inflow", "inflow")]
for b in bcsets:
    for eqn, sv in (("glm-mhd", 7), ("euler", 8), ("i-mhd", 4)):
        worst = max(worst, cmp(Problem(ndim=3, NG=(10, 8, 6), eqn=eqn, solver=sv, artviscosity=1, xmax=(1.0, 0.8, 0.6), bcs=b)))
        worst = max(worst, cmp(Problem(ndim=2, NG=(10, 8, 1), eqn=eqn, solver=sv, artviscosity=4, xmax=(1.0, 0.8, 0.6), bcs=b, ntracer=1)))
worst = max(worst, cmp(Problem(ndim=1, NG=(64, 1, 1), eqn="i-mhd", solver=7, artviscosity=1, bcs=("outflow", "outflow") + ("periodic",) * 4)))
worst = max(worst, cmp(Problem(ndim=1, NG=(64, 1, 1), eqn="euler", solver=8, artviscosity=1, bcs=("reflecting", "inflow") + ("periodic",) * 4)))
worst = max(worst, cmp(Problem(ndim=3, NG=(12, 10, 8), eqn="glm-mhd", solver=7, artviscosity=1, ooa=1, xmax=(1.2, 1.0, 0.8))))
print("WORST", worst)
