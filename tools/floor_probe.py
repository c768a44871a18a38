"""Developer script: GPU vs oracle in the regime where UtoP resets negative pressures (cold, highly supersonic random flow,
p ~ 1e-7 rho v^2): fix-up counters, dt and state differences.  See DESIGN.md section 6, "open item"."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
from harness import GpuSim, OracleSim, random_state, rel_err
from cases import case_2d, case_3d
def cold(p):
    P = random_state(p, 3, amp=3.0); P[1] *= 1e-7; return P
for prob in (case_2d("euler", 8, 1, bcs="outflow"), case_3d("i-mhd", 7, 1, bcs="outflow", NG=(40, 26, 20)), case_3d("euler", 4, 1, bcs="mixed1", ntracer=1, NG=(40, 26, 20))):
    o, g = OracleSim(prob), GpuSim(prob)
    for s in (o, g):
        s.set_state(cold(prob)); s.init_after_state()
    do, dg = o.run(3), g.run(3)
    print(prob.eqn, prob.solver, prob.ndim, "counts", o.error_counts(), g.error_counts(), "dt", float(np.max(np.abs(do - dg) / do)), "err", rel_err(g.get_state(0), o.get_state(0), nphys=prob.nvar - prob.ntracer).max())
    o.close(); g.close()
