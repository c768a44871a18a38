"""Developer script: GPU vs one golden fixture, printing the dt and state differences.  usage: golden_diff.py <name>"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
from harness import GpuSim, rel_err, tables_for
from test_golden import load
name = sys.argv[1]
prob, nsteps, z = load(name)
g = GpuSim(prob, tables=tables_for(prob))
g.set_state(z["P0"]); g.init_after_state()
dts = g.run(nsteps)
P = g.get_state(0)
print("dt rel diff per step", np.abs(dts - z["dts"]) / z["dts"])
print("state rel err per variable", rel_err(P, z["P"], nphys=prob.nvar - prob.ntracer))
d = np.abs(P - z["P"]); v = int(np.argmax(d.max(axis=(1, 2, 3)))); i = int(np.argmax(d[v, 0, 0]))
print("worst variable", v, "cell", i, P[v, 0, 0, i - 2:i + 3], z["P"][v, 0, 0, i - 2:i + 3], "riemann failures", g.ctx.riemann_failures())
g.close()
