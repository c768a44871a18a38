import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, os
if os.environ.get("PION_DBG_LIB"):
    import pion_b200.capi as capi; capi.load_library(os.environ["PION_DBG_LIB"])
from cases import case_2d, case_3d
from harness import GpuSim, OracleSim, random_state
prob = case_2d("euler", 8, 0)
o, g = OracleSim(prob), GpuSim(prob)
P = random_state(prob, 7)
for s in (o, g):
    s.set_state(P); s.init_after_state()
# ghost update alone: scramble the ghost cells of Ph and P on both sides, then TimeUpdate BCs
Q = o.get_state(0).copy()
rng = np.random.default_rng(3)
mask = np.ones(Q.shape, bool); mask[prob.interior()] = False
Q[mask] = rng.random(Q.shape)[mask]
for s in (o, g):
    s.set_state(Q)
    s.update_bcs(2, 2)
for which in (0, 1):
    d = np.abs(o.get_state(which) - g.get_state(which))
    print("array", which, "max diff after update_bcs alone:", d.max())
    for v in range(1):
        print("\n".join("".join("X" if x > 1e-14 else "." for x in row) for row in d[v, 0]))
o.close(); g.close()
