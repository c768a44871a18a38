"""Developer script: where do the TMA-sweep results differ from the oracle after ONE step (interior indices)."""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from harness import *
from cases import case_3d
eqn, solver, av = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
prob = case_3d(eqn, solver, av, bcs="reflect-outflow", NG=(40, 26, 20))
o = OracleSim(prob); P = random_state(prob, 7); o.set_state(P); o.init_after_state(); o.run(1); Po = o.get_state(0)[prob.interior()]; o.close()
g = GpuSim(prob); g.set_state(P); g.init_after_state(); g.run(1); Pg = g.get_state(0)[prob.interior()]; g.close()
for v in range(Po.shape[0]):
    d = np.abs(Pg[v] - Po[v]); bad = np.argwhere(d > 1e-10 * np.max(np.abs(Po[v])))
    if len(bad) == 0: print("var", v, "ok"); continue
    print("var", v, "nbad", len(bad), "k", sorted(set(bad[:, 0]))[:30], "j", sorted(set(bad[:, 1]))[:30], "i", sorted(set(bad[:, 2]))[:45], "max", d.max())
