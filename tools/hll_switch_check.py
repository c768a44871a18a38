"""Developer script: parity with the HLLD->HLL switch active (hot-sphere state): oracle vs TMA sweep / LDG sweep /
gather kernel, 3-D and 2-D."""
import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from harness import *
from cases import case_2d, case_3d

def gpu(prob, P, nsteps, env):
    for k in ("PION_B200_NO_TMA", "PION_B200_GATHER"): os.environ.pop(k, None)
    os.environ.update(env)
    g = GpuSim(prob); g.set_state(P); g.init_after_state(); g.run(nsteps); out = g.get_state(0); g.close(); return out

for eqn in ("i-mhd", "glm-mhd"):
    for av in (0, 1):
        for prob, nm in ((case_3d(eqn, 7, av, bcs="reflect-outflow", NG=(40, 26, 20)), "3d"), (case_2d(eqn, 7, av, bcs="outflow", NG=(48, 40, 1)), "2d")):
            P = hot_sphere_state(prob)
            for nsteps in (1, 2):
                o = OracleSim(prob); o.set_state(P); o.init_after_state(); o.run(nsteps); Po = o.get_state(0); o.close()
                res = []
                for name, env in (("tma", {}), ("ldg", {"PION_B200_NO_TMA": "1"}), ("gather", {"PION_B200_GATHER": "1"})):
                    Pg = gpu(prob, P, nsteps, env)
                    e = rel_err(Pg, Po, nphys=prob.nvar)
                    v = int(np.argmax(e)); nbad = int(np.sum(np.abs(Pg[v] - Po[v]) > 1e-9 * np.max(np.abs(Po[v]))))
                    res.append("%s %.1e(v%d,n%d)" % (name, e.max(), v, nbad))
                print(eqn, av, nm, "steps", nsteps, " | ".join(res), flush=True)
