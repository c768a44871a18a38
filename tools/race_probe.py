"""Developer script: a few steps of the TMA sweep kernel on small multi-tile 3-D grids (HLLD, tracers, cooling + wind),
a smoke driver for manual checks (compute-sanitizer is closed on this pool: gpurun refuses it):
  python tools/race_probe.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import dataclasses
import numpy as np
from harness import GpuSim, random_state, load_cooling_tables
from cases import case_3d, case_cooling

cases = [case_3d("glm-mhd", 7, 1, bcs="reflect-outflow", NG=(40, 26, 20)),
         case_3d("glm-mhd", 7, 1, bcs="mixed1", ntracer=1, NG=(40, 26, 20)),
         case_3d("euler", 8, 1, bcs="mixed1", ntracer=2, NG=(40, 26, 20)),
         case_3d("i-mhd", 4, 0, bcs="outflow", NG=(34, 13, 10))]
for prob in cases:
    g = GpuSim(prob)
    g.set_state(random_state(prob, 7))
    g.init_after_state()
    g.run(2)
    P = g.get_state(0)
    print(prob.eqn, prob.solver, prob.ntracer, g.ctx.describe()[:90], float(np.abs(P).max()), flush=True)
    g.close()
print("race_probe done")
