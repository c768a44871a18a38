#!/usr/bin/env python
"""Developer script: bench-state parity (DTE 64^3, 5 steps) of one library build against the CPU checker,
per-variable errors and dt errors printed.  usage: gpu_diag.py [lib.so]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
if len(sys.argv) > 1:
    import pion_b200.capi as capi
    capi.load_library(sys.argv[1])
from bench import dte_problem, dte_state
from harness import GpuSim, OracleSim, rel_err
L = 3.086e19
for n, steps in ((64, 5), (96, 3)):
    prob = dte_problem((n,) * 3, (-L,) * 3, (L,) * 3)
    P0 = dte_state(prob, (-L,) * 3, (L,) * 3)
    o = OracleSim(prob); o.set_state(P0); o.init_after_state(); do = o.run(steps); Po = o.get_state(0); o.close()
    g = GpuSim(prob); g.set_state(P0); g.init_after_state(); dg = g.run(steps); Pg = g.get_state(0); g.close()
    print(n, "dt rel err", np.abs(do - dg) / do)
    print(n, "state err ", rel_err(Pg, Po, nphys=9))
