#!/usr/bin/env python
"""Throughput of the five BASELINE.json configurations on one B200 next to the reference's CPU
implementation (oracle/_ref, one host core, reduced grid).  Prints a markdown table and writes
profiles/r01_configs.json.  Developer/bench script -- TEST INFRASTRUCTURE side for the CPU leg.

  python tools/bench_configs.py [--quick]
"""
import argparse
import dataclasses
import json
import math
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from cases import case_cooling  # noqa: E402
from harness import GpuSim, OracleSim, Problem, RefSim, have_ref, load_cooling_tables  # noqa: E402


def dmr_state(prob):
    """Double Mach reflection ICs (ics/basic_tests.cpp:736-812): Mach-10 shock at 60 degrees through x=1/6."""
    shp = prob.padded_shape()
    g, dx = prob.nbc, prob.dx
    x = prob.xmin[0] + (np.arange(shp[3]) - g + 0.5) * dx
    y = prob.xmin[1] + (np.arange(shp[2]) - g + 0.5) * dx
    X, Y = np.meshgrid(x, y)
    post = X < 1.0 / 6.0 + Y / math.tan(math.pi / 3.0)
    P = np.zeros(shp)
    P[0, 0] = np.where(post, 8.0, 1.4)
    P[1, 0] = np.where(post, 116.5, 1.0)
    P[2, 0] = np.where(post, 7.14470958, 0.0)
    P[3, 0] = np.where(post, -4.125, 0.0)
    return P


def fieldloop_state(prob):
    shp = prob.padded_shape()
    g, dx = prob.nbc, prob.dx
    x = prob.xmin[0] + (np.arange(shp[3]) - g + 0.5) * dx
    y = prob.xmin[1] + (np.arange(shp[2]) - g + 0.5) * dx
    X, Y = np.meshgrid(x, y)
    Az = lambda xx, yy: 1.0e-3 * np.maximum(0.0, 0.3 - np.sqrt(xx * xx + yy * yy))
    P = np.zeros(shp)
    P[0], P[1], P[2], P[3] = 1.0, 1.0, 2.0, 1.0
    P[5, 0] = (Az(X, Y + dx) - Az(X, Y - dx)) / (2 * dx)
    P[6, 0] = -(Az(X + dx, Y) - Az(X - dx, Y)) / (2 * dx)
    return P


def sphere_state(prob, ambient, hot_p, radius, centre):
    shp = prob.padded_shape()
    g, dx = prob.nbc, prob.dx
    ax = [prob.xmin[a] + (np.arange(shp[3 - a]) - g + 0.5) * dx for a in range(3)]
    P = np.zeros(shp)
    for v, val in enumerate(ambient):
        P[v] = val
    for k in range(shp[1]):
        r2 = (ax[2][k] - centre[2]) ** 2 + (ax[1][:, None] - centre[1]) ** 2 + (ax[0][None, :] - centre[0]) ** 2
        P[1, k] = np.where(r2 < radius ** 2, hot_p, ambient[1])
    return P


def configs(quick):
    s = 2 if quick else 1
    L = 3.086e19
    out = []
    for n, solver in ((260, 4), (520, 4), (520, 8)):
        p = Problem(ndim=2, NG=(n // s, n * 80 // 260 // s, 1), eqn="euler", solver=solver, artviscosity=1, etav=0.1, gamma=1.4, cfl=0.4,
                    xmax=(3.25, 1.0, 1.0), bcs=("inflow", "outflow", "reflecting", "DMR", "periodic", "periodic"),
                    internal_bcs=("DMR2",), finishtime=0.2)
        out.append((f"1 DMR 2-D Euler {p.NG[0]}x{p.NG[1]} solver {solver}", p, dmr_state, 200, 1))
    for solver in (7, 4):
        p = Problem(ndim=2, NG=(512 // s, 256 // s, 1), eqn="glm-mhd", solver=solver, artviscosity=1, etav=0.1, gamma=5.0 / 3.0, cfl=0.4,
                    xmin=(-1.0, -0.5, 0.0), xmax=(1.0, 0.5, 1.0), finishtime=2.0)
        out.append((f"2 FieldLoop 2-D GLM-MHD {p.NG[0]}x{p.NG[1]} solver {solver}", p, fieldloop_state, 200, 1))
    n = 256 // s
    p = Problem(ndim=3, NG=(n, n, n), eqn="euler", solver=4, artviscosity=1, etav=0.1, gamma=5.0 / 3.0, cfl=0.3, xmax=(30.86e18,) * 3,
                bcs=("reflecting", "outflow") * 3, finishtime=1.58e12)
    bw = lambda pr: sphere_state(pr, (2.34e-22, 1.38e-11, 0, 0, 0), 3.0 * 1.0e51 * (pr.gamma - 1) / (4 * math.pi * (8 * pr.dx) ** 3),
                                 8 * pr.dx, (0.0, 0.0, 0.0))
    out.append((f"3 blast wave 3-D Euler {n}^3 Roe-CV", p, bw, 20, 4))
    n = 512 // s
    p = Problem(ndim=3, NG=(n, n, n), eqn="glm-mhd", solver=7, artviscosity=1, etav=0.15, gamma=5.0 / 3.0, cfl=0.2, xmin=(-L,) * 3, xmax=(L,) * 3,
                bcs=("outflow",) * 6, finishtime=1e30,
                refvec=(2.338e-22, 1.518e-12, 1e6, 1e6, 1e6, 4e-6, 4e-6, 4e-6, 4e-6) + (1.0,) * 7)
    dte = lambda pr: sphere_state(pr, (2.338e-22, 1.518e-12, 0, 0, 0, 14.2e-6 / math.sqrt(4 * math.pi), 0, 0, 0), 200 * 1.518e-12,
                                  0.25 * 2 * L, (0.0, 0.0, 0.0))
    out.append((f"4 DTE3D-style 3-D GLM-MHD {n}^3 HLLD", p, dte, 10, 8))
    n = 384 // s
    base = case_cooling("euler", 8, NG=(n, n, n), ntracer=1, mp_limit=1)
    Lw = 3.160064e18
    wind = dict(pos=(0.0, 0.0, 0.0), radius=1.543e17 * (128 / 128), mdot=1.0e-7, vinf=1500.0, vrot=0.0, temp=3.0e4, rstar=6.96e11,
                bsrf=10.0, tr=(1.0, 0.0, 0.0, 0.0))
    p = dataclasses.replace(base, xmax=(Lw,) * 3, internal_bcs=("stellar-wind",), winds=(wind,),
                            bcs=("reflecting", "one-way-outflow") * 3)
    amb = lambda pr: sphere_state(pr, (2.124e-24, 2.209e-12, 0, 0, 0, 0.0), 2.209e-12, 0.0, (0.0, 0.0, 0.0))
    out.append((f"5 Wind3D-style 3-D Euler+cooling+wind {n}^3 HLL", p, amb, 10, 6))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated config numbers, e.g. 3,5")
    ap.add_argument("--lib", default=None, help="another build of libpion_b200.so (kernel A/B experiments)")
    args = ap.parse_args()
    if args.lib:
        from pion_b200.capi import load_library
        load_library(args.lib)
    tables = load_cooling_tables()
    rows = []
    for name, prob, icfn, nsteps, cpu_div in configs(args.quick):
        if args.only and name.split()[0] not in args.only.split(","):
            continue
        tab = tables if prob.cooling else None
        P0 = icfn(prob)
        g = GpuSim(prob, tables=tab)
        g.set_state(P0)
        g.init_after_state()
        g.run(3)
        g.ctx.sync()
        t0 = time.perf_counter()
        g.run(nsteps)
        g.ctx.sync()
        tg = time.perf_counter() - t0
        errs = g.error_counts()
        g.close()
        ncell = prob.NG[0] * prob.NG[1] * prob.NG[2]
        gpu_rate = ncell * nsteps / tg
        cpu_rate, cpu_cells = None, None
        if not args.no_cpu:
            # CPU leg: same problem family on a grid reduced by cpu_div per axis, one host core
            NGc = tuple(max(8, n // cpu_div) if a < prob.ndim else 1 for a, n in enumerate(prob.NG))
            pc = dataclasses.replace(prob, NG=NGc)
            if prob.winds:
                w = dict(prob.winds[0]); w["radius"] = w["radius"] * cpu_div / 1.0 if False else w["radius"]
                pc = dataclasses.replace(pc, winds=(w,))
            sim = (RefSim if have_ref() else OracleSim)(pc) if not prob.cooling else ((RefSim(pc)) if have_ref() else OracleSim(pc, tables=tab))
            sim.set_state(icfn(pc))
            sim.init_after_state()
            sim.run(1)
            nc = max(2, min(nsteps, 6))
            t0 = time.perf_counter()
            sim.run(nc)
            tc = time.perf_counter() - t0
            sim.close()
            cpu_cells = NGc[0] * NGc[1] * NGc[2]
            cpu_rate = cpu_cells * nc / tc
        nvar = prob.nvar
        frac = gpu_rate * 5 * nvar * 8 / 6556.5e9
        rows.append(dict(config=name, cells=ncell, steps=nsteps, gpu_cell_updates_per_s=gpu_rate, ms_per_step=1e3 * tg / nsteps,
                         hbm_roofline_frac=frac, cpu_1core_cell_updates_per_s=cpu_rate, cpu_cells=cpu_cells,
                         cpu_kind="reference" if have_ref() else "port", neg_rho=errs[0], neg_p_fixups=errs[1]))
        print(f"| {name} | {ncell:.3g} | {gpu_rate:.3e} | {1e3 * tg / nsteps:.3f} | {100 * frac:.1f} % | "
              f"{cpu_rate if cpu_rate is None else format(cpu_rate, '.3e')} | {'' if cpu_rate is None else format(gpu_rate / cpu_rate, '.0f')} |", flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "r01_configs.json").write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
