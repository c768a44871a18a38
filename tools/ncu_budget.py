#!/usr/bin/env python
"""Developer script: per-component instruction budget of the stage kernels from an `ncu --set full --import-source on`
capture.  Reads the report's source page (`ncu -i rep --page source --csv --print-source cuda,sass`), attributes every
SASS instruction to the device function whose source line it came from (-lineinfo: the innermost inlined callee), and
sums executed thread instructions per component and per cell-update.

  python tools/ncu_budget.py gpurun_out/prof.ncu-rep <cells per launch> > profiles/xxx_budget.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

# device function -> component of the finite-volume update
COMPONENT = {
    "minmod": "reconstruction", "add_limited": "reconstruction", "edge_states_tile": "reconstruction", "lds_prim": "tile loads (LDS)",
    "tracer_edges_tile": "reconstruction",
    "fast_rcp": "rcp / sqrt / rsqrt (Newton steps)", "fast_sqrt": "rcp / sqrt / rsqrt (Newton steps)", "fast_sqrt_pos": "rcp / sqrt / rsqrt (Newton steps)",
    "fast_rsqrt": "rcp / sqrt / rsqrt (Newton steps)", "rcp_refine": "rcp / sqrt / rsqrt (Newton steps)", "pdiv": "rcp / sqrt / rsqrt (Newton steps)",
    "psqrt": "rcp / sqrt / rsqrt (Newton steps)",
    "cfast2_ir": "fast speeds (HLLD signal speeds + FKJ98 mean state)", "cfast_components": "fast speeds (HLLD signal speeds + FKJ98 mean state)",
    "mhd_HLLD": "HLLD (Miyoshi & Kusano, one-sided form)", "mhd_HLL": "HLL fallback (HLLD->HLL switch) + its UtoP", "hlld_speeds": "HLL fallback (HLLD->HLL switch) + its UtoP",
    "intercell_flux": "GLM Dedner state + FKJ98 viscosity (InterCellFlux)",
    "acc_sources": "Powell + GLM sources", "acc_flux_diff": "flux difference / dU", "cons_diff": "flux difference / dU",
    "cons_shfl_down": "flux exchange (shfl / smem)", "cons_to_smem": "flux exchange (shfl / smem)", "cons_from_smem": "flux exchange (shfl / smem)",
    "PtoU": "CellAdvanceTime (PtoU, UtoP, floors, CFL dt)", "UtoP": "CellAdvanceTime (PtoU, UtoP, floors, CFL dt)", "PUtoFlux": "HLL fallback (HLLD->HLL switch) + its UtoP",
    "PtoU_mhd_ideal": "HLL fallback (HLLD->HLL switch) + its UtoP",
    "cell_advance_time_pb": "CellAdvanceTime (PtoU, UtoP, floors, CFL dt)", "cell_advance_time": "CellAdvanceTime (PtoU, UtoP, floors, CFL dt)",
    "cell_time_step": "CellAdvanceTime (PtoU, UtoP, floors, CFL dt)", "store_prim": "CellAdvanceTime (PtoU, UtoP, floors, CFL dt)",
    "load_prim": "CellAdvanceTime (PtoU, UtoP, floors, CFL dt)", "stage_block_epilogue": "CellAdvanceTime (PtoU, UtoP, floors, CFL dt)",
    "pmax": "max / min selects", "pmin": "max / min selects", "sq": "other",
    "k_stage_sweep_tma": "kernel body (loop control, barriers, TMA issue, flags, address arithmetic)",
    "mbar_wait_spin": "barrier waits (try_wait spins)", "mbar_arrive": "kernel body (loop control, barriers, TMA issue, flags, address arithmetic)",
    "tma_load_plane": "kernel body (loop control, barriers, TMA issue, flags, address arithmetic)", "mbar_expect_tx": "kernel body (loop control, barriers, TMA issue, flags, address arithmetic)",
}
FUNC_RE = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__)[^;(]*?\b([A-Za-z_]\w*)\s*\(")
FP64 = ("DFMA", "DMUL", "DADD")


def function_map(path):
    """start line -> device function name for one source file (the working tree's copy: run this on a capture of
    the current build)."""
    starts = []
    try:
        lines = open(path).read().splitlines()
    except OSError:
        return starts
    for n, ln in enumerate(lines, 1):
        if ln.lstrip().startswith("//"):
            continue
        m = FUNC_RE.match(ln)
        if m:
            starts.append((n, m.group(1)))
        elif n >= 2 and "__global__" in lines[n - 2] and re.match(r"\s+(k_\w+)\(", ln):  # __global__ ... \n    k_name(
            starts.append((n - 1, re.match(r"\s+(k_\w+)\(", ln).group(1)))
    return starts


def func_of(starts, line):
    name = "?"
    for n, f in starts:
        if n <= line:
            name = f
        else:
            break
    return name


def main():
    rep, ncell = sys.argv[1], float(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # sections: "File Path", "Function Name" (kernel), a header row, then per CUDA source line one row (Line No, Source)
    # followed by its SASS rows (address, SASS, counters).  Every section is listed twice and an instruction can appear
    # under several lines: count each (kernel, address) once, under the first line it is listed with.
    data = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, 0]))  # kernel -> function -> [fp64, all, samples]
    seen = set()
    maps = {}
    fname = kernel = None
    ix = None
    line = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1]
            if fname not in maps:
                maps[fname] = function_map(fname)
            continue
        if r[0] == "Function Name":
            kernel = r[1]; continue
        if r[0] == "Line No":
            ix = {}
            for i, h in enumerate(r):
                ix.setdefault(h, i)
            continue
        if ix is None or len(r) < 4:
            continue
        if r[0].strip():
            line = int(r[0]); continue
        addr, sass = r[2], r[3].strip()
        if not addr.startswith("0x") or (kernel, addr) in seen:
            continue
        seen.add((kernel, addr))
        try:
            ti = int(r[ix["Thread Instructions Executed"]] or 0)
            sm = int(r[ix["# Samples"]] or 0)
        except (ValueError, KeyError, IndexError):
            continue
        op = re.match(r"(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", sass)
        op = op.group(1) if op else "?"
        f = func_of(maps[fname], line) if fname.startswith("/root/repo") or "pion_b200" in fname else "cuda headers (shfl, atomics, math)"
        d = data[kernel][f]
        d[1] += ti
        d[2] += sm
        if op in FP64:
            d[0] += ti
    print(f"# Instruction budget per component, `{rep}` ({ncell:.0f} cells per launch)\n")
    print("Thread instructions per cell-stage; FP64 = DFMA + DMUL + DADD (an FMA counts once); `samples` = share of the kernel's warp-state samples.\n")
    tot_all = collections.Counter(); tot_fp = collections.Counter()
    for kernel, funcs in data.items():
        comp = collections.defaultdict(lambda: [0, 0, 0])
        for f, (fp, al, sm) in funcs.items():
            c = COMPONENT.get(f, f"other ({f})")
            comp[c][0] += fp; comp[c][1] += al; comp[c][2] += sm
        tfp = sum(v[0] for v in comp.values()); tal = sum(v[1] for v in comp.values()); tsm = sum(v[2] for v in comp.values()) or 1
        print(f"## `{kernel}`\n")
        print("| component | FP64 arith / cell | all inst / cell | samples |\n|---|---|---|---|")
        for c, (fp, al, sm) in sorted(comp.items(), key=lambda kv: -kv[1][1]):
            if al / ncell < 0.5:
                continue
            print(f"| {c} | {fp / ncell:.1f} | {al / ncell:.1f} | {100 * sm / tsm:.1f} % |")
            tot_all[c] += al / ncell; tot_fp[c] += fp / ncell
        print(f"| **total** | **{tfp / ncell:.1f}** | **{tal / ncell:.1f}** | 100 % |\n")
    print("## both stages (one cell-update)\n\n| component | FP64 arith | all inst |\n|---|---|---|")
    for c, al in sorted(tot_all.items(), key=lambda kv: -kv[1]):
        print(f"| {c} | {tot_fp[c]:.1f} | {al:.1f} |")
    print(f"| **total** | **{sum(tot_fp.values()):.1f}** | **{sum(tot_all.values()):.1f}** |")


if __name__ == "__main__":
    main()
