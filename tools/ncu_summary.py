#!/usr/bin/env python
"""Developer script: condense an `ncu --set full` report (.ncu-rep, read here with
`ncu -i ... --page raw --csv`) into the few counters DESIGN.md / profiles/ quote.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [ncell_per_launch] > profiles/xxx.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe inst % of peak"),
    ("smsp__issue_active.avg.per_cycle_active", "issue slots busy (per SMSP cycle)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread inst"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread inst"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread inst"),
    ("sass__inst_executed_register_spilling", "spill instructions (warp)"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local loads (warp inst)"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local stores (warp inst)"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    ncell = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{rep}`\n")
    for r in rows[2:]:
        get = lambda k: (r[hdr.index(k)], units[hdr.index(k)]) if k in hdr else (None, None)
        print(f"## launch {r[hdr.index('ID')]}: `{r[hdr.index('Kernel Name')]}`\n")
        print("| counter | value |\n|---|---|")
        vals = {}
        for k, label in KEYS:
            v, u = get(k)
            if v is None:
                continue
            vals[k] = v
            print(f"| {label} (`{k}`) | {v} {u} |")
        # derived
        try:
            cyc = float(vals["sm__cycles_elapsed.avg"])
            fp = 0.0
            for k in ("dfma", "dmul", "dadd"):
                kk = f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum"
                if kk in vals:
                    fp += float(vals[kk])
                else:
                    v, _ = get(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed")
                    fp += float(v) * cyc
            print(f"| FP64 arithmetic thread inst (dfma+dmul+dadd) | {fp:.4g} |")
            if ncell:
                print(f"| FP64 arithmetic inst per cell | {fp / ncell:.1f} |")
                wi = float(vals["smsp__inst_executed.sum"]) * float(vals["smsp__thread_inst_executed_per_inst_executed.ratio"])
                print(f"| all thread inst per cell | {wi / ncell:.1f} |")
        except Exception as e:  # noqa: BLE001
            print(f"| derived | unavailable ({e}) |")
        print("\nstall reasons (warps per issue-active cycle):\n")
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    st.append((float(r[i]), h[len(STALLS):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        for v, n in sorted(st, reverse=True)[:8]:
            print(f"- {n}: {v:.3f}")
        print()


if __name__ == "__main__":
    main()
