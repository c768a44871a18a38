#!/usr/bin/env python
"""Developer script: write profiles/ncu_traffic.json (what bench.py's `roofline.traffic` / `roofline.fp64` quote) from
an `ncu --set full` capture of one predictor + one corrector launch of the stage kernel.

  python tools/ncu_traffic.py gpurun_out/prof_sweep_TAG.ncu-rep <cells per axis> <csrc digest of the captured build> \
         <profiles/ summary the numbers are also written to>

The digest is bench.py's csrc_digest() of the sources the captured library was built from (the GPU script prints it
next to the capture); bench.py reports `same_build` by comparing it with the digest of the sources it runs on.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    rep, size, digest, summary = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ncell = float(size) ** 3

    def val(r, k):
        v = float(r[hdr.index(k)].replace(",", ""))
        u = units[hdr.index(k)]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u, 1.0)
        return v * scale

    stages = {}
    fp64_inst = 0.0
    pipe = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if "k_stage_sweep" not in name:
            continue
        # the last-but-one template argument is ORDER (k_stage_sweep_tma<EQ, SOLVER, FKJ, TY, MINB, NTR, ORDER, RING>)
        targs = [t.strip() for t in name[name.index("<") + 1:name.index(">")].split(",")]
        order = int(targs[6].strip("()int "))
        key = "predictor" if order == 1 else "corrector"
        cyc = val(r, "sm__cycles_elapsed.avg")
        fp = 0.0
        for op in ("dfma", "dmul", "dadd"):
            k = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum"
            fp += val(r, k) if k in hdr else val(r, k + ".per_cycle_elapsed") * cyc
        nvar = 9
        stages[key] = {
            "kernel": name,
            "read": val(r, "dram__bytes_read.sum"),
            "write": val(r, "dram__bytes_write.sum"),
            "algorithmic": int((2 if order == 1 else 3) * nvar * 8 * ncell),
            "duration_ms_ncu": val(r, "gpu__time_duration.sum"),
            "fp64_arith_inst_per_cell": fp / ncell,
            "fp64_pipe_active_pct": val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": val(r, "launch__registers_per_thread"),
            "block_size": val(r, "launch__block_size"),
        }
        fp64_inst += fp / ncell
        pipe.append((stages[key]["fp64_pipe_active_pct"], stages[key]["duration_ms_ncu"]))
    assert set(stages) == {"predictor", "corrector"}, stages.keys()
    per_launch = 0.5 * sum(s["read"] + s["write"] for s in stages.values())
    tw = sum(d for _, d in pipe)
    p = ROOT / "profiles" / "ncu_traffic.json"
    d = json.loads(p.read_text()) if p.exists() else {}
    d["source"] = f"{summary} (ncu --set full --clock-control none, one predictor + one corrector launch, {size}^3)"
    d["csrc_digest"] = digest
    d[str(size)] = {"bytes_per_launch": per_launch, **stages}
    d["fp64"] = {
        "fp64_arith_inst_per_cell_update": round(fp64_inst, 1),
        "peak_ginst_per_s": 16940.0,
        "ncu_pipe_fp64_cycles_active_pct": round(sum(a * b for a, b in pipe) / tw, 1),
        "source": f"DFMA+DMUL+DADD thread instructions per cell, predictor + corrector, {summary}; peak = 33.88 TFLOP/s / 2 measured by "
                  "tools/micro/fp64_pipe.cu (nominal 148 SM x 64 lanes x 1.965 GHz = 18.6e12 inst/s)",
    }
    p.write_text(json.dumps(d, indent=1) + "\n")
    print(json.dumps({k: d[k] for k in ("source", "csrc_digest", "fp64")}, indent=1))
    print("bytes_per_launch", per_launch)


if __name__ == "__main__":
    main()
