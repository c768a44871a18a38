#!/bin/bash
# Developer script: BASELINE config 5 (Wind3D-style 384^3 Euler + cooling + wind + tracer) on one GPU: bench line,
# ncu launch list, full ncu capture of the cooling kernels and the Euler stage kernel.  usage: tools/gpu_wind_prof.sh <tag>
T=${1:-x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/bench_wind_mgpu.py --size 384 --steps 10 > gpurun_out/wind_1gpu_$T.log 2>&1; echo "wind1 exit $?"; grep -h '^{' gpurun_out/wind_1gpu_$T.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_wind384_$T.csv python tools/bench_wind_mgpu.py --size 384 --steps 2 --warmup 2 > gpurun_out/ncu_lw_$T.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cooling_dU|k_mp_dt|k_stage_sweep' -s 8 -c 5 -o gpurun_out/prof_wind384_$T -f python tools/bench_wind_mgpu.py --size 384 --steps 2 --warmup 2 > gpurun_out/ncu_w_$T.log 2>&1; echo "ncu full exit $?"
