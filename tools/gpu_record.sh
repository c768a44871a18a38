#!/bin/bash
# Developer script: the numbers that go under profiles/ for a kernel revision: default bench line (e2e +
# CPU baseline), ncu launch list of the same command, per-config throughput.  usage: tools/gpu_record.sh <tag>
TAG=${1:-x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench512_$TAG.log 2>&1; echo "bench exit $?"; grep -h '^{' gpurun_out/bench512_$TAG.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu launches exit $?"
timeout 1200 python tools/bench_configs.py --no-cpu > gpurun_out/configs_$TAG.log 2>&1; echo "configs exit $?"; grep "^|" gpurun_out/configs_$TAG.log
cp gpurun_out/r01_configs.json gpurun_out/configs_$TAG.json 2>/dev/null
