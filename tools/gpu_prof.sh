#!/bin/bash
# Developer script: launch list of the default bench + one full ncu capture of the two stage kernels (predictor,
# corrector).  usage: tools/gpu_prof.sh <tag> [pytest -k expression]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-p}; K=${2:-}
if [ -n "$K" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -k "$K" > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log
  grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_$T.log | tail -8
fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/bench_$T.log 2>&1; echo "bench exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_l_$T.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_sweep -s 6 -c 2 -o gpurun_out/prof_sweep_$T -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_$T.log 2>&1; echo "ncu full exit $?"
