#!/bin/bash
# Developer script: one GPU session = parity suite, A/B bench (TMA sweep vs LDG sweep), launch list and
# one full ncu capture of the stage kernel.  usage: tools/gpu_round.sh <tag> [quick]
TAG=${1:-x}; QUICK=${2:-}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -z "$QUICK" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
else
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$QUICK" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
fi
tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$TAG.log 2>&1; echo "bench exit $?"
PION_B200_NO_TMA=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_ldg.log 2>&1
grep -h '^{' gpurun_out/bench_$TAG.log gpurun_out/bench_${TAG}_ldg.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('value=%.4g c-u/s  ms/step=%.3f  stage_avg_ms=%.3f  frac=%.4f  clocks=%s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], d['clocks']['sm_mhz']))
"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_sweep -s 6 -c 2 -o gpurun_out/prof_sweep_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu exit $?"
