#!/bin/bash
# Developer script: full GPU suite on the current build + headline / Wind3D lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02s}
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log; tail -3 gpurun_out/pytest_$T.log
timeout 300 python tools/bench_wind_mgpu.py --size 384 --steps 10 > gpurun_out/wind_1gpu_$T.log 2>&1; grep -h '^{' gpurun_out/wind_1gpu_$T.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ab_${T}_default.log 2>&1
grep -h '^{' gpurun_out/ab_${T}_default.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('headline value=%.4g c-u/s  ms/step=%.3f  stage_avg_ms=%.3f' % (d['value'], d['ms_per_step'], r['avg_launch_ms']))
"
