#!/bin/bash
# Developer script: cooling / wind / tracer parity with the one-launch cooling kernel, then A/B of the Wind3D-style
# 384^3 line and the headline line over the library variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02q}
timeout 900 python -m pytest tests -m gpu -x -q -k "cool or wind or tracer or golden or host_cpp or binding" > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log; tail -3 gpurun_out/pytest_$T.log
for lib in pion_b200/libpion_b200.so pion_b200/variants/*.so; do
  name=$(basename $lib .so)
  timeout 300 python tools/bench_wind_mgpu.py --lib $PWD/$lib --size 384 --steps 10 > gpurun_out/wind_${T}_$name.log 2>&1
  echo "$name: $(grep -h '^{' gpurun_out/wind_${T}_$name.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('value=%.4g ms/step=%.3f stage_share=%.3f' % (d['value'], d['ms_per_step'], d['stage_share_of_step']))
")"
done
timeout 900 python tools/shock_bound.py > gpurun_out/shock_bound_$T.log 2>&1; echo "shock bound exit $?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ab_${T}_default.log 2>&1
grep -h '^{' gpurun_out/ab_${T}_default.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('headline value=%.4g c-u/s  ms/step=%.3f  stage_avg_ms=%.3f' % (d['value'], d['ms_per_step'], r['avg_launch_ms']))
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_wind384_$T.csv python tools/bench_wind_mgpu.py --size 384 --steps 2 --warmup 2 > gpurun_out/ncu_lw_$T.log 2>&1; echo "ncu launches exit $?"
