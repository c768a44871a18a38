#!/bin/bash
# Developer script: time every library build under pion_b200/variants (plus the default)
# on the same bench workload.  usage: tools/ab_bench.sh [size] [steps]
SIZE=${1:-256}; STEPS=${2:-4}
cd "$(dirname "$0")/.."
for lib in pion_b200/libpion_b200.so pion_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  out=$(PION_B200_LIB=$PWD/$lib python bench.py --size $SIZE --steps $STEPS --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('value=%.4g c-u/s  ms/step=%.3f  stage_avg_ms=%.3f  frac=%.4f  clocks=%s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], d['clocks']['sm_mhz']))
")
  echo "$(basename $lib): $out"
done
