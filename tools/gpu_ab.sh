#!/bin/bash
# Developer script: A/B bench of the default library and every build under pion_b200/variants, then an
# optional full ncu capture of one variant.  usage: tools/gpu_ab.sh <tag> [ncu_variant.so|default] [size]
TAG=${1:-x}; NCUV=${2:-}; SIZE=${3:-512}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in pion_b200/libpion_b200.so pion_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so)
  timeout 600 python bench.py --lib $PWD/$lib --size $SIZE --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ab_${TAG}_$name.log 2>&1
  echo "$name: $(grep -h '^{' gpurun_out/ab_${TAG}_$name.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('value=%.4g c-u/s  ms/step=%.3f  stage_avg_ms=%.3f  frac=%.4f  clocks=%s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], d['clocks']['sm_mhz']))
")"
done
if [ -n "$NCUV" ]; then
  lib=pion_b200/libpion_b200.so; [ "$NCUV" != default ] && lib=pion_b200/variants/$NCUV
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_sweep -s 6 -c 2 -o gpurun_out/prof_sweep_$TAG -f python bench.py --lib $PWD/$lib --size $SIZE --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu exit $?"
fi
