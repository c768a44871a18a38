#!/bin/bash
# Developer script: parity of a variant library on the 3-D cases, then A/B against the default build.  usage: tools/gpu_variant_ab.sh <tag> <variant>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02x}; V=${2:-ymid}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bench_state.py tests/test_golden.py -m gpu -x -q -k "3d or tile or hot_sphere or strong or bench or golden or full_size or tracer or cool or wind" --pion-lib pion_b200/variants/$V.so > gpurun_out/pytest_${T}_$V.log 2>&1; echo "pytest($V) exit $?" | tee -a gpurun_out/pytest_${T}_$V.log; tail -3 gpurun_out/pytest_${T}_$V.log
tools/gpu_ab.sh $T
for lib in pion_b200/libpion_b200.so pion_b200/variants/$V.so; do
  timeout 300 python tools/bench_wind_mgpu.py --lib $PWD/$lib --size 384 --steps 10 2>/dev/null | grep -h '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('wind $lib value=%.4g ms/step=%.3f' % (d['value'], d['ms_per_step']))
"
done
