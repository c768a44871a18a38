#!/bin/bash
# Developer script: parity of an Euler-kernel variant, then A/B on BASELINE configs 3 and 5.  usage: gpu_r02z.sh <tag> <variant>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02z}; V=${2:-eunroll}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "euler or tracer or cool or wind or golden or tile" --pion-lib pion_b200/variants/$V.so > gpurun_out/pytest_${T}_$V.log 2>&1; echo "pytest($V) exit $?" | tee -a gpurun_out/pytest_${T}_$V.log; tail -3 gpurun_out/pytest_${T}_$V.log
for lib in pion_b200/libpion_b200.so pion_b200/variants/$V.so; do
  name=$(basename $lib .so)
  echo "== $name"
  timeout 600 python tools/bench_configs.py --no-cpu --only 3,5 --lib $PWD/$lib 2>&1 | grep "^|" | tee gpurun_out/configs_${T}_$name.log
done
