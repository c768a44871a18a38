#!/bin/bash
# Developer script: full GPU suite, then A/B of the Euler tile shapes (rows per tile / blocks per SM) on BASELINE configs 1, 3, 5.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02v}
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log; tail -3 gpurun_out/pytest_$T.log
for lib in pion_b200/libpion_b200.so pion_b200/variants/e16.so pion_b200/variants/e12.so; do
  name=$(basename $lib .so)
  echo "== $name"
  timeout 600 python tools/bench_configs.py --no-cpu --only 1,3,5 --lib $PWD/$lib 2>&1 | grep "^|" | tee gpurun_out/configs_${T}_$name.log
done
