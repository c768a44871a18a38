#!/bin/bash
# Developer script: parity of the DEFAULT build on the Euler / tracer / cooling / golden cases, then A/B against a variant on
# BASELINE configs 3 and 5.  usage: gpu_r02E.sh <tag> <variant>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02E}; V=${2:-nopbs}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "euler or tracer or cool or wind or golden or tile" > gpurun_out/pytest_${T}.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_${T}.log; tail -3 gpurun_out/pytest_${T}.log
for lib in pion_b200/libpion_b200.so pion_b200/variants/$V.so; do
  name=$(basename $lib .so)
  echo "== $name"
  timeout 600 python tools/bench_configs.py --no-cpu --only 3,5 --lib $PWD/$lib 2>&1 | grep "^|" | tee gpurun_out/configs_${T}_$name.log
done
