#!/bin/bash
# Developer script: parity of the per-order tile build (3-D cases), then A/B of the variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02o}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bench_state.py tests/test_golden.py -m gpu -x -q -k "3d or tile or hot_sphere or strong or bench or golden or full_size" > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log; tail -3 gpurun_out/pytest_$T.log
tools/gpu_ab.sh $T default
