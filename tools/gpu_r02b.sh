#!/bin/bash
# Developer script (round 2, second GPU session): full GPU suite on the new host paths (box-table shell launch,
# per-axis ghost fill, binding), A/B of tile-height and spin-backoff variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02b}
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log
grep -E "passed|failed|FAILED" gpurun_out/pytest_$T.log | tail -8
tools/gpu_ab.sh $T
