#!/bin/bash
# Developer script: GPU suite (optionally a -k subset) + A/B bench of the default library and every build under
# pion_b200/variants.  usage: tools/gpu_quick.sh <tag> [pytest -k expression]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-q}; K=${2:-}
if [ -n "$K" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log
else
  timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log
fi
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_$T.log | tail -8
tools/gpu_ab.sh $T
