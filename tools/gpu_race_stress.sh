#!/bin/bash
# Developer script: the 3-D parity / golden / bench-state tests against the -DPION_RACE_STRESS build (random per-warp delays at
# every hand-over point of the TMA sweep kernel), then the default build's bench line and a full ncu capture of its stage kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02D}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_bench_state.py tests/test_golden.py -m gpu -q -k "3d or tile or hot_sphere or strong or bench or golden or tracer or cool or wind" --pion-lib pion_b200/variants/stress.so > gpurun_out/pytest_${T}_stress.log 2>&1; echo "pytest(stress) exit $?" | tee -a gpurun_out/pytest_${T}_stress.log; tail -4 gpurun_out/pytest_${T}_stress.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$T.log 2>&1; echo "bench exit $?"; grep -h '^{' gpurun_out/bench_$T.log | cut -c1-250
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_sweep -s 6 -c 2 -o gpurun_out/prof_sweep_$T -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_$T.log 2>&1; echo "ncu full exit $?"
