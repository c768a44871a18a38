#!/bin/bash
# Developer script: end-of-round verification = full GPU suite, repeated runs of the flag-heavy MHD-HLLD cases,
# smoke(), default bench line, reference arm, launch list, full ncu capture.  usage: tools/gpu_final.sh <tag>
TAG=${1:-x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_$TAG.log
for av in 0 1; do python tools/tma_check.py i-mhd 7 $av 5 2>&1 | cut -c1-70 | sort | uniq -c; done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench512_$TAG.log 2>&1; echo "bench exit $?"; grep -h '^{' gpurun_out/bench512_$TAG.log | cut -c1-300
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.log 2>&1; echo "ref exit $?"; grep -h '^{' gpurun_out/bench_ref_$TAG.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_sweep -s 6 -c 2 -o gpurun_out/prof_sweep_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu full exit $?"
