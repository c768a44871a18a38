#!/bin/bash
# Developer script: round-2 record of one build = full GPU suite, smoke(), default bench line (e2e + CPU baseline +
# parity), reference arm, ncu launch list, full ncu capture of the stage kernel (predictor + corrector) and of the
# cooling kernel of the Wind3D-style configuration, per-config throughput.  usage: tools/gpu_record_all.sh <tag>
TAG=${1:-r02n}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench512_$TAG.log 2>&1; echo "bench exit $?"; grep -h '^{' gpurun_out/bench512_$TAG.log | cut -c1-300
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.log 2>&1; echo "ref exit $?"; grep -h '^{' gpurun_out/bench_ref_$TAG.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_sweep -s 6 -c 2 -o gpurun_out/prof_sweep_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu full exit $?"
timeout 300 python tools/bench_wind_mgpu.py --size 384 --steps 10 > gpurun_out/wind_1gpu_$TAG.log 2>&1; echo "wind1 exit $?"; grep -h '^{' gpurun_out/wind_1gpu_$TAG.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_wind384_$TAG.csv python tools/bench_wind_mgpu.py --size 384 --steps 2 --warmup 2 > gpurun_out/ncu_lw_$TAG.log 2>&1; echo "ncu wind launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cooling_dU|k_mp_dt|k_stage_sweep' -s 6 -c 4 -o gpurun_out/prof_wind384_$TAG -f python tools/bench_wind_mgpu.py --size 384 --steps 2 --warmup 2 > gpurun_out/ncu_w_$TAG.log 2>&1; echo "ncu wind full exit $?"
timeout 1200 python tools/bench_configs.py --no-cpu > gpurun_out/configs_$TAG.log 2>&1; echo "configs exit $?"; grep "^|" gpurun_out/configs_$TAG.log
cp gpurun_out/r01_configs.json gpurun_out/configs_$TAG.json 2>/dev/null
