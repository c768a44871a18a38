#!/bin/bash
# Developer script (round 2, first GPU session): full GPU suite, the default bench line, the strict build's
# parity, A/B of the library variants, launch list + one full ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02a
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/smi_$T.log 2>&1
nproc >> gpurun_out/smi_$T.log
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log
tail -3 gpurun_out/pytest_$T.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_bench_state.py -m gpu -q --pion-lib pion_b200/variants/strict.so > gpurun_out/pytest_${T}_strict.log 2>&1; echo "strict pytest exit $?" | tee -a gpurun_out/pytest_${T}_strict.log
tail -3 gpurun_out/pytest_${T}_strict.log
timeout 900 python bench.py > gpurun_out/bench_$T.log 2>&1; echo "bench exit $?"
grep -h '^{' gpurun_out/bench_$T.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('value=%.4g  ms/step=%.3f  stage_avg_ms=%.3f  frac=%.4f  share=%.3f e2e=%.3g parity=%s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['stage_share_of_step'], d['e2e']['value'], d.get('parity', {}).get('max_rel_err')))
"
tools/gpu_ab.sh $T
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_l_$T.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_sweep -s 6 -c 2 -o gpurun_out/prof_sweep_$T -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_$T.log 2>&1; echo "ncu full exit $?"
