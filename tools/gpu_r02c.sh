#!/bin/bash
# Developer script (round 2, third GPU session): GPU suite on the current build, default bench line,
# launch list, BASELINE configs, one full ncu capture of the cooling kernel (Wind3D-style step).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02c}
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log
grep -E "passed|failed|FAILED" gpurun_out/pytest_$T.log | tail -8
timeout 900 python bench.py > gpurun_out/bench_$T.log 2>&1; echo "bench exit $?"
grep -h '^{' gpurun_out/bench_$T.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print('value=%.4g  ms/step=%.3f  stage_avg_ms=%.3f  frac=%.4f  share=%.3f e2e=%.3g parity_ok=%s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['stage_share_of_step'], d['e2e']['value'], d.get('parity', {}).get('ok')))
"
timeout 900 python tools/bench_configs.py --no-cpu > gpurun_out/configs_$T.log 2>&1; grep "^|" gpurun_out/configs_$T.log
cp gpurun_out/r01_configs.json gpurun_out/configs_$T.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_wind_$T.csv python tools/wind_prof.py > gpurun_out/ncu_lw_$T.log 2>&1; echo "ncu wind list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cooling_dU -s 4 -c 1 -o gpurun_out/prof_cooling_$T -f python tools/wind_prof.py > gpurun_out/ncu_cool_$T.log 2>&1; echo "ncu cooling exit $?"
