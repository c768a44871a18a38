#!/bin/bash
# Developer script: full GPU suite on the current build, then the parity / golden / bench-state tests against the
# -DPION_STRICT build (the reference's literal expressions: IEEE divisions, two-term sources, fmax / fmin).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02t}
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$T.log; tail -3 gpurun_out/pytest_$T.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_bench_state.py -m gpu -q --pion-lib pion_b200/variants/strict.so > gpurun_out/pytest_${T}_strict.log 2>&1; echo "strict pytest exit $?" | tee -a gpurun_out/pytest_${T}_strict.log
tail -4 gpurun_out/pytest_${T}_strict.log
timeout 600 python bench.py --lib $PWD/pion_b200/variants/strict.so --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${T}_strict.log 2>&1
grep -h '^{' gpurun_out/bench_${T}_strict.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']; p = d.get('parity') or {}
    print('strict build: value=%.4g c-u/s  ms/step=%.3f  stage_avg_ms=%.3f parity_ok=%s max=%s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], p.get('ok'), p.get('max_rel_err_well_conditioned_variables(rho,p,vx,Bx)')))
"
