#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02f
for lib in pion_b200/libpion_b200.so pion_b200/variants/base.so pion_b200/variants/oldfkj.so; do
  echo "== diag $lib"; timeout 300 python tools/gpu_diag.py $PWD/$lib 2>&1 | tail -4
done
tools/gpu_ab.sh $T
for lib in pion_b200/libpion_b200.so pion_b200/variants/base.so; do
  n=$(basename $lib .so)
  timeout 600 ncu --metrics smsp__inst_executed.sum,sm__cycles_elapsed.max,sm__inst_executed_pipe_fp64.sum,smsp__issue_active.avg,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct --clock-control none -k regex:k_stage_sweep -s 6 -c 2 --csv --log-file gpurun_out/metrics_${T}_$n.csv python bench.py --lib $PWD/$lib --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_m_${T}_$n.log 2>&1; echo "ncu $n exit $?"
done
