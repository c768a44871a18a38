#!/bin/bash
# Developer script (8-GPU box, trimmed): BASELINE config 5 (Wind3D-style 384^3 GLOBAL) on 8 GPUs with the reworked cooling
# kernel, and the default bench line at N = 8 (weak + strong block + parity_mgpu, no e2e leg).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02y}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29522 tools/bench_wind_mgpu.py --size 384 --steps 20 > gpurun_out/wind_8gpu_$T.log 2>&1; echo "wind8 exit $?"
timeout 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_8gpu_$T.log 2>&1; echo "bench8 exit $?"
grep -h '^{' gpurun_out/bench_8gpu_$T.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); s = d.get('strong') or {}; p = d.get('parity_mgpu') or {}
    print('bench8 weak=%.4g ms=%.2f | strong=%.4g ms=%.3f | parity_mgpu=%s ok=%s' % (d['value'], d['ms_per_step'], s.get('value', 0), s.get('ms_per_step', 0), p.get('max_rel_err'), p.get('ok')))
"
grep -h '^{' gpurun_out/wind_8gpu_$T.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); p = d.get('parity_mgpu') or {}
    print('wind8 value=%.4g ms=%.3f stage_share=%.3f parity=%s ok=%s fails=%s' % (d['value'], d['ms_per_step'], d['stage_share_of_step'], p.get('max_rel_err'), p.get('ok'), d.get('cooling_integration_failures')))
"
