#!/bin/bash
# Developer script (8-GPU box): the default bench line at N = 8 (weak line + strong block + parity_mgpu), the
# strong block at N = 2 and 4, and BASELINE config 5 (Wind3D-style 384^3, cooling + wind) on 8 and 1 GPUs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02m}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/smi8_$T.log
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_$T.log 2>&1; echo "bench8 exit $?"
timeout 400 $TR --nproc-per-node 8 --master-port 29522 tools/bench_wind_mgpu.py --size 384 --steps 10 > gpurun_out/wind_8gpu_$T.log 2>&1; echo "wind8 exit $?"
timeout 400 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_4gpu_$T.log 2>&1; echo "bench4 exit $?"
timeout 400 $TR --nproc-per-node 2 --master-port 29524 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_2gpu_$T.log 2>&1; echo "bench2 exit $?"
timeout 300 python tools/bench_wind_mgpu.py --size 384 --steps 10 > gpurun_out/wind_1gpu_$T.log 2>&1; echo "wind1 exit $?"
timeout 300 $TR --nproc-per-node 2 --master-port 29525 tools/bench_wind_mgpu.py --size 384 --steps 10 > gpurun_out/wind_2gpu_$T.log 2>&1; echo "wind2 exit $?"
for f in bench_8gpu bench_4gpu bench_2gpu; do grep -h '^{' gpurun_out/${f}_$T.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); s = d.get('strong') or {}; p = d.get('parity_mgpu') or {}
    print('$f weak=%.4g ms=%.2f | strong=%.4g ms=%.3f | parity_mgpu=%s ok=%s' % (d['value'], d['ms_per_step'], s.get('value', 0), s.get('ms_per_step', 0), p.get('max_rel_err'), p.get('ok')))
"; done
for f in wind_8gpu wind_2gpu wind_1gpu; do grep -h '^{' gpurun_out/${f}_$T.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); p = d.get('parity_mgpu') or {}
    print('$f value=%.4g ms=%.3f parity=%s ok=%s fails=%s' % (d['value'], d['ms_per_step'], p.get('max_rel_err'), p.get('ok'), d.get('cooling_integration_failures')))
"; done
