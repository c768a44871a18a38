#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in pion_b200/variants/mmfkj.so pion_b200/variants/oldmm.so pion_b200/libpion_b200.so; do
  echo "== diag $lib"; timeout 300 python tools/gpu_diag.py $PWD/$lib 2>&1 | tail -6
done
tools/gpu_ab.sh r02g
