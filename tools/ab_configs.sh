#!/bin/bash
# Developer script: run tools/bench_configs.py --only "$1" for the default library and every variant build
cd "$(dirname "$0")/.."
for lib in pion_b200/libpion_b200.so pion_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  echo "== $(basename $lib)"
  python tools/bench_configs.py --lib $PWD/$lib --no-cpu --only "$1" 2>&1 | grep "^|"
done
