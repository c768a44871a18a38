#!/bin/bash
# Developer script: build library variants into pion_b200/variants/<name>.so
# usage: tools/variants.sh name "EXTRA flags" [name "flags" ...]
cd "$(dirname "$0")/../pion_b200/csrc"
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  rm -rf /tmp/pion_var_$name; mkdir -p /tmp/pion_var_$name
  make -j8 OBJDIR=/tmp/pion_var_$name OUTDIR=/tmp/pion_var_$name EXTRA="$flags" 2>&1 | grep -E "error|spill" | grep -v " 0 bytes spill" | head -5
  cp /tmp/pion_var_$name/libpion_b200.so ../variants/$name.so && echo "built $name"
done
