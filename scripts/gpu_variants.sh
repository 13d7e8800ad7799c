#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants.log
for f in avatarcraft_b200/_variants/*.so; do
  AC_LIB_PATH=$PWD/$f timeout 120 python scripts/variant_bench.py 2>&1 | grep -E "VARIANT|Error|error" | tail -3 >> gpurun_out/variants.log
done
cat gpurun_out/variants.log
