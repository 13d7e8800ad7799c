#!/bin/bash
# round 2: A/B of the 16-byte corner-pair loads (base.so = AC_PAIR_LOADS=0, pair.so = 1), each with and without the shifted table
mkdir -p gpurun_out
: > gpurun_out/r02t_variants.log
for f in base pair; do
  for sh in 1 0; do
    AC_TABLE_SHIFT=$sh AC_LIB_PATH=$PWD/avatarcraft_b200/_variants/$f.so timeout 120 python scripts/variant_bench.py ${f}_shift$sh 2>&1 | grep -E "VARIANT|Error|error" | tail -3 >> gpurun_out/r02t_variants.log
  done
done
cat gpurun_out/r02t_variants.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kernel.py -q > gpurun_out/r02t_pytest.log 2>&1; tail -3 gpurun_out/r02t_pytest.log
