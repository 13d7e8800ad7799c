#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_training.py -q -x > gpurun_out/r02ag_pytest.log 2>&1; tail -3 gpurun_out/r02ag_pytest.log
timeout 200 python scripts/profile_full_step.py > gpurun_out/r02ag_full_step_kernels.txt 2>&1
grep -E "pass 2|forward_sdf_tc|scatter|backward_mma|Self CUDA" gpurun_out/r02ag_full_step_kernels.txt | cut -c1-72,150-250
AC_STENCIL_FWD=flat timeout 200 python scripts/profile_full_step.py 2>&1 | grep -E "pass 2|forward_sdf_tc" | cut -c1-72,150-250
