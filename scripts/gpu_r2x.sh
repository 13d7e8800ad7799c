#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_full_step.py > gpurun_out/r02X_full_step_kernels.txt 2>&1
grep -E "pass 2|^void|^\(anon|Self C" gpurun_out/r02X_full_step_kernels.txt | cut -c1-72,150-250
