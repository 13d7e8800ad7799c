#!/bin/bash
# round 2: tcgen05 SDF backward (v2) -- training tests, per-kernel timing of a step
mkdir -p gpurun_out
TAG=${1:-r02d}
timeout 900 python -m pytest tests/test_gpu_training.py -q -s > gpurun_out/${TAG}_pytest_train.log 2>&1
grep -n "passed\|failed\|Error\|rel L2\|relative L2" gpurun_out/${TAG}_pytest_train.log | head -60
timeout 300 python scripts/profile_train_step.py > gpurun_out/${TAG}_train_step_kernels.txt 2>&1
grep -n "kernel\|Memset\|Memcpy\|aten::" gpurun_out/${TAG}_train_step_kernels.txt | head -30
