#!/bin/bash
# round 2: red.global.add.v4.f32 corner-pair merging in the training scatter -- parity tests + kernel table
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_training.py tests/test_gpu_frame_ops.py -q -x > gpurun_out/r02w_pytest.log 2>&1; tail -3 gpurun_out/r02w_pytest.log
timeout 200 python scripts/profile_train_step.py > gpurun_out/r02w_train_step_kernels.txt 2>&1; grep -E "scatter|sdf_backward|forward_sdf|Self CUDA time" gpurun_out/r02w_train_step_kernels.txt | cut -c1-60,150-260
