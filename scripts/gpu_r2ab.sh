#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sds.py -q -x > gpurun_out/r02ab_pytest_sds.log 2>&1; tail -3 gpurun_out/r02ab_pytest_sds.log
timeout 300 python scripts/bench_unet_graph.py 2>&1 | grep -v Warn | tee gpurun_out/r02ab_unet_graph.txt
timeout 200 python scripts/profile_guidance.py 2>&1 | grep -E "pixel_gradient|native VAE|Self C" | tee gpurun_out/r02ab_guidance.txt
