#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_guidance.py > gpurun_out/r02Y_guidance_kernels.txt 2>&1
grep -E "pixel_gradient|native VAE|^void|^\(anon|Memset|Self C" gpurun_out/r02Y_guidance_kernels.txt | cut -c1-72,150-250
timeout 500 python -m pytest tests/test_gpu_warp.py -q -x > gpurun_out/r02Z_pytest_warp.log 2>&1; tail -5 gpurun_out/r02Z_pytest_warp.log
timeout 200 python scripts/bench_warp.py --profile --exact-all > gpurun_out/r02Z_warp_exact.txt 2>&1
timeout 200 python scripts/bench_warp.py --profile > gpurun_out/r02Z_warp_skip.txt 2>&1
for f in exact skip; do grep -E "ms_per_frame|^void|^\(anon|Self C" gpurun_out/r02Z_warp_$f.txt | cut -c1-62,130-230; done
