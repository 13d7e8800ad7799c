#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_guidance.py > gpurun_out/r02Y_guidance_kernels.txt 2>&1
grep -E "pixel_gradient|native VAE|^void|^\(anon|Memset|Self C" gpurun_out/r02Y_guidance_kernels.txt | cut -c1-72,150-250
