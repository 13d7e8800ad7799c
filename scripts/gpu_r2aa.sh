#!/bin/bash
# round 2: sd_gemm with a two-stage ring (3 CTAs/SM) vs three stages (2 CTAs/SM), both with the two-half epilogue
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sds.py -q -x > gpurun_out/r02aa_pytest_sds.log 2>&1; tail -3 gpurun_out/r02aa_pytest_sds.log
timeout 200 python scripts/profile_guidance.py > gpurun_out/r02aa_guidance_st2.txt 2>&1
AC_LIB_PATH=$PWD/avatarcraft_b200/_variants/st3.so timeout 200 python scripts/profile_guidance.py > gpurun_out/r02aa_guidance_st3.txt 2>&1
for f in st2 st3; do echo == $f; grep -E "pixel_gradient|native VAE|sd_gemm|Self CUDA" gpurun_out/r02aa_guidance_$f.txt | cut -c1-72,150-250; done
