#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the render kernel.
# Usage: scripts/profile_render.sh <tag> [kernel regex]
TAG=${1:-r02}
KRE=${2:-nsr_render}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 24 --csv --log-file gpurun_out/${TAG}_launches.csv \
    env AC_BENCH_SKIP_SDS=1 python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 3 -c 1 -f -o gpurun_out/${TAG}_render \
    env AC_BENCH_SKIP_SDS=1 python bench.py --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
