#!/bin/bash
# round 2: ncu evidence of the final kernels -- training-path kernels (set full), render kernel (set full, traffic), launch lists
mkdir -p gpurun_out
TAG=${1:-r02p}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sdf_backward_mma|sdf_scatter|shade_backward|shade_forward|forward_sdf_tc" \
    -s 15 -c 5 -f -o gpurun_out/${TAG}_train_kernels python scripts/profile_train_step.py --plain > gpurun_out/${TAG}_ncu_train.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_train.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nsr_render_tc -s 3 -c 1 -f -o gpurun_out/${TAG}_render \
    env AC_BENCH_SKIP_SDS=1 AC_BENCH_SKIP_WARP=1 python bench.py --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_render.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_render.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    env AC_BENCH_SKIP_SDS=1 AC_BENCH_SKIP_WARP=1 python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_step_launches.csv \
    python scripts/profile_train_step.py --plain > gpurun_out/${TAG}_train_under_ncu.log 2>&1
ls -la gpurun_out | grep ${TAG}
