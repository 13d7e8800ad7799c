#!/bin/bash
# round 2 final ncu evidence: training-path kernels after the feature cache / stencil-ordered scatter; launch list of one native step
mkdir -p gpurun_out
TAG=${1:-r02H}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sdf_backward_mma|sdf_scatter|forward_sdf_tc" \
    -s 9 -c 3 -f -o gpurun_out/${TAG}_train_kernels python scripts/profile_train_step.py --plain > gpurun_out/${TAG}_ncu_train.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_train.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_step_launches.csv \
    python scripts/profile_train_step.py --plain > gpurun_out/${TAG}_train_under_ncu.log 2>&1
ls -la gpurun_out | grep ${TAG}
