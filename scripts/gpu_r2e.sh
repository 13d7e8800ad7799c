#!/bin/bash
# round 2: ncu --set full of the SDF backward kernel (v2 = tcgen05 MLP, v1 = scalar MLP) inside a native patch step
mkdir -p gpurun_out
TAG=${1:-r02e}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdf_backward_mma -s 3 -c 1 -f -o gpurun_out/${TAG}_sdf_bwd_v2 \
    python scripts/profile_train_step.py --plain > gpurun_out/${TAG}_ncu_v2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_v2.log
