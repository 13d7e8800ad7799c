#!/bin/bash
# round 2: ncu --set full of the stencil forward with the packed (sample, direction) rounds
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"forward_sdf_tc" -s 3 -c 1 -f -o gpurun_out/r02ai_stencil_forward \
    python scripts/profile_train_step.py --plain > gpurun_out/r02ai_ncu.log 2>&1
tail -2 gpurun_out/r02ai_ncu.log
