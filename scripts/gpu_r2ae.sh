#!/bin/bash
# round 2: ncu --set full of the scatter kernel after the v4 reductions + adaptive scan
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sdf_scatter" -s 3 -c 1 -f -o gpurun_out/r02ae_scatter \
    python scripts/profile_train_step.py --plain > gpurun_out/r02ae_ncu.log 2>&1
tail -2 gpurun_out/r02ae_ncu.log; ls -la gpurun_out | grep r02ae
