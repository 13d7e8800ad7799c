#!/bin/bash
# round 2: the fused shade kernels + native patch step (training path) -- tests, per-kernel timing, ncu launch list of a step
mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 900 python -m pytest tests/test_gpu_training.py -x -q -s > gpurun_out/${TAG}_pytest_train.log 2>&1
tail -30 gpurun_out/${TAG}_pytest_train.log
timeout 300 python scripts/profile_train_step.py > gpurun_out/${TAG}_train_step_kernels.txt 2>&1
tail -40 gpurun_out/${TAG}_train_step_kernels.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_step_launches.csv \
    python scripts/profile_train_step.py --plain > gpurun_out/${TAG}_train_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_train_step_launches.csv")) if len(r) > 5 and r[0].isdigit()]
print(len(rows), "launches")
agg = collections.OrderedDict()
for r in rows[-200:]:
    pass
PY
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_all.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_all.log
