#!/bin/bash
# round 2: 2-GPU driver-style bench after the UNet graph replay and the scatter / warp changes (+ the NCCL gradient-equality test)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02ad_pytest_multi.log 2>&1; tail -2 gpurun_out/r02ad_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/r02ad_bench_2gpu.json 2> gpurun_out/r02ad_bench_2gpu.err
grep -i "warn.*graph\|capture" gpurun_out/r02ad_bench_2gpu.err | head -5
python - <<PY
import json
line=[l for l in open("gpurun_out/r02ad_bench_2gpu.json") if l.startswith("{")][-1]
d=json.loads(line)
print(d["value"], d["ms_per_step"], d["n_gpus"])
s=d["sds_step"]; print(s.get("value"), s.get("ms_per_step"), s.get("phases_ms"), s.get("error"))
print(s["nerf_side_only"]["ms_per_step"], s["coarse_stage_nerf_side_only"]["ms_per_step"], s["config5_512_multibbox_sd21"].get("ms_per_step"))
print(d["warp_frame"].get("ms_per_frame"), d["warp_frame"].get("value"))
PY
