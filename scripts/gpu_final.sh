#!/bin/bash
# Round-end validation on the GPU box: full GPU suite, smoke(), both bench arms.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r01j_pytest.log 2>&1
tail -4 gpurun_out/r01j_pytest.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r01j_bench.json 2> gpurun_out/r01j_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r01j_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])
s=d['sds_step']; print('sds', s.get('value'), s.get('ms_per_step'), s.get('nerf_side_only'), s.get('coarse_stage_nerf_side_only'), s.get('error'))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01j_bench_reference.json 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/r01j_bench_reference.json').read().strip().splitlines()[-1]); print('ref', d['value'], d['cpu_baseline']['cores'])"
timeout 300 python scripts/profile_train_step.py > gpurun_out/r01j_train_step_kernels.txt 2>&1
grep -E "sdf_backward|forward_sdf|render_tc|aten::mm |adam|Self CUDA time" gpurun_out/r01j_train_step_kernels.txt
