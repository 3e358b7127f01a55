#!/bin/bash
timeout 900 python -m pytest tests/test_train_ops_gpu.py tests/test_train_model_gpu.py tests/test_boundary_cpu.py -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-inference > gpurun_out/r02_bench_v8.json 2>gpurun_out/r02_bench_v8.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_v8.json').read().strip().splitlines()[-1]); print('value', d['value'], d['ms_per_step'], d['train_step']['fwd_ms'], d['train_step']['bwd_ms'], 'launches/step', d['train_step']['gpu_launches_per_step'])
print('e2e', d['e2e'].get('value'), d['e2e'].get('ms_per_step'), d['e2e'].get('device_sampling'), d['e2e'].get('error'))
P
tail -3 gpurun_out/r02_bench_v8.err
