#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_model_gpu.py tests/test_dropin_gpu.py tests/test_train_ops_gpu.py -q -x 2>&1 | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-inference > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_v4.json'))
print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"], "fwd/bwd", d["train_step"]["fwd_ms"], d["train_step"]["bwd_ms"], "full", d["train_step"]["full_iteration"])
print("e2e", d["e2e"])
P
tail -3 gpurun_out/r02_bench_v4.err
