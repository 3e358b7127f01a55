#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_ops_gpu.py tests/test_train_model_gpu.py -q -x 2>&1 | tail -6
timeout 300 python tools/wgrad_probe.py > gpurun_out/r02_wgrad_probe_tma.txt 2>&1; cut -c1-160 gpurun_out/r02_wgrad_probe_tma.txt
PRN_WGRAD_TMA=0 timeout 300 python tools/wgrad_probe.py l0_1x1_64_256 l2_1x1_256_1024 l2_1x1_1024_256 l3_1x1_2048_512 > gpurun_out/r02_wgrad_probe_old.txt 2>&1; cut -c1-160 gpurun_out/r02_wgrad_probe_old.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-inference > gpurun_out/r02_bench_v5.json 2> gpurun_out/r02_bench_v5.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_v5.json'))
print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"], "fwd/bwd", d["train_step"]["fwd_ms"], d["train_step"]["bwd_ms"])
P
tail -3 gpurun_out/r02_bench_v5.err
