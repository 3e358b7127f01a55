#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench_v6.json 2> gpurun_out/r02_bench_v6.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_v6.json'))
print("MAXCONN=32 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), "fwd/bwd", d["train_step"]["fwd_ms"], d["train_step"]["bwd_ms"])
print("inference", d["inference"]["value"], d["inference"]["ms_per_step"], "e2e", d["inference"]["e2e"]["value"])
P
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench_v6b.json 2> gpurun_out/r02_bench_v6b.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_v6b.json'))
print("MAXCONN=8 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), "fwd/bwd", d["train_step"]["fwd_ms"], d["train_step"]["bwd_ms"])
print("inference", d["inference"]["value"], d["inference"]["ms_per_step"], "e2e", d["inference"]["e2e"]["value"])
P
