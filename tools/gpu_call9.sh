#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-inference > gpurun_out/r02_bench_2gpu_v1.json 2> gpurun_out/r02_bench_2gpu_v1.err; python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_2gpu_v1.json').read().strip().splitlines()[-1])
    print("2gpu value", d["value"], "ms", d["ms_per_step"], "fwd/bwd", d["train_step"]["fwd_ms"], d["train_step"]["bwd_ms"], "full", d["train_step"]["full_iteration"])
    print("allreduce", d["train_step"]["grad_allreduce"])
    print("e2e", d["e2e"].get("value"), d["e2e"].get("error"))
except Exception as e:
    print("parse failed", e)
P
tail -5 gpurun_out/r02_bench_2gpu_v1.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-inference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu value', d['value'], d['ms_per_step'])"
