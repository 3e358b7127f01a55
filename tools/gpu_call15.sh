#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/dp_check.py after p2p nccl_overlap > gpurun_out/r02_dp_check_2gpu_v2.txt 2>&1; grep -E "DP_CHECK|Error|error" gpurun_out/r02_dp_check_2gpu_v2.txt | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-inference 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu value', d['value'], d['ms_per_step'])"
