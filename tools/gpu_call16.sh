#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/dp_check.py after nccl_overlap > gpurun_out/r02_dp_check_2gpu_v3.txt 2>&1; grep -E "DP_CHECK|NCCL all|Error|error" gpurun_out/r02_dp_check_2gpu_v3.txt | head -20
