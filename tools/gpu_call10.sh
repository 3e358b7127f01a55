#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/dp_check.py > gpurun_out/r02_dp_check_2gpu.txt 2>&1; grep -E "DP_CHECK|Error|error|assert" gpurun_out/r02_dp_check_2gpu.txt | head -20; tail -5 gpurun_out/r02_dp_check_2gpu.txt | cut -c1-300
