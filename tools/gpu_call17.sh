#!/bin/bash
mkdir -p gpurun_out
export PRN_DP_BIG_ONLY=1
echo "--- 2 procs, no collective"; PRN_DP_OP=skip timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/dp_check.py after 2>&1 | grep -E "DP_CHECK|Error"
echo "--- 1 proc under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29520 tools/dp_check.py after 2>&1 | grep -E "DP_CHECK|Error"
echo "--- 2 procs, sum"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/dp_check.py after 2>&1 | grep -E "DP_CHECK|Error"
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
