#!/bin/bash
# round-2 GPU call 1: hardware probe, assembled loss check, GPU reference timings, baseline GPU test suite
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 ./tools/probe/umma_probe > gpurun_out/r02_probe.txt 2>&1; echo "probe rc=$?"
tail -60 gpurun_out/r02_probe.txt
timeout 300 python tools/loss_module_check.py > gpurun_out/r02_loss_check.txt 2>&1; echo "loss rc=$?"
tail -40 gpurun_out/r02_loss_check.txt
for mode in fwd_dense fwd_e2e train_cot train_loss; do
  timeout 300 python baseline/ref_runner.py --device cuda --mode $mode --steps 10 --warmup 3 2>/dev/null | tail -1 | tee -a gpurun_out/r02_gpu_reference.jsonl
done
for mode in fwd_dense train_cot; do
  timeout 300 python baseline/ref_runner.py --device cuda --mode $mode --steps 10 --warmup 3 --amp --channels-last 2>/dev/null | tail -1 | tee -a gpurun_out/r02_gpu_reference.jsonl
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
