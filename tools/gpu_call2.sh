#!/bin/bash
# round-2 GPU call 2: parity of the TMA conv kernel (conv cases, model tests), probe old vs new
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -q 2>&1 | tail -25
timeout 600 python tools/conv_probe.py > gpurun_out/r02_conv_probe_tma_v1.txt 2>&1; tail -40 gpurun_out/r02_conv_probe_tma_v1.txt
PRN_CONV_TMA=0 timeout 600 python tools/conv_probe.py l0_3x3_64 l0_1x1_64_256 l2_3x3_256 l2_1x1_256_1024 fpn0_3x3_256 mask0_3x3_256_128 deconv4_up_256_64 > gpurun_out/r02_conv_probe_old.txt 2>&1; tail -10 gpurun_out/r02_conv_probe_old.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_bench_shapes_gpu.py tests/test_loss_kernels_gpu.py -x -q 2>&1 | tail -15
timeout 300 python baseline/ref_runner.py --device cuda --mode train_loss --steps 5 --warmup 2 2>&1 | tail -12
