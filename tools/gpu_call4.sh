#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_shapes_gpu.py::test_train_step_480x640_bs8_graph_equals_eager tests/test_dropin_gpu.py tests/test_pointwise_gpu.py -q -x -s 2>&1 | tail -40
S="l0_1x1_64_256 l0_1x1_256_64 l2_1x1_256_1024 l2_3x3_256 fpn0_3x3_256"
timeout 600 python tools/conv_probe.py $S > gpurun_out/r02_conv_probe_tma_v3.txt 2>&1; cat gpurun_out/r02_conv_probe_tma_v3.txt
timeout 600 python tools/loss_profile.py 8 > gpurun_out/r02_loss_profile_v1.txt 2>&1; tail -12 gpurun_out/r02_loss_profile_v1.txt
