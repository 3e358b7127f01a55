#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./tools/probe/umma_probe > gpurun_out/r02_probe_v2.txt 2>&1; grep T5 gpurun_out/r02_probe_v2.txt
S="l0_1x1_64_256 l0_1x1_256_64 l1_1x1_128_512 l2_1x1_256_1024 l2_1x1_1024_256 l2_3x3_256 mask0_3x3_256_128 ppa_dyn stem_k192_64"
timeout 600 python tools/conv_probe.py $S > gpurun_out/r02_conv_probe_tma_v2.txt 2>&1; cat gpurun_out/r02_conv_probe_tma_v2.txt
PRN_TMA_STORE=0 timeout 600 python tools/conv_probe.py $S > gpurun_out/r02_conv_probe_tma_v2_direct.txt 2>&1; cat gpurun_out/r02_conv_probe_tma_v2_direct.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err; tail -c 6000 gpurun_out/r02_bench_v1.json; tail -5 gpurun_out/r02_bench_v1.err
