#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_v7.json 2> gpurun_out/r02_bench_v7.err; tail -c 300 gpurun_out/r02_bench_v7.json; tail -3 gpurun_out/r02_bench_v7.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_v7_reference.json 2> gpurun_out/r02_bench_v7_reference.err; cat gpurun_out/r02_bench_v7_reference.json | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
