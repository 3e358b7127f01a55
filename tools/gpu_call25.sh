#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_probe.py > gpurun_out/r02_conv_probe_tma_v5.txt 2>&1; cut -c1-100 gpurun_out/r02_conv_probe_tma_v5.txt
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench_v9.json 2> gpurun_out/r02_bench_v9.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_v9.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"].get("value"), d["e2e"].get("device_sampling",{}).get("value"), "fwd/bwd", d["train_step"]["fwd_ms"], d["train_step"]["bwd_ms"])
print("inference", d["inference"]["value"], d["inference"]["ms_per_step"], "e2e", d["inference"]["e2e"]["value"])
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], json.dumps(d["roofline"]["by_kind"]))
P
tail -3 gpurun_out/r02_bench_v9.err
