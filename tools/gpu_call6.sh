#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_loss_kernels_gpu.py tests/test_postprocess_gpu.py tests/test_dropin_gpu.py tests/test_bench_shapes_gpu.py::test_train_step_480x640_bs8_graph_equals_eager -q -x -s 2>&1 | grep -v "^$" | tail -15
timeout 600 python tools/loss_profile.py 8 > gpurun_out/r02_loss_profile_v2.txt 2>&1; tail -8 gpurun_out/r02_loss_profile_v2.txt
# launch lists (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6200 -c 1800 --csv --log-file gpurun_out/r02_train_launches_v1.csv python bench.py --steps 2 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-inference > gpurun_out/r02_ncu_train.log 2>&1; tail -2 gpurun_out/r02_ncu_train.log | cut -c1-300
python tools/summarize_launches.py gpurun_out/r02_train_launches_v1.csv > gpurun_out/r02_train_launches_v1_summary.txt 2>&1; head -24 gpurun_out/r02_train_launches_v1_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v1.csv python tools/profile_step.py > gpurun_out/r02_ncu_fwd.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_v1.csv > gpurun_out/r02_launches_v1_summary.txt 2>&1; head -20 gpurun_out/r02_launches_v1_summary.txt
# full captures of the conv kernel on the layers that carry the step
for L in fpn0_3x3_256 deconv4_subpixel_256_4x64 mask0_3x3_256_128 l0_1x1_64_256 l2_3x3_256 l2_1x1_256_1024; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tma -s 3 -c 1 -f -o gpurun_out/r02_full_$L python tools/conv_probe.py $L > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/r02_full_dcn_l2_256 python tools/conv_probe.py dcn_l2_256 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_v3.json'))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "fwd/bwd", d["train_step"]["fwd_ms"], d["train_step"]["bwd_ms"])
print("inference", d["inference"]["value"], d["inference"]["ms_per_step"], "e2e", d["inference"]["e2e"]["value"])
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"])
print("gpu_reference", {k:(v.get("value") if isinstance(v,dict) else None) for k,v in d["gpu_reference"].items()})
print("cpu", d["cpu_baseline"])
P
tail -3 gpurun_out/r02_bench_v3.err
