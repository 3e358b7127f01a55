#!/bin/bash
# ncu launch lists (gpu__time_duration.sum, --clock-control none) of the final code: eval forward (eager) and one replayed training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v2.csv \
    python tools/profile_step.py PlaneRecNet_101_config 8 f16 > gpurun_out/ncu_fwd_v2.log 2>&1
tail -1 gpurun_out/ncu_fwd_v2.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --graph-profiling node --csv \
    --log-file gpurun_out/r02_train_launches_v2.csv python tools/profile_train_step.py > gpurun_out/ncu_train_v2.log 2>&1
tail -1 gpurun_out/ncu_train_v2.log
wc -l gpurun_out/r02_launches_v2.csv gpurun_out/r02_train_launches_v2.csv
