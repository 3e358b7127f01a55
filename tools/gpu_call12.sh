#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointwise_gpu.py -x -q 2>&1 | grep -E "Error|assert|rel_l2|^E" | head -20
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_post_launches.csv python tools/profile_post.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_post_launches.csv | head -30
PRN_CONV_TMA=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_post_launches_old.csv python tools/profile_post.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_post_launches_old.csv | head -12
