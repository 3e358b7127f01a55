#!/bin/bash
# ncu --set full captures of the training-side kernels inside the replayed step (a few launches of each, spread over the step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ncu_train   # the .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB): only the summary travels
for spec in "chan_reduce:20:5:5" "bn_bwd_apply:15:5:5" "bn_finalize_apply:15:5:5" "wgrad_umma:10:6:6" "dcn_col2im_bwd:2:3:1" "dcn_im2col:2:3:1" \
            "gn_bwd_reduce:4:3:2" "gn_bwd_apply:4:3:2" "maxpool3s2_bwd:0:1:1" "reflect_fold:2:3:1" "pack_multi:0:1:1" "copy_multi:0:1:1"; do
  IFS=: read name skip count stride <<< "$spec"
  timeout 240 ncu --set full --clock-control none --profile-from-start off --graph-profiling node -k regex:$name -s $skip -c $count \
      -f -o /tmp/ncu_train/r02_full_train_$name python tools/profile_train_step.py > gpurun_out/ncu_train_$name.log 2>&1
  tail -1 gpurun_out/ncu_train_$name.log | cut -c1-120
done
python tools/ncu_summary_multi.py /tmp/ncu_train/r02_full_train_*.ncu-rep > gpurun_out/r02_ncu_train_kernels_summary.txt
cat gpurun_out/r02_ncu_train_kernels_summary.txt | cut -c1-170
cp /tmp/ncu_train/r02_full_train_chan_reduce.ncu-rep /tmp/ncu_train/r02_full_train_wgrad_umma.ncu-rep gpurun_out/ 2>/dev/null
