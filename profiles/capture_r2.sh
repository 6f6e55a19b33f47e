#!/bin/bash
# Round-2 profile capture (one GPU): plain bench first, then the ncu launch list of the same command, then
# `ncu --set full` of the kernels of the two small drivers (each only after the same driver exited 0 without ncu).
# Outputs land in gpurun_out/; profiles/make_summary_r2.py turns them into profiles/r2_summary.md.
set -x
T=${1:-r2}
NCU="ncu --set full --clock-control none --import-source on -f"
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_ncu_launch.log 2>&1
# training step (D0 B=64): third forward iteration, third forward+gradient iteration
python profiles/train_profile.py 3 transient > gpurun_out/${T}_train_plain.log 2>&1 || exit 1
$NCU -k regex:'assign_gt_kernel|loss_flat|loss_patch_kernel' -s 6 -c 3 -o gpurun_out/${T}_train_fwd \
    python profiles/train_profile.py 3 transient > gpurun_out/${T}_ncu_train.log 2>&1
$NCU -k regex:'assign_gt_kernel|loss_flat|loss_patch_kernel' -s 15 -c 3 -o gpurun_out/${T}_train_grad \
    python profiles/train_profile.py 3 transient >> gpurun_out/${T}_ncu_train.log 2>&1
# post-process (D3 B=32): sample + collect + tail of the fused entry point, hard then soft suppression
ODK_SHOW_TIMELINE=1 python profiles/pp_fused_profile.py d3 32 10 dense,planted staged > gpurun_out/${T}_pp_plain.log 2>&1 || exit 1
ODK_SHOW_TIMELINE=0 $NCU -k regex:'sample_kernel|topk_collect_kernel|post_tail_kernel' -s 59 -c 3 -o gpurun_out/${T}_pp_hard \
    python profiles/pp_fused_profile.py d3 32 2 dense staged > gpurun_out/${T}_ncu_pp.log 2>&1
ODK_SHOW_TIMELINE=0 ODK_PP_SOFT=1 $NCU -k regex:'post_tail_kernel' -s 4 -c 1 -o gpurun_out/${T}_pp_soft \
    python profiles/pp_fused_profile.py d3 32 2 dense staged >> gpurun_out/${T}_ncu_pp.log 2>&1
ls -la gpurun_out/ | tail -20
tail -n 4 gpurun_out/${T}_train_plain.log
grep "odk_postprocess" gpurun_out/${T}_pp_plain.log
