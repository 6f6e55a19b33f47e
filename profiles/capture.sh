#!/bin/bash
# Round profile capture (one GPU): plain bench first, then the ncu launch list of the same command, then
# `ncu --set full` of the kernels of the two small drivers.  Outputs land in gpurun_out/.
set -x
T=${1:-r1}
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_ncu_launch.log 2>&1
python profiles/train_profile.py 3 > gpurun_out/${T}_train_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'assign_gt_kernel|loss_kernel' -s 4 -c 2 -f \
    -o gpurun_out/${T}_train_fwd python profiles/train_profile.py 3 > gpurun_out/${T}_ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'loss_kernel' -s 5 -c 1 -f \
    -o gpurun_out/${T}_train_grad python profiles/train_profile.py 3 >> gpurun_out/${T}_ncu_train.log 2>&1
python profiles/pp_profile.py d3 32 3 > gpurun_out/${T}_pp_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'topk_|detect_kernel' -s 10 -c 5 -f \
    -o gpurun_out/${T}_pp python profiles/pp_profile.py d3 32 3 > gpurun_out/${T}_ncu_pp.log 2>&1
python profiles/run_configs.py > gpurun_out/${T}_configs.log 2> gpurun_out/${T}_configs.err
tail -n 3 gpurun_out/${T}_train_plain.log gpurun_out/${T}_pp_plain.log gpurun_out/${T}_configs.log
