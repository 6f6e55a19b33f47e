#!/bin/bash
# assign_gt_kernel duration (ncu launch list of the training driver), anchors recomputed vs gathered:
#   gpurun -- bash profiles/assign_exp.sh label
for mode in 1 0; do
  ODK_TEST_ANCHOR_GEN=$mode ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/assign_$1_$mode.csv \
      python profiles/train_profile.py 3 transient > /dev/null 2>&1
  echo "anchor generator = $mode"; python profiles/launch_summary.py gpurun_out/assign_$1_$mode.csv assign_gt
done
