#!/bin/bash
# sweep of (warps per SM, stages, lsm, rgl); args: "W:S:LSM:RGL"
mkdir -p gpurun_out
out=gpurun_out/sweep_sclp5.log
: > $out
for a in "$@"; do
  IFS=: read w s l r <<< "$a"
  PC_SCLP_WARPS_PER_SM=$w PC_SCLP_STAGES=$s PC_SCLP_LSM=$l PC_SCLP_RGL=$r python scripts/sweep_sclp.py --mode ${MODE:-sym} >> $out 2>&1
done
grep -E "SWEEP|Error|error" $out
