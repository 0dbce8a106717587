#!/bin/bash
# sweep of (warps per SM, bulk-copy stages); args: "W:S" pairs
mkdir -p gpurun_out
out=gpurun_out/sweep_sclp3.log
: > $out
for ws in "$@"; do
  w=${ws%%:*}; s=${ws##*:}
  PC_SCLP_WARPS_PER_SM=$w PC_SCLP_STAGES=$s python scripts/sweep_sclp.py --mode probs >> $out 2>&1
done
grep -E "SWEEP|Error|error" $out
