#!/bin/bash
# short sweep: warps per SM in "$@" (default 16 20 24), probability-pair input
mkdir -p gpurun_out
out=gpurun_out/sweep_sclp2.log
: > $out
for w in ${@:-16 20 24}; do
  PC_SCLP_WARPS_PER_SM=$w python scripts/sweep_sclp.py --mode probs >> $out 2>&1
done
grep -E "SWEEP|Error|error" $out
