#!/bin/bash
# knob sweep of the path-per-lane list decoder (scl_path.cu); one process per setting
mkdir -p gpurun_out
out=gpurun_out/sweep_sclp.log
: > $out
for w in 8 12 16 20 24 32; do
  PC_SCLP_WARPS_PER_SM=$w python scripts/sweep_sclp.py --mode probs >> $out 2>&1
done
for w in 16 24; do
  PC_SCLP_WARPS_PER_SM=$w python scripts/sweep_sclp.py --mode sym >> $out 2>&1
done
PC_SCLP_WARPS_PER_SM=20 PC_SCLP_NOFUSE=1 python scripts/sweep_sclp.py --mode probs >> $out 2>&1
PC_SCLP_WARPS_PER_SM=16 PC_SCLP_LSM=2 python scripts/sweep_sclp.py --mode probs >> $out 2>&1
PC_SCLP_WARPS_PER_SM=24 PC_SCLP_LSM=3 python scripts/sweep_sclp.py --mode probs >> $out 2>&1
PC_SCLP_WARPS_PER_SM=20 PC_SCLP_RGL=9 python scripts/sweep_sclp.py --mode probs >> $out 2>&1
PC_SCLP_WARPS_PER_SM=20 python scripts/sweep_sclp.py --mode probs --ebn0 1.0 >> $out 2>&1
grep SWEEP $out
