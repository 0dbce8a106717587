#!/bin/bash
# A/B of the prune's arg-max: butterfly shuffles (SCLP_SELECT_SHFL=1, the .so in the tree) vs three redux.sync (=0, rebuilt here)
mkdir -p gpurun_out
out=gpurun_out/sweep_select.log
: > $out
for i in 1 2; do python scripts/sweep_sclp.py --mode sym >> $out 2>&1; done
echo "--- rebuild SCLP_SELECT_SHFL=0" >> $out
PC_NVCC_EXTRA="-DSCLP_SELECT_SHFL=0" python polarcub_b200/build.py -f >> $out 2>&1
for i in 1 2; do python scripts/sweep_sclp.py --mode sym >> $out 2>&1; done
grep -E "SWEEP|rebuild|rror" $out
