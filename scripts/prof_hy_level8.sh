#!/bin/bash
# ncu --set full of one hy_level8_kernel<0> launch (level 17 of a C4 step: the largest streamed byte-state level) and of one
# hy_level_sym8_kernel launch; plain run first
mkdir -p gpurun_out
CMD="python bench.py --workload sc2p20 --steps 1 --warmup 0 --e2e-frames 32 --cpu-frames 4 --no-secondary"
$CMD > gpurun_out/plain_c4b.log 2>&1 || { tail -3 gpurun_out/plain_c4b.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:hy_level8_kernel -s 0 -c 1 -o gpurun_out/prof_r2_hy8 -f $CMD > gpurun_out/ncu_r2_hy8.log 2>&1; tail -1 gpurun_out/ncu_r2_hy8.log
ncu --set full --clock-control none --import-source on -k regex:hy_level_sym8_kernel -s 0 -c 1 -o gpurun_out/prof_r2_sym8 -f $CMD > gpurun_out/ncu_r2_sym8.log 2>&1; tail -1 gpurun_out/ncu_r2_sym8.log
