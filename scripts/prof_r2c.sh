#!/bin/bash
# re-capture of sc_decode_kernel<symbols> (C1) after sc_binary.cu changed (the hybrid walk's symbol stage; the C1 kernel's code is the same)
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload sc1024 --frames 1048576 --e2e-frames 32768 --cpu-frames 2048 --steps 1 --warmup 1 --no-secondary"
$CMD > gpurun_out/plain_sc.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/plain_sc.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:sc_decode_kernel -s 1 -c 1 -o gpurun_out/prof_r2_sc -f $CMD > gpurun_out/ncu_r2_sc.log 2>&1
tail -1 gpurun_out/ncu_r2_sc.log
