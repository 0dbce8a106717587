#!/bin/bash
# round-2 evidence pass for the kernels that changed late in the round (one GPU): ncu --set full of sclp_kernel (C2) and
# qsc_decode_kernel<3> (C3), each only after the same command has exited 0 without ncu, and the launch list of the default
# bench command.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
run() {  # tag, kernel regex, skip, count, command...
  local tag=$1 k=$2 skip=$3 cnt=$4; shift 4
  "$@" > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run of $tag failed"; tail -3 gpurun_out/plain_$tag.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -o gpurun_out/prof_r2_$tag -f "$@" > gpurun_out/ncu_r2_$tag.log 2>&1
  tail -1 gpurun_out/ncu_r2_$tag.log
}
run sclp sclp_kernel 1 1 python scripts/sweep_sclp.py --mode sym --frames 9472 --steps 1
run qsc qsc_decode_kernel 1 1 python bench.py --workload qsc2048 --frames 151552 --e2e-frames 4096 --cpu-frames 64 --steps 1 --warmup 1 --no-secondary
CMD="python bench.py --steps 2 --warmup 1 --no-secondary"
$CMD > gpurun_out/plain_r2_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_default.csv $CMD > gpurun_out/ncu_ll_r2_default.log 2>&1
tail -c 300 gpurun_out/plain_r2_default.log | head -c 200; echo
