#!/bin/bash
# launch list of one C4 step (byte-state hybrid walk): which kernels the 0.5 s go to
mkdir -p gpurun_out
CMD="python bench.py --workload sc2p20 --steps 1 --warmup 1 --e2e-frames 32 --cpu-frames 4 --no-secondary"
$CMD > gpurun_out/plain_c4.log 2>&1 || { tail -5 gpurun_out/plain_c4.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hy_|sc_decode|egress|ingest|transform" -c 9000 --csv --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_c4.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_c4.csv
tail -c 600 gpurun_out/plain_c4.log | head -c 400
