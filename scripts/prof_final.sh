#!/bin/bash
# last evidence pass of the round: ncu --set full of hy_level8_bulk_kernel and hy_level_sym8_bulk_kernel (C4) and of sc_decode_kernel<symbols> (C1, sc_binary.cu
# changed again), then the bench lines of C4, C1 and the default workload
mkdir -p gpurun_out
CMD="python bench.py --workload sc2p20 --steps 1 --warmup 0 --e2e-frames 32 --cpu-frames 4 --no-secondary"
ncu --set full --clock-control none --import-source on -k regex:hy_level8_bulk_kernel -s 0 -c 1 -o gpurun_out/prof_r2_hy8bulk -f $CMD > gpurun_out/ncu_r2_hy8bulk.log 2>&1; tail -1 gpurun_out/ncu_r2_hy8bulk.log
ncu --set full --clock-control none --import-source on -k regex:hy_level_sym8_bulk_kernel -s 0 -c 1 -o gpurun_out/prof_r2_sym8bulk -f $CMD > gpurun_out/ncu_r2_sym8bulk.log 2>&1; tail -1 gpurun_out/ncu_r2_sym8bulk.log
bash scripts/prof_r2c.sh
python bench.py --workload sc2p20 --no-secondary > gpurun_out/bench_r2_sc2p20.json 2> gpurun_out/bench_r2_sc2p20.err
python bench.py --workload sc1024 --no-secondary > gpurun_out/bench_r2_sc1024.json 2> gpurun_out/bench_r2_sc1024.err
