export PC_SCL_THREADS=64 PC_SCL_LSM=4 PC_SCL_RGL=12
CMD="python bench.py --frames 8192 --e2e-frames 256 --cpu-frames 32 --steps 1 --warmup 1 --no-secondary"
$CMD > gpurun_out/plain_scl_t64.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scl2_kernel -s 1 -c 1 -o gpurun_out/prof_scl2_t64 -f $CMD > gpurun_out/ncu_scl_t64.log 2>&1
tail -c 600 gpurun_out/plain_scl_t64.log
