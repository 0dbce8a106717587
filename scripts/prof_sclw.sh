CMD="python bench.py --frames 8192 --e2e-frames 256 --cpu-frames 32 --steps 1 --warmup 1 --no-secondary"
$CMD > gpurun_out/plain_sclw.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sclw_kernel -s 1 -c 1 -o gpurun_out/prof_sclw_${TAG:-a} -f $CMD > gpurun_out/ncu_sclw.log 2>&1
tail -c 300 gpurun_out/plain_sclw.log
