CMD="python bench.py --workload qsc2048 --construction ga --frames 65536 --e2e-frames 4096 --cpu-frames 64 --steps 1 --warmup 1 --no-secondary"
$CMD > gpurun_out/plain_qsc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qsc_decode_kernel -s 1 -c 1 -o gpurun_out/prof_qsc_${TAG:-a} -f $CMD > gpurun_out/ncu_qsc.log 2>&1
tail -c 200 gpurun_out/plain_qsc.log
