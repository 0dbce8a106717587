CMD="python bench.py --workload sc1024 --frames 1048576 --e2e-frames 32768 --cpu-frames 2048 --steps 1 --warmup 1 --no-secondary"
$CMD > gpurun_out/plain_sc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sc_decode_kernel -s 1 -c 1 -o gpurun_out/prof_sc_${TAG:-b} -f $CMD > gpurun_out/ncu_sc.log 2>&1
tail -c 1500 gpurun_out/plain_sc.log | head -c 900
