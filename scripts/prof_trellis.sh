# ncu --set full of one trellis_step_kernel launch of the deletion-channel workload: bash scripts/prof_trellis.sh
CMD="python bench.py --workload del256 --steps 1 --warmup 1 --cpu-frames 64 --no-secondary"
$CMD > gpurun_out/plain_trellis.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:trellis_step_kernel -s ${SKIP:-8} -c 2 -o gpurun_out/prof_trellis_${TAG:-a} -f $CMD > gpurun_out/ncu_trellis.log 2>&1
tail -c 300 gpurun_out/plain_trellis.log | head -c 200
