# ncu --set full of one sub-block launch of the hybrid decoder at N = 2^16 (same kernel and occupancy as the N = 2^20 batch):
# bash scripts/prof_hybrid_sub.sh -> gpurun_out/prof_hysub_<TAG>.ncu-rep
export PC_BENCH_LARGE_N=16 PC_SC_HYBRID=1
CMD="python bench.py --workload sc2p20 --frames 32768 --e2e-frames 64 --cpu-frames 8 --steps 1 --warmup 1 --no-secondary"
$CMD > gpurun_out/plain_hysub.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:${KERN:-sc_decode8_kernel} -s ${SKIP:-40} -c 1 -o gpurun_out/prof_hysub_${TAG:-a} -f $CMD > gpurun_out/ncu_hysub.log 2>&1
tail -c 600 gpurun_out/plain_hysub.log | head -c 400
