CMD="python bench.py --workload sc1024 --frames 8192 --e2e-frames 8192 --cpu-frames 256 --steps 1 --warmup 1 --no-secondary"
$CMD > gpurun_out/plain_sc_small.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sc_decode_kernel -s 1 -c 1 -o gpurun_out/prof_sc_small -f $CMD > gpurun_out/ncu_sc_small.log 2>&1
python -c "
import json; d=json.loads(open('gpurun_out/plain_sc_small.log').read().strip().splitlines()[-1]); print(d['roofline']['kernel_ms_per_launch'], d['frames_per_s'])"
