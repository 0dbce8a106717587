# usage: bash scripts/sweep_sclw.sh "<warps_per_sm> <rgl> [lsm]" ...   (frame-per-warp SCL kernel tuning sweep, bench workload scl4096)
for cfg in "$@"; do set -- $cfg
PC_SCLW_WARPS_PER_SM=$1 PC_SCLW_RGL=$2 PC_SCLW_LSM=${3:--1} python bench.py --frames ${FRAMES:-32768} --e2e-frames 1024 --cpu-frames 64 --steps 2 --warmup 1 --no-secondary 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warps/sm $1 rgl $2 lsm ${3:--1}', round(d['frames_per_s']), d['value'])"
done
