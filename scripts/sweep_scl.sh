# usage: bash scripts/sweep_scl.sh "<lsm> <threads> <rgl> <ctas_per_sm>" ...   (frame-per-CTA SCL kernel tuning sweep, bench workload scl4096)
for cfg in "$@"; do set -- $cfg
PC_SCL_LSM=$1 PC_SCL_THREADS=$2 PC_SCL_RGL=$3 PC_SCL_CTAS_PER_SM=${4:-0} python bench.py --construction ga --frames ${FRAMES:-16384} --e2e-frames 1024 --cpu-frames 64 --steps 2 --warmup 1 --no-secondary 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lsm $1 threads $2 rgl $3 ctas/sm ${4:-0}', round(d['frames_per_s']), d['value'])"
done
