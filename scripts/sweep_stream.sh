# usage: bash scripts/sweep_stream.sh "<threads> <lsm> <frames>" ...
for cfg in "$@"; do set -- $cfg
PC_STREAM_THREADS=$1 PC_STREAM_LSM=$2 python bench.py --workload sc2p20 --frames $3 --e2e-frames 64 --cpu-frames 8 --steps 2 --warmup 1 --no-secondary 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads $1 lsm $2 frames $3', round(d['frames_per_s']), d['value'], d['ms_per_step'])"
done
