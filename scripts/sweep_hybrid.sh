for fr in "$@"; do
PC_SC_HYBRID_FRAMES=$fr python bench.py --workload sc2p20 --frames $fr --e2e-frames 64 --cpu-frames 8 --steps 2 --warmup 1 --no-secondary 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('frames $fr', round(d['frames_per_s']), d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_check'])"
done
