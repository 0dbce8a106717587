#!/bin/bash
# all bench lines of the round on one GPU: gpurun_out/bench_r2_<workload>.json (+ the reference arm of the default workload)
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err
for w in sc1024 qsc2048 sc2p20 del256; do python bench.py --workload $w --no-secondary > gpurun_out/bench_r2_$w.json 2> gpurun_out/bench_r2_$w.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_r2_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("bench_r2_")[1], d.get("config",{}).get("workload"), "value", round(d.get("value",0),4), d.get("unit"), "e2e", round(d.get("e2e",{}).get("value",0),4), "fps", d.get("frames_per_s"), "parity", (d.get("parity_check") or {}).get("identical"), "/", (d.get("parity_check") or {}).get("frames_compared"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
P
