#!/bin/bash
# parity tests of the large-block paths + C4 bench lines: byte-state level kernel with bulk-copy staging (default) vs plain loads (PC_HY_BULK=0)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sc_stream.py -x -q -m gpu > gpurun_out/pytest_c4_check.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_c4_check.log
one() {  # tag, env
  env $2 python bench.py --workload sc2p20 --no-secondary --steps 3 --warmup 1 --e2e-frames 32 --cpu-frames 4 2>gpurun_out/bench_c4_$1.err > gpurun_out/bench_c4_$1.json
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_c4_$1.json") if x.startswith("{")]
j=json.loads(l[-1]); print("$1", j["value"], j["frames_per_s"], j["ms_per_step"], (j.get("parity_check") or {}).get("identical"))
PY
}
one bulk PC_HY_BULK=1
one plain PC_HY_BULK=0
