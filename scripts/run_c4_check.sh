#!/bin/bash
# parity tests of the large-block paths + C4 bench lines: two levels per byte-state pass (default) vs one (PC_HY_FUSE2=0), and a
# batch of one full sub-block wave (37,888 frames)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sc_stream.py tests/test_gpu_sc_binary.py -x -q -m gpu > gpurun_out/pytest_c4_check.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_c4_check.log
one() {  # tag, env, extra args
  env $2 python bench.py --workload sc2p20 --no-secondary --steps 3 --warmup 1 --e2e-frames 32 --cpu-frames 4 $3 2>gpurun_out/bench_c4_$1.err > gpurun_out/bench_c4_$1.json
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_c4_$1.json") if x.startswith("{")]
j=json.loads(l[-1]); print("$1", j["value"], j["frames_per_s"], j["ms_per_step"], (j.get("parity_check") or {}).get("identical"))
PY
}
one fuse2 PC_HY_FUSE2=1 ""
one fuse1 PC_HY_FUSE2=0 ""
one wave PC_HY_FUSE2=1 "--frames 37888"
