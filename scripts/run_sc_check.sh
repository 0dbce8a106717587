#!/bin/bash
# parity tests of the binary SC paths + the C1 bench line with the block schedule (default) and the leaf-by-leaf one (PC_SC_BLOCK=0)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sc_binary.py tests/test_gpu_genie.py tests/test_gpu_sc_stream.py tests/test_gpu_host_pipeline.py tests/test_gpu_trellis.py -x -q -m gpu > gpurun_out/pytest_sc_check.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_sc_check.log
for b in 1 0; do
PC_SC_BLOCK=$b python bench.py --workload sc1024 --no-secondary --steps 10 --warmup 3 2>gpurun_out/bench_sc_check.err > gpurun_out/bench_sc_check_$b.json
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_sc_check_$b.json") if x.startswith("{")]
j=json.loads(l[-1]); print("block=$b", j["value"], j["e2e"]["value"], j.get("parity_check"), j["roofline"]["achieved"])
PY
done
