#!/bin/bash
# A/B of the fused sweeps of qsc_decode_kernel: QSC_FUSE_DEPTH 2 (the .so in the tree) vs 1 (variant library built beside it)
mkdir -p gpurun_out
out=gpurun_out/sweep_qsc_fuse.log
: > $out
python -m pytest tests/test_gpu_sc_qary.py tests/test_gpu_log.py tests/test_gpu_sim_qary.py -x -q -m gpu > gpurun_out/pytest_qsc_fuse.log 2>&1; echo "pytest rc=$?" >> $out
for d in 2 1; do
  if [ $d = 1 ]; then cp polarcub_b200/libpolarcub_b200.so /tmp/keep.so; cp gpurun_in_qfd1.so polarcub_b200/libpolarcub_b200.so; fi
  echo "--- QSC_FUSE_DEPTH=$d" >> $out
  python bench.py --workload qsc2048 --no-secondary --steps 10 --warmup 3 2>>$out | tee gpurun_out/bench_qsc_fuse_$d.json >> $out
done
cp /tmp/keep.so polarcub_b200/libpolarcub_b200.so
grep -E "rc=|---" $out; for d in 2 1; do python - <<PY
import json
l=[x for x in open("gpurun_out/bench_qsc_fuse_$d.json") if x.startswith("{")]
j=json.loads(l[-1]); print($d, j["value"], j["e2e"]["value"], j.get("parity_check"), j["roofline"]["achieved"])
PY
done
