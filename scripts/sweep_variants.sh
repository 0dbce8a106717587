#!/bin/bash
# A/B of variant libraries variants/<name>.so against the .so in the tree: bench.py --workload $1 for each; $2.. = variant names
mkdir -p gpurun_out
w=$1; shift
cp polarcub_b200/libpolarcub_b200.so /tmp/keep.so
for v in tree "$@"; do
  if [ $v != tree ]; then cp variants/$v.so polarcub_b200/libpolarcub_b200.so; fi
  python bench.py --workload $w --no-secondary --steps 10 --warmup 3 2>gpurun_out/bench_var_$v.err > gpurun_out/bench_var_$v.json
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_var_$v.json") if x.startswith("{")]
j=json.loads(l[-1]); print("$v", j["value"], j["e2e"]["value"], (j.get("parity_check") or {}).get("identical"), j["roofline"]["achieved"])
PY
done
cp /tmp/keep.so polarcub_b200/libpolarcub_b200.so
