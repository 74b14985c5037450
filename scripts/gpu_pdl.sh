#!/bin/bash
# A/B of programmatic dependent launch: GPU tests with PDL on, then bench with and without
mkdir -p gpurun_out
export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in 1 0; do
  MANET_PDL=$v timeout -s KILL 300 python bench.py --steps 40 --warmup 5 | tail -1 > gpurun_out/bench_pdl$v.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pdl$v.log").read())
print("PDL=$v value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "single", round(d["single_stream"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "stream", round(d["e2e"]["streaming"]["value"],1))
PY
done
