#!/bin/bash
mkdir -p gpurun_out
for v in 0 2; do
  echo "== full GPU tests with MANET_GM_VARIANT=$v"
  MANET_GM_VARIANT=$v timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
done
export MANET_BENCH_CPU=0
for v in 0 1 2; do
  MANET_GM_VARIANT=$v MANET_BENCH_SHARDED=1 timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s$v.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s$v.log').read().strip().splitlines()[-1])
print('variant $v: 480p umma ms', round(d['roofline']['kernel_ms'],4), ' 1080p sharded leg:', d['sharded_global_1080p']['ms'], 'ms', round(d['sharded_global_1080p']['algorithmic_tflops'],1),'TF/s')
PY
done
