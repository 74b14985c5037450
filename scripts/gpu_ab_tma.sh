#!/bin/bash
mkdir -p gpurun_out
export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
for rep in 1 2; do
for tma in 1 0; do
  MANET_SH_DW_TMA=$tma timeout -s KILL 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tma$tma.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_tma$tma.log').read().strip().splitlines()[-1])
print('TMA=$tma seghead ms %.4f  prop ms/frame %.4f  fps %.1f' % (d['seghead']['ms'], d['propagation_50']['device_ms_per_frame'], d['propagation_50']['frames_per_s']))
PY
done
done
