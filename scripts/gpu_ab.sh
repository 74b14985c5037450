#!/bin/bash
# A/B: single-CTA vs CTA-pair global matching kernel
mkdir -p gpurun_out
export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
timeout -s KILL 120 python scripts/umma_probe.py 120 214 120 6 2>&1 | tail -3
for mode in 0 1; do
  echo "== MANET_GM_CTA_PAIR=$mode"
  MANET_GM_CTA_PAIR=$mode timeout -s KILL 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$mode.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$mode.log').read().strip().splitlines()[-1]); r=d['roofline']
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'umma ms',round(r['kernel_ms'],4),'frac',round(r['frac'],4),'window',round(r['local']['window_kernel_ms'],4),'min',round(r['local']['min_kernel_ms'],4))
PY
done
timeout -s KILL 600 python -m pytest tests -m gpu -q -k "global or smoke or session" 2>&1 | tail -3
