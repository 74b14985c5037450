#!/bin/bash
# A/B of the global-matching kernel variants: 0 single CTA, 1 multicast pair, 2 cta_group::2 pair
mkdir -p gpurun_out
export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
for mode in ${VARIANTS:-0 1 2}; do
  echo "== MANET_GM_VARIANT=$mode"
  MANET_GM_VARIANT=$mode timeout -s KILL 120 python scripts/umma_probe.py 120 214 120 6 2>&1 | tail -2
  MANET_GM_VARIANT=$mode timeout -s KILL 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$mode.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$mode.log').read().strip().splitlines()[-1]); r=d['roofline']
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'umma ms',round(r['kernel_ms'],4),'frac',round(r['frac'],4),'e2e',round(d['e2e']['value'],1))
PY
done
