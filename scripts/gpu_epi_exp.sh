#!/bin/bash
# timing experiment: epilogue reading 1/4, 2/4, 4/4 of the accumulator columns (results wrong for <4)
export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
cp cvpr2020_manet_b200/lib/libmanet_b200.so /tmp/lib_full.so
for n in 1 2 4; do
  if [ $n != 4 ]; then cp cvpr2020_manet_b200/lib/libmanet_dbg$n.so cvpr2020_manet_b200/lib/libmanet_b200.so; else cp /tmp/lib_full.so cvpr2020_manet_b200/lib/libmanet_b200.so; fi
  for v in 0 2; do
  MANET_GM_VARIANT=$v timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > /tmp/b.log 2>&1
  python - <<PY
import json
d=json.loads(open('/tmp/b.log').read().strip().splitlines()[-1]); print('chunks $n variant $v: umma ms', round(d['roofline']['kernel_ms'],4))
PY
  done
done
