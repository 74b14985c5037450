#!/bin/bash
# gm_fr_kernel bottleneck experiments: normal | loads only | arithmetic only | cycle trace   (python scripts/build_fr_variants.py first)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_filter_refine.py -m gpu -q -x 2>&1 | tail -2
cp cvpr2020_manet_b200/lib/libmanet_b200.so /tmp/normal.so
for v in normal nocompute noload; do
  if [ $v != normal ]; then cp cvpr2020_manet_b200/lib/libmanet_b200_$v.so cvpr2020_manet_b200/lib/libmanet_b200.so; fi
  echo "== $v"; timeout -s KILL 200 python scripts/fr_variant_time.py 12 2>&1 | grep gm_fr
done
cp cvpr2020_manet_b200/lib/libmanet_b200_trace.so cvpr2020_manet_b200/lib/libmanet_b200.so
timeout -s KILL 300 python scripts/gm_once.py 2 2>&1 | grep -v "^cta 80" | tail -44 > gpurun_out/fr_trace.txt
grep "MMA tl\|EPI cta 0\|cta 0 warp" gpurun_out/fr_trace.txt | tail -14
cp /tmp/normal.so cvpr2020_manet_b200/lib/libmanet_b200.so
