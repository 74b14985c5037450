#!/bin/bash
# gm_fr_kernel bottleneck experiments: normal | loads only | arithmetic only
mkdir -p gpurun_out
cp cvpr2020_manet_b200/lib/libmanet_b200.so /tmp/normal.so
for v in normal nocompute noload; do
  if [ $v != normal ]; then cp cvpr2020_manet_b200/lib/libmanet_b200_$v.so cvpr2020_manet_b200/lib/libmanet_b200.so; fi
  echo "== $v"; timeout -s KILL 200 python scripts/fr_variant_time.py 12 2>&1 | tail -4
done
cp /tmp/normal.so cvpr2020_manet_b200/lib/libmanet_b200.so
