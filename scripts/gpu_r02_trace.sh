#!/bin/bash
# in-kernel cycle trace of gm_fr_kernel (python scripts/build_fr_variants.py first)
mkdir -p gpurun_out
cp cvpr2020_manet_b200/lib/libmanet_b200_trace.so cvpr2020_manet_b200/lib/libmanet_b200.so
timeout -s KILL 300 python scripts/gm_once.py 2 2>&1 | grep -v "^cta 80" | tail -44 | tee gpurun_out/fr_trace.txt
