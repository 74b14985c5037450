#!/bin/bash
# in-kernel cycle trace of gm_fr_kernel (library built with -DFR_TRACE as lib/libmanet_b200_trace.so)
mkdir -p gpurun_out
cp cvpr2020_manet_b200/lib/libmanet_b200_trace.so cvpr2020_manet_b200/lib/libmanet_b200.so
timeout -s KILL 300 python scripts/gm_once.py 6 2>&1 | tail -14 | tee gpurun_out/fr_trace.txt
