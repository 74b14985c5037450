#!/bin/bash
cp cvpr2020_manet_b200/lib/libmanet_b200.so /tmp/lib_full.so
cp cvpr2020_manet_b200/lib/libmanet_trace.so cvpr2020_manet_b200/lib/libmanet_b200.so
timeout -s KILL 120 python scripts/umma_probe.py 120 214 120 6 2>&1 | tail -24
cp /tmp/lib_full.so cvpr2020_manet_b200/lib/libmanet_b200.so
