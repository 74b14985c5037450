#!/bin/bash
# one full ncu capture (with source counters) of gm_fr_kernel on the bench tensors
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"gm_fr_kernel" -s 2 -c 1 -f -o gpurun_out/r02_fr python scripts/gm_once.py 4 > gpurun_out/ncu_fr.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fr.log
