#!/bin/bash
# one full ncu capture of the seghead kernels (dw<256>, pw<0>) at 480p
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"${1:-sh_dw_kernel}" -s ${2:-6} -c ${3:-1} \
    -f -o gpurun_out/prof_seghead python scripts/seghead_times.py > gpurun_out/ncu_seghead.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_seghead.log; ls -la gpurun_out/
