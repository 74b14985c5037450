#!/bin/bash
# one full ncu capture of the seghead kernels (dw<256>, pw<0>) at 480p
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"sh_dw_kernel|sh_pw_kernel" -s 10 -c 4 \
    -f -o gpurun_out/prof_seghead python scripts/seghead_times.py > gpurun_out/ncu_seghead.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_seghead.log; ls -la gpurun_out/
