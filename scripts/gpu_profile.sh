#!/bin/bash
# ncu launch list + one full capture of the hot kernels (B200_PROFILING.md recipe). 1 GPU only.
# Usage: bash scripts/gpu_profile.sh [kernel-regex] [skip] [count]
mkdir -p gpurun_out
export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
REGEX=${1:-"gm_umma_kernel|window_dist|upsample_mask_min|gm_convert|gm_scan|gm_finalize|avg_pool2|local_map_store"}
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
    -k regex:"$REGEX" -s ${2:-16} -c ${3:-9} \
    -f -o gpurun_out/prof_full python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/
