#!/bin/bash
# A/B of compile-time variants of the local tcgen05 kernel: bash scripts/gpu_lm_var.sh "<nvcc defines>" ...
mkdir -p gpurun_out
for v in "$@"; do
  MANET_NVCC_EXTRA="$v" python -m cvpr2020_manet_b200.build --force > gpurun_out/var_build.log 2>&1 || { echo "build failed for $v"; tail -5 gpurun_out/var_build.log; continue; }
  echo "=== variant [$v]"
  timeout -s KILL 200 python scripts/lm_debug.py 2>&1 | grep -E "self\] H=120 W=214 C=100 N=6 d=12|rand\] H=120 W=214 C=100 N=6 d=12|tcgen05:"
  timeout -s KILL 200 python scripts/kernel_times.py 2>&1 | grep "lm_umma"
done
