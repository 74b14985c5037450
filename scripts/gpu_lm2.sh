#!/bin/bash
# local tcgen05 engine: parity vs CUDA-core engine + timing, local GPU tests, optional cycle trace / ncu
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/lm_debug.py > gpurun_out/lm_debug.log 2>&1; echo "debug rc=$?"; tail -20 gpurun_out/lm_debug.log
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -k "local or session or step or smoke" > gpurun_out/pytest_local.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_local.log
if [ "${1:-}" = "ncu" ]; then
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"lm_" -s 6 -c 3 -f -o gpurun_out/prof_lm python scripts/lm_prof.py > gpurun_out/ncu_lm.log 2>&1; echo "ncu rc=$?"
fi
if [ "${1:-}" = "trace" ] || [ "${2:-}" = "trace" ]; then
  MANET_NVCC_EXTRA=-DLM_TRACE python -m cvpr2020_manet_b200.build --force > gpurun_out/lm_trace_build.log 2>&1
  timeout -s KILL 200 python scripts/lm_prof.py > gpurun_out/lm_trace.log 2>&1; echo "trace rc=$?"; sort gpurun_out/lm_trace.log | uniq | tail -24
fi
