#!/bin/bash
# local-matching engine: parity/timing script, then an in-kernel cycle trace (rebuilt with -DLM_TRACE on the box)
mkdir -p gpurun_out
if [ "${2:-debug}" = "debug" ]; then timeout -s KILL 400 python scripts/lm_debug.py > gpurun_out/lm_debug.log 2>&1; echo "debug rc=$?"; tail -22 gpurun_out/lm_debug.log; fi
if [ "${1:-trace}" = "trace" ]; then
  MANET_NVCC_EXTRA=-DLM_TRACE python -m cvpr2020_manet_b200.build --force > gpurun_out/lm_trace_build.log 2>&1
  timeout -s KILL 200 python scripts/lm_prof.py > gpurun_out/lm_trace.log 2>&1; echo "trace rc=$?"; tail -12 gpurun_out/lm_trace.log
fi
