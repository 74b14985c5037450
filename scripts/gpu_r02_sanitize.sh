#!/bin/bash
# compute-sanitizer: racecheck over the filter-and-refine tests, memcheck over the parity / session / seghead tests
mkdir -p gpurun_out
timeout -s KILL 900 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_filter_refine.py -m gpu -q -x > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_racecheck.log
timeout -s KILL 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_autograd.py -m gpu -q -x > gpurun_out/sanitize_memcheck2.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck2.log
