#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_reference_corr.py tests/test_gpu_parity.py -m gpu -q -x -k "corr" > gpurun_out/pytest_corr.log 2>&1; echo "pytest(corr) rc=$?"; tail -6 gpurun_out/pytest_corr.log
timeout -s KILL 300 python scripts/corr_times.py gpurun_out/r02_correlation_timing.json 2>&1 | tail -4
