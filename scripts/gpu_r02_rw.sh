#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_filter_refine.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_autograd.py -m gpu -q -x > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_g.log
timeout -s KILL 200 python scripts/fr_variant_time.py 20 2>&1 | tail -4
