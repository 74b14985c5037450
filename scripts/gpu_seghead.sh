#!/bin/bash
# seghead bring-up: tests (guarded by timeouts: a hung tcgen05 kernel must not hold the box), then kernel times
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_seghead.py -m gpu -x -q > gpurun_out/seghead_tests.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/seghead_tests.log
timeout -s KILL 200 python scripts/seghead_times.py > gpurun_out/seghead_times.log 2>&1; echo "times rc=$?"; tail -20 gpurun_out/seghead_times.log
