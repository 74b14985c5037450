#!/bin/bash
# r02 first call: TMEM drain micro-benchmark, the new full-size / interaction-head / grad tests, bench (both arms).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 120 python scripts/tmem_ld_bench.py gpurun_out/r02_tmem_ld.json > gpurun_out/tmem_ld.log 2>&1; echo "tmem rc=$?"; cat gpurun_out/tmem_ld.log | tail -20
timeout -s KILL 600 python -m pytest tests/test_gpu_fullsize.py tests/test_seghead.py tests/test_gpu_autograd.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest(new) rc=$?"; tail -15 gpurun_out/pytest_new.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.log
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench(reference) rc=$?"; tail -c 1200 gpurun_out/bench_ref.log
