#!/bin/bash
# Full GPU check: all GPU tests, smoke, bench (default + CUDA-core local engine), ncu launch list + full capture of the hot kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.log
MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0 MANET_LM_ENGINE=simt timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_simt.log 2>&1; echo "bench(simt local) rc=$?"
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench(reference) rc=$?"; tail -c 600 gpurun_out/bench_ref.log
if [ "${1:-}" = "ncu" ]; then
  export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
  timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"gm_umma2_kernel|lm_umma_kernel|lm_pool|lm_convert|gm_convert|gm_scan|gm_finalize|local_map_store" -s 24 -c 8 -f -o gpurun_out/prof_full python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
fi
ls -la gpurun_out/ | head -30
