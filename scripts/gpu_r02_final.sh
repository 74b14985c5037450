#!/bin/bash
# r02 round check: all GPU tests, smoke, bench (both arms), ncu launch list, full ncu captures of the matching kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench.log
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench(reference) rc=$?"; tail -c 300 gpurun_out/bench_ref.log
if [ "${1:-}" = "ncu" ]; then
  export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0 MANET_BENCH_SEGHEAD=0
  timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"gm_fr_kernel|gm_refine_kernel|gm_rescan_kernel|gm_convert_kernel|gm_scan_kernel|lm_pool_kernel|lm_convert_kernel|lm_umma_kernel|local_map_store_select" -s 22 -c 9 -f -o gpurun_out/r02_matching python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "full ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
fi
ls -la gpurun_out/ | head -30
