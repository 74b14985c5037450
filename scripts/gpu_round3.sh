#!/bin/bash
# Round check with the seghead: all GPU tests, smoke, bench, reference arm, ncu launch list, full ncu of the seghead kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.log
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench(reference) rc=$?"; tail -c 400 gpurun_out/bench_ref.log
if [ "${1:-}" = "ncu" ]; then
  export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
  timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"sh_dw_kernel|sh_pw_kernel" -s 10 -c 4 -f -o gpurun_out/prof_seghead python scripts/seghead_times.py > gpurun_out/ncu_seghead.log 2>&1; echo "seghead ncu rc=$?"
fi
ls -la gpurun_out/ | head -30
