#!/bin/bash
# Staged GPU check run under gpurun: every stage has its own hard timeout and log so that a hang
# in one stage cannot eat the others.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== probe" ; timeout -s KILL 180 python scripts/umma_probe.py 24 30 24 3 > gpurun_out/probe_small.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/probe_small.log
echo "== probe big" ; timeout -s KILL 180 python scripts/umma_probe.py 120 214 120 6 > gpurun_out/probe_big.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/probe_big.log
echo "== pytest (everything except the tcgen05 engine)"; timeout -s KILL 900 python -m pytest tests -m gpu -q -k "not tcgen05 and not full_480p and not seeded_medium and not session" -x > gpurun_out/pytest_simt.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_simt.log
echo "== pytest (full)"; timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench"; timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench.log
