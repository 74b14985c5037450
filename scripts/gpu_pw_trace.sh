#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/seghead_times.py > gpurun_out/pw_trace.log 2>&1; echo "rc=$?"
grep "^pw" gpurun_out/pw_trace.log | tail -16
