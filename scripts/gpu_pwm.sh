#!/bin/bash
mkdir -p gpurun_out
for cl in 2 4; do
  export MANET_SH_PW_CLUSTER=$cl
  timeout -s KILL 300 python -m pytest tests/test_seghead.py -m gpu -x -q > gpurun_out/seghead_tests_cl$cl.log 2>&1; echo "cluster $cl pytest rc=$?"; tail -3 gpurun_out/seghead_tests_cl$cl.log
  timeout -s KILL 200 python scripts/seghead_times.py > gpurun_out/seghead_times_cl$cl.log 2>&1; echo "times rc=$?"; grep -E "seghead forward|sh_pw|sh_dw" gpurun_out/seghead_times_cl$cl.log
done
