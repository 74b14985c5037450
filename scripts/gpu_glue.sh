#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_seghead.py -m gpu -x -q > gpurun_out/seghead_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/seghead_tests.log
MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_short.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_short.log').read().strip().splitlines()[-1])
    print("value",round(d["value"],1),"seghead ms",d["seghead"]["ms"],"prop",d.get("propagation_50",{}).get("frames_per_s"),"session",d.get("session_8_rounds"))
except Exception as e:
    print('parse failed',e); print(open('gpurun_out/bench_short.log').read()[-3000:])
PY
