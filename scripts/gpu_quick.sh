#!/bin/bash
# quick check: GPU tests + short bench. Usage: bash scripts/gpu_quick.sh [pytest -k expr]
mkdir -p gpurun_out
export MANET_BENCH_SHARDED=${MANET_BENCH_SHARDED:-0} MANET_BENCH_CPU=${MANET_BENCH_CPU:-0}
timeout -s KILL 900 python -m pytest tests -m gpu -q ${1:+-k "$1"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    r=d['roofline']
    print('value',round(d['value'],1),'fps  ms/step',round(d['ms_per_step'],4),' e2e',round(d['e2e']['value'],1))
    print('umma ms',r['kernel_ms'],'frac',r['frac'],' local main ms',r['local']['main_kernel_ms'],' prepass ms',r['local']['prepass_ms'])
    print('clocks',d['clocks'])
except Exception as e:
    print('bench parse failed',e); print(open('gpurun_out/bench.log').read()[-2000:])
PY
