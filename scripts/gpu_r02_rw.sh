#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_g.log
MANET_BENCH_CPU=0 MANET_BENCH_SHARDED=0 MANET_BENCH_SEGHEAD=0 timeout -s KILL 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]; d=json.loads(l)
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',d['e2e'])
r=d['roofline']; print({k:r.get(k) for k in ('frac','kernel_ms','refine_kernel_ms','rescan_kernel_ms','other_engine_chain_ms','prepass_ms','core_ms','frac_filter_kernel_only')})
PY
