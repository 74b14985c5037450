#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_filter_refine.py -m gpu -q -x > gpurun_out/pytest_fr.log 2>&1; echo "pytest(fr) rc=$?"; tail -5 gpurun_out/pytest_fr.log
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_autograd.py -m gpu -q -x > gpurun_out/pytest_parity.log 2>&1; echo "pytest(parity) rc=$?"; tail -5 gpurun_out/pytest_parity.log
timeout -s KILL 120 python scripts/gm_once.py 4 2>&1 | tail -3
MANET_BENCH_CPU=0 MANET_BENCH_SEGHEAD=0 MANET_BENCH_SHARDED=0 timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]; d=json.loads(l)
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'full',d['e2e']['full_copy']['value'])
    r=d['roofline']; print({k:r[k] for k in ('achieved','frac','kernel_ms','refine_ms','core_ms','frac_filter_kernel_only','frac_executed')}); print(r['local']['main_kernel_ms'],r['local']['prepass_ms'])
    print('single_stream',d['single_stream'])
except Exception as e:
    print('parse failed',e); print(open('gpurun_out/bench.log').read()[-3000:])
PY
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"gm_fr_kernel|gm_refine_kernel|gm_rescan_kernel" -s 6 -c 3 -f -o gpurun_out/r02_gm_fr2 python scripts/gm_once.py 4 > gpurun_out/ncu_gm.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_gm.log
timeout -s KILL 120 python scripts/gm_reuse_probe.py 2>&1 | tail -3; timeout -s KILL 200 python scripts/lm_error_vs_g.py gpurun_out/r02_lm_error_vs_g.json 2>&1 | tail -14
