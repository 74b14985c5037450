#!/bin/bash
# r02 second call: filter-and-refine global engine -- its tests first (short timeout: a hang must not cost the box), then the
# global parity tests, the fixed TMEM micro-benchmark, bench.
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_filter_refine.py -m gpu -q -x > gpurun_out/pytest_fr.log 2>&1; echo "pytest(fr) rc=$?"; tail -25 gpurun_out/pytest_fr.log
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_autograd.py tests/test_deeplab.py -m gpu -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest(parity) rc=$?"; tail -25 gpurun_out/pytest_parity.log
timeout -s KILL 120 python scripts/tmem_ld_bench.py gpurun_out/r02_tmem_ld.json > gpurun_out/tmem_ld.log 2>&1; echo "tmem rc=$?"; tail -20 gpurun_out/tmem_ld.log
MANET_BENCH_CPU=0 timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1]; d=json.loads(l)
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'full',d['e2e']['full_copy']['value'])
    r=d['roofline']; print({k:r[k] for k in ('achieved','frac','kernel_ms','refine_ms','core_ms','frac_filter_kernel_only','frac_executed')}); print(r['local']['main_kernel_ms'],r['local']['prepass_ms'])
    print('single_stream',d['single_stream']); print('intvos',json.dumps(d.get('intvos_forward'))[:1500])
    print('sharded',json.dumps(r.get('sharded_global_1080p'))[:1800])
    print('prop',d['propagation_50']['frames_per_s'],'session',d['session_8_rounds']['frames_per_s'], d['session_8_rounds']['interaction_branch_ms'])
except Exception as e:
    print('parse failed',e); print(open('gpurun_out/bench.log').read()[-3000:])
PY
MANET_GM_ENGINE=exact3 MANET_BENCH_CPU=0 MANET_BENCH_SEGHEAD=0 MANET_BENCH_SHARDED=0 timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_exact3.log 2>&1; echo "bench(exact3) rc=$?"; tail -c 400 gpurun_out/bench_exact3.log | head -c 400
