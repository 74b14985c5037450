#!/bin/bash
# session-4 first call: full GPU tests, bench with both local engines, local-engine debug/timing, lm ncu capture
mkdir -p gpurun_out
export MANET_BENCH_SHARDED=0 MANET_BENCH_CPU=0
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_umma.log 2>&1; echo "bench(umma local) rc=$?"
MANET_LM_ENGINE=simt timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_simt.log 2>&1; echo "bench(simt local) rc=$?"
python - <<'PY'
import json
for f in ('bench_umma','bench_simt'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
        r=d['roofline']
        print(f,'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'single',round(d['single_stream']['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),
              'umma ms',round(r['kernel_ms'],4),'local main ms',r['local']['main_kernel_ms'],'prepass ms',r['local']['prepass_ms'])
    except Exception as e:
        print(f,'parse failed',e); print(open(f'gpurun_out/{f}.log').read()[-1500:])
PY
timeout -s KILL 400 python scripts/lm_debug.py > gpurun_out/lm_debug.log 2>&1; echo "debug rc=$?"; tail -22 gpurun_out/lm_debug.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"lm_umma_kernel|lm_pool" -s 4 -c 2 -f -o gpurun_out/prof_lm python scripts/lm_prof.py > gpurun_out/ncu_lm.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/
