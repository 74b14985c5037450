#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_g.log
run() {
  tag=$1; shift
  env "$@" MANET_BENCH_CPU=0 MANET_BENCH_SHARDED=0 MANET_BENCH_SEGHEAD=0 timeout -s KILL 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/bench_{tag}.log') if x.startswith('{')]
if not l: print(tag,'NO LINE'); print(open(f'gpurun_out/bench_{tag}.log').read()[-1500:]); sys.exit()
d=json.loads(l[-1]); print(tag,{k:d[k] for k in ('value','ms_per_step','gpu_launches')},'e2e',d['e2e']['value'],'serial',d['single_stream']['ms_per_step'], 'first',d['first_frame']['ms_per_step'])
PY
}
run aux X=1
run noaux MANET_STEP_AUX=0
run aux2 X=1
run noaux2 MANET_STEP_AUX=0
