#!/bin/bash
# step enqueue order / stream priority matrix
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" MANET_BENCH_CPU=0 MANET_BENCH_SHARDED=0 MANET_BENCH_SEGHEAD=0 MANET_BENCH_LEGS=0 timeout -s KILL 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/bench_{tag}.log') if x.startswith('{')]
if not l: print(tag,'NO LINE'); print(open(f'gpurun_out/bench_{tag}.log').read()[-1500:]); sys.exit()
d=json.loads(l[-1]); print(tag,{k:d[k] for k in ('value','ms_per_step')},'e2e',d['e2e']['value'],'serial',d['single_stream']['ms_per_step'], 'first',d['first_frame']['ms_per_step'])
PY
}
run global_prio  X=1
run global_noprio MANET_STEP_PRIO=0
run local_prio MANET_STEP_ORDER=local
run local_noprio MANET_STEP_ORDER=local MANET_STEP_PRIO=0
run global_prio2  X=1
