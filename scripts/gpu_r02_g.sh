#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" MANET_BENCH_CPU=0 MANET_BENCH_SHARDED=0 MANET_BENCH_SEGHEAD=0 timeout -s KILL 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/bench_{tag}.log') if x.startswith('{')]
if not l: print(tag,'NO LINE'); print(open(f'gpurun_out/bench_{tag}.log').read()[-1500:]); sys.exit()
d=json.loads(l[-1]); print(tag,{k:d[k] for k in ('value','ms_per_step')},'e2e',d['e2e']['value'],'serial',d['single_stream']['ms_per_step'], 'first',d['first_frame']['ms_per_step'])
PY
}
run plain X=1
run pdlg MANET_PDL_GLOBAL=1
run plain2 X=1
run pdlg2 MANET_PDL_GLOBAL=1
