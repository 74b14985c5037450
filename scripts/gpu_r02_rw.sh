#!/bin/bash
mkdir -p gpurun_out
for v in 1 0 1; do
MANET_PROP_STREAMS=$v MANET_BENCH_CPU=0 MANET_BENCH_SHARDED=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_p$v.log 2>&1; python - $v <<'PY'
import json,sys
l=[x for x in open(f'gpurun_out/bench_p{sys.argv[1]}.log') if x.startswith('{')][-1]; d=json.loads(l)
s=d['session_8_rounds']
print('streams',sys.argv[1],'prop',d['propagation_50']['frames_per_s'],'session',s['frames_per_s'],s['ms_per_round'],'intvos',d['intvos_forward']['random_init_eval']['matching_and_head_ms'])
PY
done
