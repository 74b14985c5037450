#!/bin/bash
# N-GPU check: bench.py under torchrun (independent sequences + the NCCL-sharded 1080p leg with its parity assertion).  Usage: gpu_r02_nN.sh N
N=${1:-2}
mkdir -p gpurun_out
MANET_BENCH_CPU=0 timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
python - $N <<'PY'
import json, sys
N=sys.argv[1]
try:
    l=[x for x in open(f'gpurun_out/bench_n{N}.log') if x.startswith('{')][-1]; d=json.loads(l)
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'e2e full', d['e2e']['full_copy']['value'])
    sh=d['roofline'].get('sharded_global_1080p') or {}
    for k,v in sh.items():
        if isinstance(v,dict): print(k, {kk:v[kk] for kk in ('n_gpus','ms','match_ms','allreduce_us','allreduce_share_of_step','per_gpu_algorithmic_tflops','frac_of_sustained_peak','max_rel_err') if kk in v})
except Exception as e:
    print('parse failed', e); print(open(f'gpurun_out/bench_n{N}.err').read()[-2000:])
PY
