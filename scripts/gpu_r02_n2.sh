#!/bin/bash
# two-GPU check: bench.py under torchrun (independent sequences + the NCCL-sharded 1080p leg with its parity assertion)
mkdir -p gpurun_out
MANET_BENCH_CPU=0 timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"
python - <<'PY'
import json
try:
    l=[x for x in open('gpurun_out/bench_n2.log') if x.startswith('{')][-1]; d=json.loads(l)
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])
    print(json.dumps(d['roofline'].get('sharded_global_1080p'))[:2500])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_n2.err').read()[-2000:])
PY
