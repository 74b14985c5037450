#!/bin/bash
# multi-GPU bench check: usage bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
export MANET_BENCH_CPU=0
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/bench_n$N.log | cut -c1-1500
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-400
