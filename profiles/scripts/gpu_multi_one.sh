#!/bin/bash
# One multi-GPU bench line:  gpurun --gpus N -- 'bash profiles/scripts/gpu_multi_one.sh N TAG NAME "extra bench args"'
N=${1:-8}; TAG=${2:-x}; NAME=${3:-schnet}; EXTRA=${4:-}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/bench_${NAME}_n${N}_$TAG.json 2> gpurun_out/bench_${NAME}_n${N}_$TAG.err
head -c 260 gpurun_out/bench_${NAME}_n${N}_$TAG.json; echo; tail -2 gpurun_out/bench_${NAME}_n${N}_$TAG.err
