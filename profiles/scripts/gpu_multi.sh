#!/bin/bash
# Multi-GPU lines on one box:  gpurun --gpus N -- 'bash profiles/scripts/gpu_multi.sh N TAG'
N=${1:-8}
TAG=${2:-x}
mkdir -p gpurun_out
run() {  # name, extra bench args
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
        --steps 20 --warmup 5 --no-cpu-baseline $2 > gpurun_out/bench_$1_n${N}_$TAG.json 2> gpurun_out/bench_$1_n${N}_$TAG.err
    head -c 260 gpurun_out/bench_$1_n${N}_$TAG.json; echo; tail -2 gpurun_out/bench_$1_n${N}_$TAG.err
}
run schnet ""
run painn "--model painn"
run strong "--global-batch 256"
run var "--atoms-max 60"
