#!/bin/bash
# ncu --set full of the pair-centric cfconv kernel alone (profiles/bench_cfconv.py); the report is small enough to come back.
TAG=${1:-x}
mkdir -p gpurun_out
export GEOSSL_BENCH_CFCONV_TUNINGS=${2:-216}
ncu --set full --clock-control none --import-source on -k regex:"cfconv_pairs" --launch-skip 8 -c 2 \
    -o gpurun_out/prof_cfconv_$TAG python profiles/bench_cfconv.py > gpurun_out/ncu_cfconv_$TAG.log 2>&1
ls -la gpurun_out/prof_cfconv_$TAG.ncu-rep; tail -3 gpurun_out/ncu_cfconv_$TAG.log
