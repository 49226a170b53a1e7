#!/bin/bash
# GPU tests + cfconv A/B microbenchmark + default bench line (no ncu).
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.txt
python profiles/bench_cfconv.py > gpurun_out/cfconv_ab_$TAG.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -4 gpurun_out/pytest_$TAG.txt; cat gpurun_out/cfconv_ab_$TAG.txt; head -c 300 gpurun_out/bench_$TAG.json; echo; tail -2 gpurun_out/bench_$TAG.err
