#!/bin/bash
# Quick iteration call: selected GPU tests, the cfconv micro-benchmark and (unless MODE = "micro") the default and
# variable-size bench lines with and without the pair-centric cfconv kernel.
# Usage: gpurun --timeout 900 -- 'bash profiles/scripts/gpu_iter.sh TAG "tests/test_gpu_schnet.py -k pair" [micro]'
TAG=${1:-x}
SEL=${2:-tests}
MODE=${3:-full}
mkdir -p gpurun_out
python -m pytest $SEL -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_$TAG.txt
python profiles/bench_cfconv.py > gpurun_out/cfconv_ab_$TAG.txt 2>&1
tail -5 gpurun_out/pytest_$TAG.txt; cat gpurun_out/cfconv_ab_$TAG.txt
if [ "$MODE" != "micro" ]; then
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
GEOSSL_CFCONV_PAIRS=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nopairs_$TAG.json 2> gpurun_out/bench_nopairs_$TAG.err
python bench.py --steps 20 --warmup 5 --atoms-max 60 --no-cpu-baseline > gpurun_out/bench_var_$TAG.json 2> gpurun_out/bench_var_$TAG.err
GEOSSL_CFCONV_PAIRS=0 python bench.py --steps 20 --warmup 5 --atoms-max 60 --no-cpu-baseline > gpurun_out/bench_var_nopairs_$TAG.json 2> gpurun_out/bench_var_nopairs_$TAG.err
for f in bench_$TAG bench_nopairs_$TAG bench_var_$TAG bench_var_nopairs_$TAG; do head -c 260 gpurun_out/$f.json | tail -c 130; echo; tail -1 gpurun_out/$f.err; done
fi
