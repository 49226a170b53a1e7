#!/bin/bash
# Secondary workloads on one GPU: GPU tests, then the PaiNN-DDM (configs[2]), MD17 (configs[3]) and LBA (configs[4]) bench lines.
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_$TAG.txt
python bench.py --model painn --steps 20 --warmup 5 > gpurun_out/bench_painn_$TAG.json 2> gpurun_out/bench_painn_$TAG.err
python bench.py --workload md17 --steps 10 --warmup 3 > gpurun_out/bench_md17_$TAG.json 2> gpurun_out/bench_md17_$TAG.err
python bench.py --workload lba --steps 10 --warmup 3 > gpurun_out/bench_lba_$TAG.json 2> gpurun_out/bench_lba_$TAG.err
tail -4 gpurun_out/pytest_$TAG.txt
for w in painn md17 lba; do head -c 250 gpurun_out/bench_${w}_$TAG.json; echo; tail -3 gpurun_out/bench_${w}_$TAG.err; done
