#!/bin/bash
# GPU tests, the MD17 bench line, and the ncu launch list of one PaiNN-DDM step.
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_$TAG.txt
python bench.py --workload md17 --steps 10 --warmup 3 > gpurun_out/bench_md17_$TAG.json 2> gpurun_out/bench_md17_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_painn_$TAG.csv \
    python profiles/profile_step.py --steps 1 --model painn > gpurun_out/ncu_lp_$TAG.log 2>&1
tail -4 gpurun_out/pytest_$TAG.txt; head -c 300 gpurun_out/bench_md17_$TAG.json; echo; tail -3 gpurun_out/bench_md17_$TAG.err; tail -2 gpurun_out/ncu_lp_$TAG.log
