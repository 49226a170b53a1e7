#!/bin/bash
# Experiment runner: each argument is "tag:ENV=VAL[,ENV=VAL]" -> one default bench line per argument (no CPU baseline).
mkdir -p gpurun_out
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}; envs=${envs//,/ }
  env $envs python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_exp_$tag.json 2> gpurun_out/bench_exp_$tag.err
  echo "$tag [$envs]"; head -c 260 gpurun_out/bench_exp_$tag.json | tail -c 130; echo; tail -1 gpurun_out/bench_exp_$tag.err
done
