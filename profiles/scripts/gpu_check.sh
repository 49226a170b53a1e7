#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, the ncu launch list of one step and one `ncu --set full`
# capture of the hot kernels.  Usage (from the repo root):  gpurun --timeout 1800 -- 'bash profiles/scripts/gpu_check.sh TAG'
# The .ncu-rep is summarised ON THE BOX (gpurun_out/ comes back only below 64 MiB) and kept only if it is small.
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
if [ "$2" != "nocap" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
    python profiles/profile_step.py --steps 1 > gpurun_out/ncu_l_$TAG.log 2>&1
python profiles/summarize.py launches gpurun_out/launches_$TAG.csv > gpurun_out/launches_$TAG.txt 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"filter_bwd_tc_kernel|ddm_head_tc_kernel|cfconv_gather|cfconv_pairs|linear_wgrad_tc_kernel|filter_fwd_tc_kernel|linear_chain_tc_kernel" \
    -c 44 -o gpurun_out/prof_$TAG python profiles/profile_step.py --steps 1 > gpurun_out/ncu_f_$TAG.log 2>&1
python profiles/summarize.py full gpurun_out/prof_$TAG.ncu-rep > gpurun_out/ncu_full_$TAG.txt 2>&1
if [ $(stat -c %s gpurun_out/prof_$TAG.ncu-rep 2>/dev/null || echo 0) -gt 30000000 ]; then rm -f gpurun_out/prof_$TAG.ncu-rep; fi
fi
python bench.py --steps 20 --warmup 5 --atoms-max 60 --no-cpu-baseline > gpurun_out/bench_var_$TAG.json 2> gpurun_out/bench_var_$TAG.err
tail -3 gpurun_out/pytest_$TAG.txt; head -c 300 gpurun_out/bench_var_$TAG.json; echo; tail -2 gpurun_out/bench_var_$TAG.err; head -c 400 gpurun_out/bench_$TAG.json; echo; tail -2 gpurun_out/bench_$TAG.err
