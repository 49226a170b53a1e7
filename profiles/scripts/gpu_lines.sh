mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_v32.json 2> gpurun_out/bench_v32.err
head -c 300 gpurun_out/bench_v32.json | tail -c 170; echo
GEOSSL_GRAPH_PRIORITY=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v32_noprio.json 2> gpurun_out/bench_v32_noprio.err
head -c 300 gpurun_out/bench_v32_noprio.json | tail -c 170; echo
python bench.py --workload md17 --steps 10 --warmup 3 > gpurun_out/bench_md17_v32.json 2> gpurun_out/bench_md17_v32.err
python bench.py --workload lba --steps 10 --warmup 3 > gpurun_out/bench_lba_v32.json 2> gpurun_out/bench_lba_v32.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v32.json 2> gpurun_out/bench_ref_v32.err
for w in md17 lba ref; do head -c 420 gpurun_out/bench_${w}_v32.json; echo; tail -2 gpurun_out/bench_${w}_v32.err; done
