cd $GRAFT_REPO_ROOT
nproc; free -g | head -2
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --workload 1080p8_n7 --steps 3 --warmup 3 > gpurun_out/bench_1080p8.json 2> gpurun_out/bench_1080p8.err; tail -3 gpurun_out/bench_1080p8.err; cat gpurun_out/bench_1080p8.json
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err; tail -3 gpurun_out/bench_4k.err; cat gpurun_out/bench_4k.json
