cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full_4k.json 2> gpurun_out/bench_full_4k.err; tail -3 gpurun_out/bench_full_4k.err; cat gpurun_out/bench_full_4k.json | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_4k.json 2> gpurun_out/bench_ref_4k.err; tail -3 gpurun_out/bench_ref_4k.err; cat gpurun_out/bench_ref_4k.json
