cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tr "\n" " "; echo
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; tail -2 gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json | cut -c1-400
python __graft_entry__.py smoke 2>&1 | tail -1
