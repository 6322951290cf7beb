cd $GRAFT_REPO_ROOT
for c in 1 2 3 4; do
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --concurrent $c 2>gpurun_out/c.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'],'conc', d['config']['windows_in_flight_per_gpu'], round(d['value'],2),'fps', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],2), d['roofline']['phases_ms'])"; tail -2 gpurun_out/c.err
done
