cd $GRAFT_REPO_ROOT
for wl in 1080p8_n7 1080p10_n11 4k10_n15; do
timeout 900 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/c.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'],'conc', d['config']['windows_in_flight_per_gpu'], round(d['value'],2),'fps', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'], d['clocks'])"; tail -2 gpurun_out/c.err
done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tr "\n" " "; echo
