cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --workload 1080p8_n7 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b1.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1080p8', round(d['value'],2),'fps', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],2), 'int_frac',round(d['roofline']['int']['frac'],4))"
timeout 600 python bench.py --workload 1080p10_n11 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b2.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1080p10', round(d['value'],2),'fps', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],2), 'int_frac',round(d['roofline']['int']['frac'],4))"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/b3.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('4k10', round(d['value'],2),'fps', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],2), 'int_frac',round(d['roofline']['int']['frac'],4), d['roofline']['phases_ms'])"
tail -2 gpurun_out/b3.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tr "\n" " "; echo
