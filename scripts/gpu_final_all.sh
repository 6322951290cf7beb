# final single-GPU records for README / profiles (default steps)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tr "\n" " "; echo
for wl in 4k10_n15 1080p10_n11 1080p8_n7; do
  timeout 900 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -1 gpurun_out/bench_$wl.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), 'int frac', round(d['roofline']['frac'],4), 'exec', round(d['roofline']['executed']['frac'],4), 'verified', d['verified'], 'cpu1', d['cpu_baseline']['value'], d['cpu_baseline'].get('generic_c',{}).get('value'), d['clocks'])"
  timeout 600 python bench.py --workload $wl --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$wl.json 2>gpurun_out/bench_ref_$wl.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_ref_$wl.json')); print(' ref', round(d['value'],3), d['cpu_baseline'])"
done
timeout 600 python bench.py --impl reference --ref-simd c --steps 1 --warmup 1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(' ref generic C 4k', round(d['value'],3), d['cpu_baseline']['cores'])"
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --concurrent 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('conc1 4k', round(d['value'],2), 'ms/window', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2))"
python bench.py --workload 1080p8_n7 --steps 8 --warmup 3 --no-cpu-baseline --concurrent 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('conc1 1080p8', round(d['value'],2), 'ms/window', round(d['ms_per_step'],2))"
python __graft_entry__.py smoke 2>&1 | tail -1
python scripts/parity_report.py ${1:-r02} 300 2>&1 | tail -3
