cd $GRAFT_REPO_ROOT
python scripts/parity_report.py r02 300 > gpurun_out/parity_r02.log 2>&1; tail -2 gpurun_out/parity_r02.log
run() { wl=$1; shift
python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys,os; d=json.loads(sys.stdin.read()); print('$wl $*', os.environ.get('TF_GPU_PRIO'), os.environ.get('TF_GPU_S16'), round(d['value'],2), {k: round(x,2) for k,x in d['roofline']['phases_ms'].items()}, d['verified'])"
}
run 4k10_n15
TF_GPU_PRIO=flat run 4k10_n15
TF_GPU_S16=single run 4k10_n15
TF_GPU_PRIO=flat TF_GPU_S16=single run 4k10_n15
run 1080p10_n11
TF_GPU_PRIO=flat run 1080p10_n11
TF_GPU_S16=single run 1080p10_n11
run 1080p10_n11 --concurrent 3
run 1080p10_n11 --concurrent 6
run 1080p8_n7 --concurrent 6
