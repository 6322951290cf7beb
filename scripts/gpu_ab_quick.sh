cd $GRAFT_REPO_ROOT
run() { wl=$1; f=$2; shift 2
TF_GPU_LIB=$GRAFT_REPO_ROOT/$f python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl $f $*', round(d['value'],2), {k: round(x,2) for k,x in d['roofline']['phases_ms'].items()}, d['verified'])"
}
for wl in 4k10_n15 1080p8_n7; do
for f in gpurun_ab/lib_*.so; do run $wl $f; done; done
