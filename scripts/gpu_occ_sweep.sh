cd $GRAFT_REPO_ROOT
run() { wl=$1; f=$2; cv=$3
if [ -n "$cv" ]; then export TF_GPU_CARVEOUT=$cv; else unset TF_GPU_CARVEOUT; fi
TF_GPU_LIB=$GRAFT_REPO_ROOT/$f python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl $f cv=$cv', round(d['value'],2), {k: round(x,2) for k,x in d['roofline']['phases_ms'].items()}, d['verified'])"
}
for wl in 4k10_n15 1080p8_n7; do
run $wl gpurun_ab/lib_a_prev.so ""
run $wl gpurun_ab/lib_b_w28.so 100
run $wl gpurun_ab/lib_c_w32.so 100
run $wl gpurun_ab/lib_a_prev.so 100
done
