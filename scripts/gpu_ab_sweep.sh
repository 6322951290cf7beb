# A/B of builds on one box: the GPU suite on the in-tree build, then every gpurun_ab/lib_*.so (built with different -D switches)
# on three workloads (TF_GPU_LIB selects the library).  usage (via gpurun): bash scripts/gpu_ab_sweep.sh
cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
run() { wl=$1; f=$2; shift 2
TF_GPU_LIB=$GRAFT_REPO_ROOT/$f python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl $f $*', round(d['value'],2), 'e2e', round(d['e2e']['value'],2) if d.get('e2e') else None, {k: round(x,2) for k,x in d['roofline']['phases_ms'].items()}, d['verified'])"
}
for wl in 4k10_n15 1080p10_n11 1080p8_n7; do
for f in gpurun_ab/lib_*.so; do run $wl $f --no-e2e; done; done
