# bench every gpurun_ab/lib_*.so on the same box (development A/B sweeps)
# usage: gpu_sweep_libs.sh [workload] [extra bench args...]
cd $GRAFT_REPO_ROOT
WL=${1:-4k10_n15}; shift
for i in 1 2; do
for f in gpurun_ab/lib_*.so; do
TF_GPU_LIB=$GRAFT_REPO_ROOT/$f python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['value'],2), {k: round(x,2) for k,x in d['roofline']['phases_ms'].items()})"
done; done
