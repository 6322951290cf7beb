# A/B on one box: GPU test suite on the current build, then bench.py alternating between the libraries
# given as arguments ("cur" = aom-av1-psy_b200/libtf_gpu.so).  usage: gpu_ab2.sh [workload] lib...
cd $GRAFT_REPO_ROOT
WL=${1:-4k10_n15}; shift
LIBS="$@"; [ -z "$LIBS" ] && LIBS="gpurun_ab/libtf_gpu_r01.so cur"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
for v in $LIBS; do
if [ $v = cur ]; then unset TF_GPU_LIB; else export TF_GPU_LIB=$GRAFT_REPO_ROOT/$v; fi
timeout 600 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],2), 'fps', {k: round(x,2) for k,x in d['roofline']['phases_ms'].items()})" || tail -5 gpurun_out/ab.err
done; done
