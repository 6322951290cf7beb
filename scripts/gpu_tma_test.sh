cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 | cut -c1-300
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
run() { wl=$1; f=$2; shift 2
TF_GPU_LIB=$GRAFT_REPO_ROOT/$f timeout 300 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl $f $*', round(d['value'],2), {k: round(x,2) for k,x in d['roofline']['phases_ms'].items()}, d['verified'])"
}
for wl in 4k10_n15 1080p10_n11 1080p8_n7; do
for f in gpurun_ab/lib_*.so; do run $wl $f; done; done
