# occupancy / carveout sweep (development)
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_search_batch.py -x -q -m gpu > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_batch.log | cut -c1-300
run() { # lib carve extra...
  lib=$1; cv=$2; shift 2
  if [ -n "$cv" ]; then export TF_GPU_CARVEOUT=$cv; else unset TF_GPU_CARVEOUT; fi
  TF_GPU_LIB=$GRAFT_REPO_ROOT/gpurun_ab/$lib python bench.py --workload 4k10_n15 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib cv=$cv $*', round(d['value'],2), {k: round(x,2) for k,x in d['roofline']['phases_ms'].items()})"
}
run lib_a_base.so ""
run lib_a_base.so 100
run lib_b_w32.so 100
run lib_c_s32lo16.so ""
run lib_c_s32lo16.so 100
run lib_d_lo16_w32.so 100
run lib_a_base.so "" --concurrent 3
run lib_d_lo16_w32.so 100 --concurrent 3
run lib_a_base.so ""
