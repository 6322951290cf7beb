# A/B of two builds on the same box: aom-av1-psy_b200/libtf_gpu.so (B, current) vs gpurun_ab/libtf_gpu_a.so (A)
cd $GRAFT_REPO_ROOT
WL=${1:-4k10_n15}
for i in 1 2 3; do
for v in a b; do
if [ $v = a ]; then export TF_GPU_LIB=$GRAFT_REPO_ROOT/gpurun_ab/libtf_gpu_a.so; else unset TF_GPU_LIB; fi
python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],2), {k: round(x,2) for k,x in d['roofline']['phases_ms'].items()})"
done; done
