# A/B of an env toggle on the default workload: usage gpu_ab.sh VAR
cd $GRAFT_REPO_ROOT
V=${1:-TF_GPU_NO_SEA}
for i in 1 2; do
for t in 0 1; do
env $V=$t python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$V=$t', round(d['value'],2), d['roofline']['phases_ms'], round(d['roofline']['int']['executed']['work_per_block_ref']['sad']))"
done; done
