# 2-GPU checks: the NCCL slab test and the driver-style bench launch (N = ${1:-2})
cd $GRAFT_REPO_ROOT
N=${1:-2}
python -m pytest tests/test_gpu_slab_nccl.py -x -q -m gpu > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_nccl.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','verified')}, d.get('e2e'))
print(json.dumps(d.get('slab'), indent=1))
PY
