cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_search_batch.py -x -q -m gpu > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_batch.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','verified')})
print(json.dumps(d.get('slab'), indent=1))
PY
