# driver-style bench launch on N GPUs (window-parallel line + slab section)
cd $GRAFT_REPO_ROOT
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','verified')}, d.get('e2e'))
s=d.get('slab') or {}
print({k:s.get(k) for k in ('ms_per_frame','single_gpu_ms_per_frame','speedup','efficiency','gather_ms','search32_chain_ms','verified')}, s.get('peer_store'))
PY
