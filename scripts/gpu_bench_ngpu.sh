# usage: gpu_bench_ngpu.sh N [slab]  -- the driver's launch line for N GPUs
cd $GRAFT_REPO_ROOT
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu_windows.json 2> gpurun_out/bench_${N}gpu_windows.err; tail -2 gpurun_out/bench_${N}gpu_windows.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_windows.json')); print('windows', d['n_gpus'], round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['scaling'], d['clocks'])"
if [ "$2" = "slab" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 --mode slab > gpurun_out/bench_${N}gpu_slab.json 2> gpurun_out/bench_${N}gpu_slab.err; tail -2 gpurun_out/bench_${N}gpu_slab.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_slab.json')); print('slab', d['n_gpus'], round(d['value'],2), 'fps', round(d['ms_per_step'],2), 'ms', d['scaling'])"
fi
