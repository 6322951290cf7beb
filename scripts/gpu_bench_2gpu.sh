cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_windows.json 2> gpurun_out/bench_2gpu_windows.err; tail -3 gpurun_out/bench_2gpu_windows.err; cut -c1-700 gpurun_out/bench_2gpu_windows.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --mode slab > gpurun_out/bench_2gpu_slab.json 2> gpurun_out/bench_2gpu_slab.err; tail -5 gpurun_out/bench_2gpu_slab.err; cut -c1-900 gpurun_out/bench_2gpu_slab.json
