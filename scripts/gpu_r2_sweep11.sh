cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
TF_GPU_CHAIN=frames python scripts/slab_latency_probe.py 2>&1 | tail -1
python scripts/slab_latency_probe.py 2>&1 | tail -1
TF_GPU_CHAIN=fused python scripts/slab_latency_probe.py 2>&1 | tail -1
