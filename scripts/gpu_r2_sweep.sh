# round-2 iteration: GPU suite on the in-tree build, then the A/B sweep of gpurun_ab/lib_*.so on the same box
cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-400
bash scripts/gpu_sweep_libs.sh ${1:-4k10_n15} 2>&1 | tee gpurun_out/sweep.log
