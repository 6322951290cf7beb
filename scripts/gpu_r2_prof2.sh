cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
bash scripts/gpu_final_profiles.sh r02
