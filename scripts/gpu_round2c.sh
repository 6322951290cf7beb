cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_new.log
bash scripts/gpu_sweep_libs.sh 4k10_n15
bash scripts/gpu_inst.sh 4k10_n15 gpurun_ab/lib_f2_b6.so 2>&1 | grep -v "^==" | tail -3
