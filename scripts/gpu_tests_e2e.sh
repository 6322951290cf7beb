# GPU suite + end-to-end aomenc comparison (stock vs CONFIG_TF_GPU=1) on one box
cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python scripts/aomenc_e2e.py --clips ${1:-cif8,cif10,hd8} --out gpurun_out/aomenc_e2e.json 2> gpurun_out/aomenc_e2e.log; echo "e2e rc=$?"; tail -4 gpurun_out/aomenc_e2e.log | cut -c1-600
