# new tests + default bench + 4K end-to-end aomenc on one box
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_concurrent.py tests/test_gpu_api.py tests/test_gpu_seam.py -x -q -m gpu > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_new.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
python scripts/aomenc_e2e.py --clips cif8,cif10,hd8,hd10,uhd10 --out gpurun_out/aomenc_e2e.json 2> gpurun_out/aomenc_e2e.log; echo "e2e rc=$?"
