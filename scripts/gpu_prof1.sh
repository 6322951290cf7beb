cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --workload 1080p8_n7 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r01.log 2>&1
tail -2 gpurun_out/launches_r01.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:tf_block -s 1 -c 1 -o gpurun_out/prof_r01_1080p8 -f python scripts/profile_step.py 1080p8_n7 2 > gpurun_out/prof_r01.log 2>&1
tail -3 gpurun_out/prof_r01.log
ls -la gpurun_out/
