cd $GRAFT_REPO_ROOT
R=${1:-r01}
# launch list of the bench command (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --concurrent 1 > gpurun_out/launches_${R}.log 2>&1
tail -1 gpurun_out/launches_${R}.log | cut -c1-200
# one full capture per kernel (4K 10-bit N=15): the launch in the middle of the window
for k in tf_search32 tf_search16; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -o gpurun_out/prof_${R}_$k -f python scripts/profile_step.py 4k10_n15 1 > gpurun_out/prof_${R}_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:tf_filter -s 0 -c 1 -o gpurun_out/prof_${R}_tf_filter -f python scripts/profile_step.py 4k10_n15 1 > gpurun_out/prof_${R}_tf_filter.log 2>&1
ls -la gpurun_out/*.ncu-rep
