cd $GRAFT_REPO_ROOT
R=${1:-r01}
for k in tf_search32 tf_search16; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_${R}_1080p8_$k -f python scripts/profile_step.py 1080p8_n7 1 > gpurun_out/prof_${R}_1080p8_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:tf_filter -s 0 -c 1 -o gpurun_out/prof_${R}_1080p8_tf_filter -f python scripts/profile_step.py 1080p8_n7 1 > gpurun_out/prof_${R}_1080p8_tf_filter.log 2>&1
ls -la gpurun_out/prof_${R}_1080p8_*.ncu-rep
