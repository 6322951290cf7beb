cd $GRAFT_REPO_ROOT
WL=${1:-4k10_n15}
ncu --set full --clock-control none --import-source on -k regex:tf_search32 -s 1 -c 1 -o gpurun_out/prof_s32 -f python scripts/profile_step.py $WL 2 > gpurun_out/prof_s32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tf_search16 -s 1 -c 1 -o gpurun_out/prof_s16 -f python scripts/profile_step.py $WL 2 > gpurun_out/prof_s16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tf_filter -s 1 -c 1 -o gpurun_out/prof_flt -f python scripts/profile_step.py $WL 2 > gpurun_out/prof_flt.log 2>&1
tail -1 gpurun_out/prof_flt.log
