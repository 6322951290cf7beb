cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:tf_block -s 1 -c 1 -o gpurun_out/prof_cur -f python scripts/profile_step.py ${1:-1080p8_n7} 2 > gpurun_out/prof_cur.log 2>&1
tail -2 gpurun_out/prof_cur.log
