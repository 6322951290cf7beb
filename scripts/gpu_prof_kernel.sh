# ncu --set full capture of one launch of one kernel.  usage: gpu_prof_kernel.sh <kernel regex> <out name> [workload] [skip]
cd $GRAFT_REPO_ROOT
K=$1; OUT=$2; WL=${3:-4k10_n15}; SKIP=${4:-20}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/$OUT python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --concurrent 1 > gpurun_out/$OUT.log 2>&1
tail -2 gpurun_out/$OUT.log | cut -c1-200
ls -la gpurun_out/$OUT.ncu-rep
