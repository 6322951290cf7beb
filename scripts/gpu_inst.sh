# Per-kernel warp-instruction counts and durations (ncu, one window) for each library given.
# usage: gpu_inst.sh [workload] lib...   ("cur" = the in-tree build)
cd $GRAFT_REPO_ROOT
WL=${1:-4k10_n15}; shift
LIBS="$@"; [ -z "$LIBS" ] && LIBS="gpurun_ab/libtf_gpu_r01.so cur"
for v in $LIBS; do
if [ $v = cur ]; then unset TF_GPU_LIB; else export TF_GPU_LIB=$GRAFT_REPO_ROOT/$v; fi
n=$(basename $v .so)
timeout 900 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tf_ -s 60 -c 31 --csv --log-file gpurun_out/inst_$n.csv python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --concurrent 1 > /dev/null 2> gpurun_out/inst_$n.err
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/inst_$n.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
agg={}
for r in rows[1:]:
    k=r[ki].split('(')[0]; a=agg.setdefault(k,{}); a.setdefault(r[mi],[]).append(float(r[vi].replace(',','')))
for k,a in agg.items():
    print('$n', k, 'launches', len(a['gpu__time_duration.sum']), 'Minst/launch', round(sum(a['smsp__inst_executed.sum'])/len(a['smsp__inst_executed.sum'])/1e6,1), 'us/launch', round(sum(a['gpu__time_duration.sum'])/len(a['gpu__time_duration.sum'])/1e3,1), 'issue%', round(sum(a['smsp__issue_active.avg.pct_of_peak_sustained_active'])/len(a['gpu__time_duration.sum']),1))
PY
done
