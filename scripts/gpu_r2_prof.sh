# ncu --set full captures of the search kernels for the libraries given: gpu_r2_prof.sh name=lib ...
cd $GRAFT_REPO_ROOT
for a in "$@"; do
  n=${a%%=*}; lib=${a#*=}
  if [ "$lib" = cur ]; then unset TF_GPU_LIB; else export TF_GPU_LIB=$GRAFT_REPO_ROOT/$lib; fi
  for k in search16 search32; do
    bash scripts/gpu_prof_kernel.sh tf_$k prof_r02_${n}_$k 4k10_n15 20
  done
done
