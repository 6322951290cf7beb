#!/bin/bash
# Builds the reference's aomenc out of tree (generic C target = the av1_*_c parity target), three flavours:
#   integration/_build/aomenc_stock        the unmodified reference
#   integration/_build/aomenc_stock_timed  reference + integration/tf_timing_only.patch (prints the wall time
#                                          spent in av1_temporal_filter(); changes no encoder decision)
#   integration/_build/aomenc_tfgpu        reference + integration/tf_gpu_seam.patch, -DCONFIG_TF_GPU=1, linked
#                                          against aom-av1-psy_b200/libtf_gpu.so (runpath /root/repo/aom-av1-psy_b200,
#                                          which is where the repository sits here and on the GPU box)
# Sources are only copied to /tmp (the reference tree is read-only); nothing of it enters the repository.
# integration/_build/ is git-ignored but travels to the GPU box.   usage: build_aomenc.sh [--all]
set -e
REF=${REF:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/integration/_build
mkdir -p $OUT
CM="-G Ninja -DAOM_TARGET_CPU=generic -DENABLE_DOCS=0 -DENABLE_TESTS=0 -DENABLE_TOOLS=0 -DENABLE_EXAMPLES=1 -DCMAKE_BUILD_TYPE=Release"
if [ ! -x $OUT/aomenc_stock ] || [ "$1" = "--all" ]; then
  mkdir -p /tmp/aom_stock && cd /tmp/aom_stock
  cmake $REF $CM > cmake.log 2>&1 && ninja aomenc aomdec > ninja.log 2>&1
  cp aomenc $OUT/aomenc_stock && cp aomdec $OUT/aomdec
fi
if [ ! -x $OUT/aomenc_stock_timed ] || [ "$1" = "--all" ]; then
  rm -rf /tmp/aom_timed_src && cp -r $REF /tmp/aom_timed_src && chmod -R u+w /tmp/aom_timed_src
  (cd /tmp/aom_timed_src && patch -p1 -s < $ROOT/integration/tf_timing_only.patch)
  mkdir -p /tmp/aom_timed && cd /tmp/aom_timed
  cmake /tmp/aom_timed_src $CM > cmake.log 2>&1 && ninja aomenc > ninja.log 2>&1
  cp aomenc $OUT/aomenc_stock_timed
fi
rm -rf /tmp/aom_gpu_src && cp -r $REF /tmp/aom_gpu_src && chmod -R u+w /tmp/aom_gpu_src
(cd /tmp/aom_gpu_src && patch -p1 -s < $ROOT/integration/tf_gpu_seam.patch)
mkdir -p /tmp/aom_gpu && cd /tmp/aom_gpu
cmake /tmp/aom_gpu_src $CM -DCONFIG_TF_GPU=1 -DTF_GPU_ROOT=$ROOT > cmake.log 2>&1
ninja aomenc > ninja.log 2>&1 || { grep -A12 FAILED ninja.log | head -60; exit 1; }
cp aomenc $OUT/aomenc_tfgpu
ls -la $OUT
