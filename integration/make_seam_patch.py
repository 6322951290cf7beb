#!/usr/bin/env python3
"""Regenerates integration/tf_gpu_seam.patch as a well-formed unified diff (applies with `patch -p1` at the
root of the reference tree).  The edits are described here as (anchor, inserted text); scratch copies of the
three reference files are edited under /tmp and diffed -- no reference source is kept in this repository.
usage: make_seam_patch.py [/root/reference]"""
import os
import shutil
import subprocess
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
TMP = "/tmp/tf_gpu_seam_patchgen"

SHIM = open(os.path.join(HERE, "tf_gpu_seam_shim.c.inc")).read()

EDITS = {
    "av1/encoder/temporal_filter.c": [
        # (anchor, replacement)
        ('#include "av1/encoder/temporal_filter.h"\n\n',
         '#include "av1/encoder/temporal_filter.h"\n\n' + SHIM + "\n"),
        ("  double *noise_levels = tf_ctx->noise_levels;\n"
         "  for (int plane = 0; plane < num_planes; ++plane) {\n"
         "    noise_levels[plane] = av1_estimate_noise_from_single_plane(\n"
         "        to_filter_frame, plane, cpi->common.seq_params->bit_depth,\n"
         "        NOISE_ESTIMATION_EDGE_THRESHOLD);\n"
         "  }\n",
         "  double *noise_levels = tf_ctx->noise_levels;\n"
         "#if CONFIG_TF_GPU\n"
         "  (void)to_filter_frame;\n"
         "  tf_gpu_noise_levels(cpi, to_filter_buf, noise_levels);\n"
         "#else\n"
         "  for (int plane = 0; plane < num_planes; ++plane) {\n"
         "    noise_levels[plane] = av1_estimate_noise_from_single_plane(\n"
         "        to_filter_frame, plane, cpi->common.seq_params->bit_depth,\n"
         "        NOISE_ESTIMATION_EDGE_THRESHOLD);\n"
         "  }\n"
         "#endif\n"),
        ("              compute_frame_diff, output_frame);\n\n  // Allocate and reset temporal filter buffers.",
         "              compute_frame_diff, output_frame);\n\n"
         "#if CONFIG_TF_GPU\n  tf_gpu_do_filtering(cpi, frame_diff);\n  return;\n#endif\n\n"
         "  // Allocate and reset temporal filter buffers."),
    ],
    "build/cmake/aom_config_defaults.cmake": [
        ('set_aom_config_var(CONFIG_TUNE_VMAF 0 "Enable encoding tuning for VMAF.")\n',
         'set_aom_config_var(CONFIG_TUNE_VMAF 0 "Enable encoding tuning for VMAF.")\n'
         'set_aom_config_var(CONFIG_TF_GPU 0\n'
         '                   "Run the temporal filter on a B200 through libtf_gpu.so.")\n'),
    ],
    "CMakeLists.txt": [
        ("  if(CONFIG_TUNE_VMAF)\n    find_package(PkgConfig)\n",
         "  if(CONFIG_TF_GPU)\n"
         "    # -DCONFIG_TF_GPU=1 -DTF_GPU_ROOT=/path/to/tf-gpu\n"
         "    target_include_directories(aom PRIVATE ${TF_GPU_ROOT}/include)\n"
         "    target_link_libraries(aom PRIVATE\n"
         "                          ${TF_GPU_ROOT}/aom-av1-psy_b200/libtf_gpu.so)\n"
         "  endif()\n\n"
         "  if(CONFIG_TUNE_VMAF)\n    find_package(PkgConfig)\n"),
    ],
}

shutil.rmtree(TMP, ignore_errors=True)
out = []
for rel, edits in EDITS.items():
    for side in ("a", "b"):
        os.makedirs(os.path.dirname(os.path.join(TMP, side, rel)), exist_ok=True)
    src = open(os.path.join(REF, rel)).read()
    open(os.path.join(TMP, "a", rel), "w").write(src)
    for anchor, repl in edits:
        assert src.count(anchor) == 1, (rel, anchor[:60], src.count(anchor))
        src = src.replace(anchor, repl)
    open(os.path.join(TMP, "b", rel), "w").write(src)
    r = subprocess.run(["diff", "-u", "--label", "a/" + rel, "--label", "b/" + rel,
                        os.path.join(TMP, "a", rel), os.path.join(TMP, "b", rel)], capture_output=True, text=True)
    assert r.returncode == 1, r.stderr
    out.append(r.stdout)
open(os.path.join(HERE, "tf_gpu_seam.patch"), "w").write("".join(out))
shutil.rmtree(TMP, ignore_errors=True)
print("wrote", os.path.join(HERE, "tf_gpu_seam.patch"))
