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
        # tf_setup_filtering_buffer(): the noise loop runs on the device
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
        # av1_temporal_filter(): the synchronous call (second-ARF caller, encode_strategy.c:819)
        ("              compute_frame_diff, output_frame);\n\n  // Allocate and reset temporal filter buffers.",
         "              compute_frame_diff, output_frame);\n\n"
         "#if CONFIG_TF_GPU\n  tf_gpu_do_filtering(cpi, frame_diff);\n  return;\n#endif\n\n"
         "  // Allocate and reset temporal filter buffers."),
        # av1_tf_info_free(): the device context dies with the TEMPORAL_FILTER_INFO that owns it
        ("void av1_tf_info_free(TEMPORAL_FILTER_INFO *tf_info) {\n"
         "  if (tf_info->is_temporal_filter_on == 0) return;\n",
         "void av1_tf_info_free(TEMPORAL_FILTER_INFO *tf_info) {\n"
         "#if CONFIG_TF_GPU\n"
         "  tf_gpu_release(tf_info);\n"
         "#endif\n"
         "  if (tf_info->is_temporal_filter_on == 0) return;\n"),
        # av1_tf_info_filtering(): KF and ARF windows submitted back to back, waited for together,
        # aom_extend_frame_borders() done on the device
        ("  const AV1_COMMON *const cm = &cpi->common;\n"
         "  for (int gf_index = 0; gf_index < gf_group->size; ++gf_index) {\n"
         "    int update_type = gf_group->update_type[gf_index];\n",
         "  const AV1_COMMON *const cm = &cpi->common;\n"
         "#if CONFIG_TF_GPU\n"
         "  uint64_t gpu_ticket[TF_INFO_BUF_COUNT];\n"
         "  int64_t gpu_diff[TF_INFO_BUF_COUNT][2];\n"
         "  int gpu_pending[TF_INFO_BUF_COUNT] = { 0 };\n"
         "  struct aom_usec_timer gpu_timer;\n"
         "  aom_usec_timer_start(&gpu_timer);\n"
         "  (void)cm;\n"
         "#endif\n"
         "  for (int gf_index = 0; gf_index < gf_group->size; ++gf_index) {\n"
         "    int update_type = gf_group->update_type[gf_index];\n"),
        ("        YV12_BUFFER_CONFIG *out_buf = &tf_info->tf_buf[buf_idx];\n"
         "        av1_temporal_filter(cpi, lookahead_idx, gf_index,\n"
         "                            &tf_info->frame_diff[buf_idx], out_buf);\n"
         "        aom_extend_frame_borders(out_buf, av1_num_planes(cm));\n",
         "        YV12_BUFFER_CONFIG *out_buf = &tf_info->tf_buf[buf_idx];\n"
         "#if CONFIG_TF_GPU\n"
         "        if (gpu_pending[buf_idx]) {  // the buffer is about to be rewritten\n"
         "          tf_gpu_wait_filtering(cpi, gpu_ticket[buf_idx], gpu_diff[buf_idx],\n"
         "                                &tf_info->frame_diff[buf_idx]);\n"
         "          gpu_pending[buf_idx] = 0;\n"
         "        }\n"
         "        init_tf_ctx(cpi, lookahead_idx, gf_index, 1, out_buf);\n"
         "        gpu_ticket[buf_idx] = tf_gpu_submit_filtering(cpi, gpu_diff[buf_idx]);\n"
         "        gpu_pending[buf_idx] = 1;\n"
         "#else\n"
         "        av1_temporal_filter(cpi, lookahead_idx, gf_index,\n"
         "                            &tf_info->frame_diff[buf_idx], out_buf);\n"
         "        aom_extend_frame_borders(out_buf, av1_num_planes(cm));\n"
         "#endif\n"),
        ("        tf_info->tf_buf_valid[buf_idx] = 1;\n"
         "      }\n"
         "    }\n"
         "  }\n"
         "}\n",
         "        tf_info->tf_buf_valid[buf_idx] = 1;\n"
         "      }\n"
         "    }\n"
         "  }\n"
         "#if CONFIG_TF_GPU\n"
         "  for (int i = 0; i < TF_INFO_BUF_COUNT; ++i) {\n"
         "    if (gpu_pending[i])\n"
         "      tf_gpu_wait_filtering(cpi, gpu_ticket[i], gpu_diff[i],\n"
         "                            &tf_info->frame_diff[i]);\n"
         "  }\n"
         "  aom_usec_timer_mark(&gpu_timer);\n"
         "  tf_seam_us_filter += aom_usec_timer_elapsed(&gpu_timer);\n"
         "#endif\n"
         "}\n"),
    ],
    "av1/encoder/temporal_filter.h": [
        ("typedef struct TEMPORAL_FILTER_INFO {\n",
         "#if CONFIG_TF_GPU\n"
         "struct tf_gpu_ctx;\n"
         "/*!\\brief Frame allocations page-locked for the device (lookahead slots + tf_buf). */\n"
         "#define TF_GPU_MAX_PINNED 64\n"
         "#endif\n"
         "typedef struct TEMPORAL_FILTER_INFO {\n"
         "#if CONFIG_TF_GPU\n"
         "  /*!\n"
         "   * B200 temporal filter context (tf_gpu.h): created on first use, destroyed\n"
         "   * by av1_tf_info_free().\n"
         "   */\n"
         "  struct tf_gpu_ctx *gpu;\n"
         "  /*!\n"
         "   * Host allocations registered with the device context.\n"
         "   */\n"
         "  void *gpu_pinned[TF_GPU_MAX_PINNED];\n"
         "  /*!\n"
         "   * Number of entries of gpu_pinned in use.\n"
         "   */\n"
         "  int gpu_num_pinned;\n"
         "#endif\n"),
        ("/*!\\brief Check whether we should apply temporal filter at all.\n",
         "#if CONFIG_TF_GPU\n"
         "struct AV1_COMP;\n"
         "/*!\\brief Uploads the frame that just entered the lookahead to the device.\n"
         " * \\param[in]   cpi            Top level encoder instance structure\n"
         " */\n"
         "void av1_tf_gpu_lookahead_push(struct AV1_COMP *cpi);\n"
         "/*!\\brief av1_estimate_noise_from_single_plane() on the device.\n"
         " * \\param[in]   cpi            Top level encoder instance structure\n"
         " * \\param[in]   frame          Frame to estimate the noise of\n"
         " * \\param[in]   in_lookahead   Whether frame is a lookahead_entry::img\n"
         " * \\param[in]   plane          Plane index\n"
         " * \\param[in]   bit_depth      Bit depth of the samples\n"
         " * \\param[in]   edge_thresh    Edge threshold\n"
         " * \\return Noise level, -1.0 when it cannot be estimated.\n"
         " */\n"
         "double av1_tf_gpu_estimate_noise(const struct AV1_COMP *cpi,\n"
         "                                 const YV12_BUFFER_CONFIG *frame,\n"
         "                                 int in_lookahead, int plane, int bit_depth,\n"
         "                                 int edge_thresh);\n"
         "#endif\n\n"
         "/*!\\brief Check whether we should apply temporal filter at all.\n"),
    ],
    "av1/encoder/encoder.c": [
        # av1_receive_raw_frame(): upload at lookahead push
        ('                       "av1_lookahead_push() failed");\n'
         "    res = -1;\n"
         "  }\n",
         '                       "av1_lookahead_push() failed");\n'
         "    res = -1;\n"
         "  }\n"
         "#if CONFIG_TF_GPU && !CONFIG_REALTIME_ONLY\n"
         "  if (res == 0) av1_tf_gpu_lookahead_push(cpi);\n"
         "#endif\n"),
        # ALLINTRA noise synthesis level
        ("      cpi->oxcf.noise_level =\n"
         "          (float)(av1_estimate_noise_from_single_plane(\n"
         "                      sd, 0, cm->seq_params->bit_depth, 16) -\n"
         "                  0.1);\n",
         "#if CONFIG_TF_GPU\n"
         "      cpi->oxcf.noise_level =\n"
         "          (float)(av1_tf_gpu_estimate_noise(cpi, sd, 0, 0,\n"
         "                                            cm->seq_params->bit_depth, 16) -\n"
         "                  0.1);\n"
         "#else\n"
         "      cpi->oxcf.noise_level =\n"
         "          (float)(av1_estimate_noise_from_single_plane(\n"
         "                      sd, 0, cm->seq_params->bit_depth, 16) -\n"
         "                  0.1);\n"
         "#endif\n"),
    ],
    "av1/encoder/encode_strategy.c": [
        # key-frame filtering gate
        ("        const double y_noise_level = av1_estimate_noise_from_single_plane(\n"
         "            frame_input->source, 0, cm->seq_params->bit_depth,\n"
         "            NOISE_ESTIMATION_EDGE_THRESHOLD);\n",
         "#if CONFIG_TF_GPU\n"
         "        // frame_input->source is the lookahead entry's image (:1385)\n"
         "        const double y_noise_level = av1_tf_gpu_estimate_noise(\n"
         "            cpi, frame_input->source, 1, 0, cm->seq_params->bit_depth,\n"
         "            NOISE_ESTIMATION_EDGE_THRESHOLD);\n"
         "#else\n"
         "        const double y_noise_level = av1_estimate_noise_from_single_plane(\n"
         "            frame_input->source, 0, cm->seq_params->bit_depth,\n"
         "            NOISE_ESTIMATION_EDGE_THRESHOLD);\n"
         "#endif\n"),
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
         "    target_include_directories(aom_av1_encoder\n"
         "                               PRIVATE ${TF_GPU_ROOT}/include)\n"
         "    target_link_libraries(aom PRIVATE\n"
         "                          ${TF_GPU_ROOT}/aom-av1-psy_b200/libtf_gpu.so)\n"
         "  endif()\n\n"
         "  if(CONFIG_TUNE_VMAF)\n    find_package(PkgConfig)\n"),
    ],
}

# Timing-only patch for the CPU arm of the end-to-end comparison (scripts/aomenc_e2e.py): the same
# TF_SEAM_TIMING line from the unmodified CPU path (wall time inside av1_temporal_filter()); it changes
# no encoder decision.
TIMING_EDITS = {
    "av1/encoder/temporal_filter.c": [
        ('#include "av1/encoder/temporal_filter.h"\n\n',
         '#include "av1/encoder/temporal_filter.h"\n\n'
         '#include <stdio.h>\n#include <stdlib.h>\n#include "aom_ports/aom_timer.h"\n'
         'static int64_t tf_seam_us_filter = 0;\nstatic int tf_seam_windows = 0;\n\n'),
        ("  TemporalFilterData *tf_data = &cpi->td.tf_data;\n"
         "  const int compute_frame_diff = frame_diff != NULL;\n",
         "  TemporalFilterData *tf_data = &cpi->td.tf_data;\n"
         "  const int compute_frame_diff = frame_diff != NULL;\n"
         "  struct aom_usec_timer seam_timer;\n"
         "  aom_usec_timer_start(&seam_timer);\n"),
        ("  // Deallocate temporal filter buffers.\n"
         "  tf_dealloc_data(tf_data, is_highbitdepth);\n",
         "  // Deallocate temporal filter buffers.\n"
         "  tf_dealloc_data(tf_data, is_highbitdepth);\n"
         "  aom_usec_timer_mark(&seam_timer);\n"
         "  tf_seam_us_filter += aom_usec_timer_elapsed(&seam_timer);\n"
         "  ++tf_seam_windows;\n"),
        ("void av1_tf_info_free(TEMPORAL_FILTER_INFO *tf_info) {\n",
         "void av1_tf_info_free(TEMPORAL_FILTER_INFO *tf_info) {\n"
         '  if (getenv("TF_SEAM_TIMING") && tf_seam_windows)\n'
         '    fprintf(stderr, "TF_SEAM_TIMING impl=cpu windows=%d filter_ms=%.3f\\n",\n'
         "            tf_seam_windows, tf_seam_us_filter * 1e-3);\n"),
    ],
}


def write_patch(edit_set, name):
    shutil.rmtree(TMP, ignore_errors=True)
    out = []
    for rel, edits in edit_set.items():
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
    open(os.path.join(HERE, name), "w").write("".join(out))
    shutil.rmtree(TMP, ignore_errors=True)
    print("wrote", os.path.join(HERE, name))


write_patch(EDITS, "tf_gpu_seam.patch")
write_patch(TIMING_EDITS, "tf_timing_only.patch")
