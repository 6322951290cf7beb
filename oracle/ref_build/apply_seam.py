#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Applies the temporal_filter.c hunks of integration/tf_gpu_seam.patch
to a scratch copy of the reference file (outside the repository) so that the seam can be
compiled into oracle/_ref/libtf_ref_seam.so.  usage: apply_seam.py <ref root> <patch> <out dir>"""
import os
import sys

ref, patch_path, out = sys.argv[1:4]
patch = open(patch_path).read().split("--- a/build/cmake")[0]
src = open(os.path.join(ref, "av1/encoder/temporal_filter.c")).read()
h1 = patch.split("@@ -41,6 +41,98 @@\n")[1].split("@@ -1291,6 +1383,11 @@")[0]
add1 = "".join(l[1:] + "\n" for l in h1.split("\n") if l.startswith("+"))
anchor1 = '#include "av1/encoder/temporal_filter.h"\n\n'
assert anchor1 in src
src = src.replace(anchor1, anchor1 + add1, 1)
h2 = patch.split("@@ -1291,6 +1383,11 @@\n")[1]
add2 = "".join(l[1:] + "\n" for l in h2.split("\n") if l.startswith("+"))
anchor2 = "              compute_frame_diff, output_frame);\n\n  // Allocate and reset temporal filter buffers."
assert anchor2 in src
src = src.replace(anchor2, "              compute_frame_diff, output_frame);\n\n" + add2 +
                  "  // Allocate and reset temporal filter buffers.", 1)
os.makedirs(os.path.join(out, "av1/encoder"), exist_ok=True)
open(os.path.join(out, "av1/encoder/temporal_filter.c"), "w").write(src)
