#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Applies integration/tf_gpu_seam.patch with patch(1) to scratch copies of the
reference files it touches (outside the repository), so that the patched temporal_filter.c can be
compiled into oracle/_ref/libtf_ref_seam.so -- and so that the patch is known to apply.
usage: apply_seam.py <ref root> <patch> <out dir>"""
import os
import re
import shutil
import subprocess
import sys

ref, patch_path, out = sys.argv[1:4]
patch_path = os.path.abspath(patch_path)
files = re.findall(r"^--- a/(\S+)$", open(patch_path).read(), re.M)
assert "av1/encoder/temporal_filter.c" in files
for rel in files:
    os.makedirs(os.path.dirname(os.path.join(out, rel)) or out, exist_ok=True)
    shutil.copyfile(os.path.join(ref, rel), os.path.join(out, rel))
subprocess.run(["patch", "-p1", "-s", "-i", patch_path], cwd=out, check=True)
