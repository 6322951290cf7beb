"""GPU: the drop-in boundary exercised from the reference's side.  oracle/_ref/libtf_ref_seam.so
is the reference's own temporal_filter.c with integration/tf_gpu_seam.patch applied (the
CONFIG_TF_GPU branch a maintainer adds to av1_temporal_filter) linked against libtf_gpu.so.
The shim fills tf_gpu_params / tf_gpu_frame from AV1_COMP, TemporalFilterCtx and the
lookahead entries exactly as in the patch; its output must equal the reference's CPU path."""
import numpy as np
import pytest

import _clips
import _params
import _ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not _ref.seam_available(), reason="oracle/_ref/libtf_ref_seam.so not built")]

CASES = [
    ("cif8_s4", 352, 288, 5, 8, {}),
    ("cif10_s4", 352, 288, 3, 10, {}),
    ("qcif8_s0", 176, 144, 3, 8, dict(speed=0)),
    ("qcif10_s3_lowq", 176, 144, 3, 10, dict(speed=3, q_factor=12)),
    ("odd_i444", 130, 70, 3, 8, dict(ss_x=0, ss_y=0)),
    ("hd8_skip", 1280, 720, 2, 8, {}),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_side_shim_matches_reference_cpu_path(case):
    name, W, H, N, bd, pkw = case
    frames = _clips.moving_texture(W, H, N, bd, ss_x=pkw.get("ss_x", 1), ss_y=pkw.get("ss_y", 1))
    p = _params.tf_params(W, H, N, bit_depth=bd, **pkw)
    cpu = _ref.RefFilter(p, frames)
    p["noise_levels"] = tuple(cpu.estimate_noise())
    cpu.close()
    cpu = _ref.RefFilter(p, frames)
    a = cpu.run(record=False)
    seam = _ref.RefFilter(p, frames, seam=True)
    b = seam.run_gpu_seam()
    mism = 0
    for x, y in zip(a["out"], b["out"]):
        # the seam writes through the host YV12 buffer: compare the crop area and the full blocks
        d = np.abs(x.astype(np.int32) - y.astype(np.int32))
        assert d.max() <= 1
        mism += int((d > 0).sum())
    assert mism == 0
    assert (a["diff"] == b["diff"]).all()
    cpu.close()
    seam.close()


@pytest.mark.parametrize("case", CASES[:5], ids=[c[0] for c in CASES[:5]])
def test_reference_side_noise_hunk_matches_reference(case):
    """The patch's second hunk: tf_setup_filtering_buffer() takes its noise levels from
    tf_gpu_estimate_noise(); they must be the doubles av1_estimate_noise_from_single_plane() returns."""
    name, W, H, N, bd, pkw = case
    frames = _clips.moving_texture(W, H, N, bd, ss_x=pkw.get("ss_x", 1), ss_y=pkw.get("ss_y", 1))
    p = _params.tf_params(W, H, N, bit_depth=bd, **pkw)
    seam = _ref.RefFilter(p, frames, seam=True)
    want = _ref.RefFilter(p, frames).estimate_noise()
    assert seam.gpu_noise_levels() == want
    seam.close()
