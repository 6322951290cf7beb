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


def _cpu_reference(p, frames, extend=False):
    cpu = _ref.RefFilter(p, frames)
    a = cpu.run(record=False)
    planes = None
    if extend:
        cpu.extend_output_borders()
        planes = [cpu.plane_with_border(-1, pl)[0] for pl in range(cpu.num_planes)]
    cpu.close()
    return a, planes


@pytest.mark.parametrize("case", CASES[:5], ids=[c[0] for c in CASES[:5]])
def test_tf_info_filtering_hunk_submits_and_extends_on_device(case):
    """The av1_tf_info_filtering() hunk: tf_gpu_submit() for every window of the GOP, one wait at the end,
    aom_extend_frame_borders() (temporal_filter.c:1372) replaced by extend_output_borders = 1.  The WHOLE
    output allocation, borders included, must equal the reference's filtered + extended frame."""
    name, W, H, N, bd, pkw = case
    frames = _clips.moving_texture(W, H, N, bd, ss_x=pkw.get("ss_x", 1), ss_y=pkw.get("ss_y", 1))
    p = _params.tf_params(W, H, N, bit_depth=bd, **pkw)
    n0 = _ref.RefFilter(p, frames)
    p["noise_levels"] = tuple(n0.estimate_noise())
    n0.close()
    a, want = _cpu_reference(p, frames, extend=True)
    seam = _ref.RefFilter(p, frames, seam=True)
    b = seam.run_gpu_seam_async(copies=2)
    assert b["all_equal"] and (a["diff"] == b["diff"]).all()
    for pl in range(seam.num_planes):
        got, _ = seam.plane_with_border(-1, pl)
        assert np.array_equal(got, want[pl]), (name, pl)
    seam.close()


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[4]], ids=[CASES[i][0] for i in (0, 1, 4)])
def test_lookahead_push_hunk_uploads_at_push_time(case):
    """The av1_receive_raw_frame() hunk: frames go through a real av1_lookahead_push() and are uploaded (and
    their slots page-locked) by av1_tf_gpu_lookahead_push(); the filter call that follows finds every frame in
    the device cache (no upload kernels in its launch count) and returns the reference's result."""
    name, W, H, N, bd, pkw = case
    frames = _clips.moving_texture(W, H, N, bd, ss_x=pkw.get("ss_x", 1), ss_y=pkw.get("ss_y", 1))
    p = _params.tf_params(W, H, N, bit_depth=bd, **pkw)
    n0 = _ref.RefFilter(p, frames)
    p["noise_levels"] = tuple(n0.estimate_noise())
    n0.close()
    a, _ = _cpu_reference(p, frames)
    seam = _ref.RefFilter(p, frames, seam=True)
    assert not seam.L.tfref_gpu_has_context(seam.h)
    seam.gpu_push_window()
    assert seam.L.tfref_gpu_has_context(seam.h)
    assert seam.L.tfref_gpu_num_pinned(seam.h) == N  # one lookahead slot per pushed frame
    b = seam.run_gpu_seam()
    launches = seam.L.tfref_gpu_last_launches(seam.h)
    nplanes = seam.num_planes
    # search32 + search16 per reference frame, the filter kernel; an upload would add one border kernel per plane
    assert launches == 2 * (N - 1) + 1, (launches, N, nplanes)
    for x, y in zip(a["out"], b["out"]):
        assert np.array_equal(x, y)
    assert (a["diff"] == b["diff"]).all()
    # the context belongs to TEMPORAL_FILTER_INFO and dies in av1_tf_info_free()
    seam.L.tfref_gpu_release(seam.h)
    assert not seam.L.tfref_gpu_has_context(seam.h)
    assert seam.L.tfref_gpu_num_pinned(seam.h) == 0
    seam.close()


def test_other_noise_callers_route_through_the_device():
    """encode_strategy.c:746-750 (key-frame gate: luma of the lookahead frame, threshold 50) and
    encoder.c:4038-4051 (ALLINTRA noise synthesis: raw input, threshold 16) through av1_tf_gpu_estimate_noise()."""
    import ctypes as C
    for (W, H, bd) in [(352, 288, 8), (176, 144, 10)]:
        frames = _clips.moving_texture(W, H, 3, bd)
        p = _params.tf_params(W, H, 3, bit_depth=bd)
        seam = _ref.RefFilter(p, frames, seam=True)
        cpu = _ref.RefFilter(p, frames)
        want50 = cpu.estimate_noise(idx=1)[0]
        assert seam.gpu_estimate_noise(1, True, 0, 50) == want50
        assert seam.gpu_estimate_noise(1, False, 0, 50) == want50
        # threshold 16: compare with the library called directly (the CPU harness only exposes threshold 50)
        import conftest
        pkg = conftest.load_package()
        ctx = pkg.TemporalFilterGpu()
        b = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"]).set_planes(*frames[1])
        assert seam.gpu_estimate_noise(1, False, 0, 16) == ctx.estimate_noise_from_single_plane(b, 0, bd, 16)
        ctx.close()
        cpu.close()
        seam.close()
