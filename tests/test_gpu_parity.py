"""GPU parity: CUDA path (through the C ABI) vs the oracle on identical seeded
inputs.  Bar: motion vectors, block errors and predictors bit-exact; filtered
pixels within +-1 LSB (float weights; in practice they are bit-exact too and the
mismatch count is reported)."""
import numpy as np
import pytest

import _clips
import _gpu
import _params

pytestmark = pytest.mark.gpu

CASES = [
    # (name, W, H, N, bit_depth, clip kwargs, param kwargs)
    ("cif8_speed4", 352, 288, 7, 8, {}, {}),
    ("cif10_speed4", 352, 288, 5, 10, {}, {}),
    ("qcif8_speed0", 176, 144, 5, 8, {}, dict(speed=0)),
    ("qcif10_speed0_hp", 176, 144, 3, 10, {}, dict(speed=0, allow_hp=1)),
    ("qcif8_speed3_q15", 176, 144, 3, 8, {}, dict(speed=3, q_factor=15)),
    ("odd8_speed2", 200, 120, 5, 8, dict(motion=(3, 5)), dict(speed=2)),
    ("i444_8", 130, 70, 3, 8, {}, dict(ss_x=0, ss_y=0)),
    ("i422_10_speed1", 130, 70, 3, 10, {}, dict(ss_x=1, ss_y=0, speed=1)),
    ("mono8", 131, 77, 4, 8, {}, dict(monochrome=1, filter_frame_idx=3)),
    ("bd12_speed3", 128, 96, 3, 12, {}, dict(speed=3, q_factor=15)),
    ("hbd8_q200", 128, 96, 3, 8, {}, dict(use_hbd=1, q_factor=200, filter_strength=2)),
    ("intmv8", 128, 96, 3, 8, {}, dict(force_integer_mv=1)),
    ("intmv10_speed0", 128, 96, 3, 10, {}, dict(force_integer_mv=1, speed=0)),
    ("hd720_8_skip", 1280, 720, 3, 8, {}, {}),
    ("hd720_10_skip", 1280, 720, 3, 10, dict(motion=(2, 7)), {}),
    ("tiny_40x24", 40, 24, 3, 8, {}, {}),
    ("tiny_17x33_10", 17, 33, 3, 10, {}, dict(speed=1)),
    ("i444_12_speed0", 72, 40, 3, 12, {}, dict(ss_x=0, ss_y=0, speed=0)),
    ("hd720_8_noskip_speed2", 736, 720, 2, 8, dict(motion=(1, 3)), dict(speed=2, use_downsampled_sad=0)),
    ("long_window_21", 96, 64, 21, 8, dict(motion=(0, 1)), dict(filter_frame_idx=10)),
    ("speed4_hp_8", 176, 144, 3, 8, dict(motion=(1, 3)), dict(allow_hp=1)),
    ("speed4_hp_10", 176, 144, 3, 10, dict(motion=(2, 1)), dict(allow_hp=1)),
    ("speed1_tree_iters2_8", 176, 144, 3, 8, dict(motion=(2, 3)), dict(speed=1)),
    ("pruned_iters2_hp", 176, 144, 3, 8, dict(motion=(1, 2)), dict(speed=3, subpel_iters_per_step=2, allow_hp=1)),
    ("pruned_more_iters2_10", 176, 144, 3, 10, dict(motion=(3, 1)), dict(subpel_iters_per_step=2)),
    ("strength0_q5", 128, 96, 3, 8, {}, dict(filter_strength=0, q_factor=5)),
    ("strength6_q255", 128, 96, 3, 10, {}, dict(filter_strength=6, q_factor=255)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_moving_texture(pkg, tfgpu, case):
    name, W, H, N, bd, ckw, pkw = case
    frames = _clips.moving_texture(W, H, N, bd, ss_x=pkw.get("ss_x", 1), ss_y=pkw.get("ss_y", 1),
                                   monochrome=pkw.get("monochrome", 0), **ckw)
    p = _params.tf_params(W, H, N, bit_depth=bd, **pkw)
    g = _gpu.run_gpu(pkg, tfgpu, p, frames)
    o = _gpu.oracle_run(p, frames)
    rep = _gpu.compare(g, o, p, tol_out=1)
    print(name, rep)
    assert rep["mvs"] == 0 and rep["mses"] == 0 and rep["pred"] == 0, rep
    assert rep["out_bad"] == 0 and rep["out_maxdiff"] <= 1, rep
    # mismatch fraction of the float-weighted output stays tiny
    assert rep["out_mismatch"] <= max(1, rep["out_total"] // 1000), rep


RANDOM_CASES = [
    ("rand8", 160, 96, 3, 8, None, {}),
    ("rand10_speed0", 160, 96, 3, 10, None, dict(speed=0)),
    ("rand8_720_kf", 1280, 720, 2, 8, None, dict(filter_frame_idx=0)),
    ("extreme8", 96, 64, 3, 8, 0, {}),
    ("extreme10", 96, 64, 3, 10, 1, {}),
]


@pytest.mark.parametrize("case", RANDOM_CASES, ids=[c[0] for c in RANDOM_CASES])
def test_random_and_extreme(pkg, tfgpu, case):
    """Uniform-random frames drive the diamond far from the start, so the mesh search and the
    skip-SAD re-run are exercised; extreme frames follow test/temporal_filter_test.cc:165-187."""
    name, W, H, N, bd, extreme, pkw = case
    frames = _clips.random_frames(W, H, N, bd, seed=3, extreme=extreme)
    p = _params.tf_params(W, H, N, bit_depth=bd, **pkw)
    g = _gpu.run_gpu(pkg, tfgpu, p, frames)
    o = _gpu.oracle_run(p, frames)
    rep = _gpu.compare(g, o, p, tol_out=1)
    print(name, rep)
    assert rep["mvs"] == 0 and rep["mses"] == 0 and rep["pred"] == 0, rep
    assert rep["out_bad"] == 0, rep


def test_noise_estimate(pkg, tfgpu):
    import _oracle
    for bd, (W, H) in ((8, (352, 288)), (10, (200, 120)), (12, (64, 48))):
        frames = _clips.moving_texture(W, H, 1, bd)
        p = _params.tf_params(W, H, 1, bit_depth=bd)
        b = pkg.Yv12Buffer(W, H, 1, 1, p["use_hbd"], p["border"]).set_planes(*frames[0])
        o = _oracle.OracleFilter(p, frames)
        exp = o.estimate_noise(0)
        got = [tfgpu.estimate_noise_from_single_plane(b, pl, bd) for pl in range(3)]
        assert got == exp, (bd, got, exp)
