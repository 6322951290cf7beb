"""GPU: seeded random sweep of the parameter space (frame size, bit depth, chroma format, speed class,
q, strength, window length / centre, motion, noise) -- CUDA path vs oracle, same bars as the fixed cases."""
import os

import numpy as np
import pytest

import _clips
import _gpu
import _params

pytestmark = pytest.mark.gpu


def _case(seed):
    r = np.random.default_rng(1000 + seed)
    W = int(r.integers(16, 260))
    H = int(r.integers(16, 200))
    bd = int(r.choice([8, 8, 10, 10, 12]))
    fmt = [(1, 1, 0), (1, 1, 0), (1, 0, 0), (0, 0, 0), (1, 1, 1)][int(r.integers(0, 5))]
    N = int(r.integers(2, 8))
    kw = dict(
        ss_x=fmt[0], ss_y=fmt[1], monochrome=fmt[2],
        speed=int(r.integers(0, 5)), q_factor=int(r.choice([3, 18, 32, 90, 140, 255])),
        filter_strength=int(r.integers(0, 7)), filter_frame_idx=int(r.integers(0, N)),
        allow_hp=int(r.integers(0, 2)), force_integer_mv=int(r.random() < 0.1),
        use_hbd=1 if bd > 8 else int(r.random() < 0.3),
        noise_levels=tuple(float(x) for x in r.uniform(-1.0, 6.0, 3)),
    )
    if r.random() < 0.3:
        kw["use_downsampled_sad"] = 1  # skip-row SAD below 720p too
    if r.random() < 0.3:
        kw["subpel_iters_per_step"] = int(r.integers(1, 3))
    clip = dict(motion=(int(r.integers(0, 5)), int(r.integers(0, 7))), sigma=float(r.choice([0.5, 3.0, 12.0])) * (1 << (bd - 8)),
                seed=int(r.integers(0, 1 << 30)))
    random_frames = r.random() < 0.2
    return W, H, N, bd, kw, clip, random_frames


# TF_FUZZ_SEEDS=N widens the sweep for soak runs.  Seeds 115, 426 and 460 are the ones that exposed a
# sign error in the high-bitdepth variance rounding (vf(src, ref) vs vf(ref, src)) during development.
_SEEDS = sorted(set(range(int(os.environ.get("TF_FUZZ_SEEDS", "200")))) | {115, 426, 460})


@pytest.mark.parametrize("seed", _SEEDS)
def test_random_configuration(pkg, tfgpu, seed):
    W, H, N, bd, kw, clip, random_frames = _case(seed)
    fk = dict(ss_x=kw["ss_x"], ss_y=kw["ss_y"], monochrome=kw["monochrome"])
    frames = (_clips.random_frames(W, H, N, bd, seed=clip["seed"], **fk) if random_frames
              else _clips.moving_texture(W, H, N, bd, **fk, **clip))
    p = _params.tf_params(W, H, N, bit_depth=bd, **kw)
    g = _gpu.run_gpu(pkg, tfgpu, p, frames)
    o = _gpu.oracle_run(p, frames)
    rep = _gpu.compare(g, o, p, tol_out=1)
    assert rep["mvs"] == 0 and rep["mses"] == 0 and rep["pred"] == 0, (seed, W, H, N, bd, kw, rep)
    assert rep["accum"] == 0 and rep["count"] == 0, (seed, rep)  # weights are bit-exact in practice
    assert rep["out_bad"] == 0 and rep["diff_equal"], (seed, rep)
