"""The AVX2-bound flavour of the compiled reference (oracle/_ref/libtf_ref_avx2.so, the CPU
*timing* baseline of bench.py, SURVEY.md 8d "CPU baseline (ii)") must compute what the generic-C
reference computes: bit-exact MVs / MSEs / predictors / FRAME_DIFF, filtered output within 1 LSB
(its weights go through a float approximation, av1/encoder/x86/temporal_filter_avx2.c)."""
import numpy as np
import pytest

import _clips
import _params
import _ref

pytestmark = pytest.mark.skipif(not (_ref.available() and _ref.avx2_available()),
                                reason="AVX2 reference flavour not built or host lacks AVX2")

CASES = [
    # w, h, bd, n, speed, extra
    (352, 288, 8, 7, 4, {}),
    (352, 288, 10, 7, 4, {}),
    (320, 192, 12, 5, 4, {}),
    (320, 192, 8, 5, 0, {}),
    (320, 192, 10, 5, 2, {"allow_hp": 1}),
    (200, 136, 8, 5, 3, {"ss_x": 0, "ss_y": 0}),
    (200, 136, 10, 5, 4, {"monochrome": 1}),
    (1280, 720, 10, 3, 4, {}),   # >= 720p: skip-row SAD kernels
    (1280, 720, 8, 3, 4, {}),
]


@pytest.mark.parametrize("w,h,bd,n,speed,extra", CASES)
def test_avx2_flavour_matches_generic_c(w, h, bd, n, speed, extra):
    clip_kw = {k: extra[k] for k in ("ss_x", "ss_y", "monochrome") if k in extra}
    frames = _clips.moving_texture(w, h, n, bd, **clip_kw)
    p = _params.tf_params(w, h, n, bit_depth=bd, speed=speed, **extra)
    a = _ref.RefFilter(p, frames).run()
    b = _ref.RefFilter(p, frames, avx2=True).run()
    assert np.array_equal(a["mvs"], b["mvs"])
    assert np.array_equal(a["mses"], b["mses"])
    assert np.array_equal(a["pred"], b["pred"])
    for x, y in zip(a["out"], b["out"]):
        assert np.abs(x.astype(np.int32) - y.astype(np.int32)).max() <= 1


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("TF_FUZZ_SEEDS", "30"))))
def test_avx2_flavour_matches_generic_c_random_configurations(seed):
    """Same seeded sweep as tests/test_gpu_fuzz.py (sizes, formats, bit depths, speed classes)."""
    from test_gpu_fuzz import _case
    W, H, N, bd, kw, clip, random_frames = _case(seed)
    fk = dict(ss_x=kw["ss_x"], ss_y=kw["ss_y"], monochrome=kw["monochrome"])
    frames = (_clips.random_frames(W, H, N, bd, seed=clip["seed"], **fk) if random_frames
              else _clips.moving_texture(W, H, N, bd, **fk, **clip))
    p = _params.tf_params(W, H, N, bit_depth=bd, **kw)
    ra, rb = _ref.RefFilter(p, frames), _ref.RefFilter(p, frames, avx2=True)
    a, b = ra.run(), rb.run()
    assert np.array_equal(a["mvs"], b["mvs"]) and np.array_equal(a["mses"], b["mses"])
    assert np.array_equal(a["pred"], b["pred"])
    for x, y in zip(a["out"], b["out"]):
        assert np.abs(x.astype(np.int32) - y.astype(np.int32)).max() <= 1
    ra.close()
    rb.close()
