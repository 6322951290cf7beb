"""GPU: the bench workloads at their full sizes (BASELINE.json configs 2-4).  The oracle is too slow for
whole 4K windows, so full size is covered by (1) the oracle on sampled block rows -- top, middle and bottom
of the frame -- against the library's slab (row-range) mode, bit-exact on every plane and on FRAME_DIFF,
(2) size-independent properties: disjoint slabs reassemble the full-frame result exactly, FRAME_DIFF adds up,
repeated runs are identical, and a window of identical frames returns the frame itself."""
import numpy as np
import pytest

import _clips
import _oracle
import _params

pytestmark = pytest.mark.gpu

WORKLOADS = [  # name, width, height, bit depth, frames, strength, seed  (bench.py WORKLOADS)
    ("4k10_n15", 3840, 2160, 10, 15, 5, 77),
    ("1080p10_n11", 1920, 1080, 10, 11, 5, 77),
    ("1080p8_n7", 1920, 1080, 8, 7, 4, 1234),
]


def _bufs(pkg, p, frames):
    out = []
    for (y, u, v) in frames:
        b = pkg.Yv12Buffer(p["width"], p["height"], 1, 1, p["use_hbd"], p["border"])
        out.append(b.set_planes(y, u, v, extend=False))
    return out


@pytest.mark.parametrize("wl", WORKLOADS, ids=[w[0] for w in WORKLOADS])
def test_full_size_rows_against_oracle_and_slab_properties(pkg, tfgpu, wl):
    _, W, H, bd, N, strength, seed = wl
    frames = _clips.moving_texture(W, H, N, bd, seed=seed)
    p = _params.tf_params(W, H, N, bit_depth=bd, q_factor=32, filter_strength=strength)
    bufs = _bufs(pkg, p, frames)
    fi = p["filter_frame_idx"]
    p["noise_levels"] = tuple(tfgpu.estimate_noise_from_single_plane(bufs[fi], pl, bd) for pl in range(3))
    mb_rows = (H + 31) // 32

    full_out = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"])
    full = tfgpu.temporal_filter(p, bufs, full_out)
    full_planes = [full_out.full_blocks(pl).copy() for pl in range(3)]

    # (2) repeated run is identical
    again_out = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"])
    again = tfgpu.temporal_filter(p, bufs, again_out)
    assert (again["diff"] == full["diff"]).all()
    for pl in range(3):
        assert (again_out.full_blocks(pl) == full_planes[pl]).all()

    # (2) three disjoint slabs reassemble the full frame, FRAME_DIFF adds up
    slab_out = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"])
    cuts = [0, mb_rows // 3, 2 * mb_rows // 3 + 1, mb_rows]
    diff = np.zeros(2, np.int64)
    for b, e in zip(cuts[:-1], cuts[1:]):
        diff += tfgpu.temporal_filter(dict(p, out_row_begin=b, out_row_end=e), bufs, slab_out)["diff"]
    assert (diff == full["diff"]).all()
    for pl in range(3):
        assert (slab_out.full_blocks(pl) == full_planes[pl]).all()

    # (1) the oracle on sampled block rows
    o = _oracle.OracleFilter(p, frames)
    for b, e in ((0, 1), (mb_rows // 2, mb_rows // 2 + 1), (mb_rows - 1, mb_rows)):
        ref = o.run(record=False, rows=(b, e))
        got = tfgpu.temporal_filter(dict(p, out_row_begin=b, out_row_end=e), bufs, slab_out)
        assert (got["diff"] == ref["diff"]).all(), (b, e)
        for pl in range(3):
            bh = 32 >> (1 if pl else 0)
            want = ref["out"][pl][b * bh:e * bh]
            have = full_planes[pl][b * bh:e * bh].astype(np.uint16)
            assert want.shape == have.shape
            assert (want == have).all(), (pl, b, e, int((want != have).sum()))
    o.close()


def test_identical_frames_return_the_frame(pkg, tfgpu):
    """Every reference frame equals the frame to filter -> zero-error matches, and the weighted mean of
    identical samples is the sample (1080p 10-bit, 7 frames)."""
    W, H, bd, N = 1920, 1080, 10, 7
    one = _clips.moving_texture(W, H, 1, bd, seed=5)[0]
    frames = [one] * N
    p = _params.tf_params(W, H, N, bit_depth=bd)
    bufs = _bufs(pkg, p, frames)
    out = pkg.Yv12Buffer(W, H, 1, 1, True, p["border"])
    r = tfgpu.temporal_filter(p, bufs, out)
    assert list(r["diff"]) == [0, 0]
    for pl, src in enumerate(one):
        h, w = src.shape
        assert (out.full_blocks(pl)[:h, :w] == src).all()
