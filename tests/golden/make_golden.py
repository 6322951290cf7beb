#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into
oracle/_ref/libtf_ref.so (run in the build container, where /root/reference exists):

    python tests/golden/make_golden.py

The fixtures pin the oracle (and through it the CUDA path) on the GPU box, where the
reference sources are absent.  Inputs are regenerated from seeds by tests/_clips.py; the
stored arrays are the reference's outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _clips  # noqa: E402
import _params  # noqa: E402
import _ref  # noqa: E402

# (name, W, H, N, bit_depth, clip kind, clip kwargs, param kwargs)
PIPELINE_CASES = [
    ("tiny8_s4", 96, 64, 3, 8, "moving", {}, {}),
    ("tiny10_s4", 96, 64, 3, 10, "moving", {}, {}),
    ("tiny8_s0", 96, 64, 3, 8, "moving", {}, dict(speed=0)),
    ("tiny10_s0_hp", 72, 40, 3, 10, "moving", dict(motion=(2, 3)), dict(speed=0, allow_hp=1)),
    ("tiny8_s3_lowq", 96, 64, 3, 8, "moving", {}, dict(speed=3, q_factor=15)),
    ("tiny12_s2", 64, 64, 3, 12, "moving", {}, dict(speed=2)),
    ("rand8", 96, 64, 3, 8, "random", {}, {}),
    ("rand10_s1", 96, 64, 2, 10, "random", {}, dict(speed=1, filter_frame_idx=0)),
    ("i444_8", 66, 38, 3, 8, "moving", {}, dict(ss_x=0, ss_y=0)),
    ("i422_10", 66, 38, 3, 10, "moving", {}, dict(ss_x=1, ss_y=0)),
    ("mono8", 67, 45, 3, 8, "moving", {}, dict(monochrome=1)),
    ("intmv8", 96, 64, 3, 8, "moving", {}, dict(force_integer_mv=1)),
    ("hd_skip8", 736, 720, 2, 8, "moving", dict(motion=(3, 5)), {}),
]


def make_frames(kind, W, H, N, bd, ckw, pkw):
    kw = dict(ss_x=pkw.get("ss_x", 1), ss_y=pkw.get("ss_y", 1), monochrome=pkw.get("monochrome", 0))
    if kind == "moving":
        return _clips.moving_texture(W, H, N, bd, **kw, **ckw)
    return _clips.random_frames(W, H, N, bd, seed=5, **kw, **ckw)


def apply_block_case(seed, bd, ss_x, ss_y, extreme=None):
    """test/temporal_filter_test.cc:130-221: one 32x32 block, fixed params."""
    import ctypes as C
    rng = np.random.default_rng(seed)
    use_hbd = bd > 8
    dt = np.uint16 if use_hbd else np.uint8
    maxv = (1 << bd) - 1
    W, H = 32, 32
    cw, ch = W >> ss_x, H >> ss_y
    stride, uvstride = 64, 64 >> ss_x  # block (0,0) inside a wider buffer

    def mk(h, w, st, flip=False):
        buf = np.zeros((h, st), dt)
        if extreme is None:
            buf[:, :w] = rng.integers(0, maxv + 1, size=(h, w))
        else:
            buf[:, :w] = maxv if (extreme ^ flip) else 0
        return buf
    sy, su, sv = mk(H, W, stride), mk(ch, cw, uvstride), mk(ch, cw, uvstride)
    num_pels = W * H + 2 * cw * ch
    if extreme is None:
        pred = rng.integers(0, maxv + 1, size=num_pels).astype(dt)
    else:
        pred = np.full(num_pels, 0 if extreme else maxv, dt)
    noise = np.array([2.1002103677063437] * 3)
    mvs = np.array([[0, 0], [5, 5], [7, 8], [2, 10]], np.int16)
    mses = np.array([15, 16, 17, 18], np.int32)
    accum = np.zeros(num_pels, np.uint32)
    count = np.zeros(num_pels, np.uint16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _ref.lib().tfref_apply_block(540, 360, ss_x, ss_y, 3, bd, int(use_hbd), p(sy), p(su), p(sv), stride, uvstride,
                                 0, 0, p(noise), p(mvs), p(mses), 12, 5, p(pred), p(accum), p(count))
    return dict(accum=accum, count=count)


def main():
    assert _ref.available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for name, W, H, N, bd, kind, ckw, pkw in PIPELINE_CASES:
        frames = make_frames(kind, W, H, N, bd, ckw, pkw)
        p = _params.tf_params(W, H, N, bit_depth=bd, **pkw)
        r = _ref.RefFilter(p, frames)
        noise = r.estimate_noise()
        p["noise_levels"] = tuple(noise) + (0.0,) * (3 - len(noise))
        r.close()
        r = _ref.RefFilter(p, frames)
        res = r.run()
        r.close()
        out[f"{name}/noise"] = np.array(p["noise_levels"])
        out[f"{name}/mvs"] = res["mvs"]
        out[f"{name}/mses"] = res["mses"]
        out[f"{name}/diff"] = res["diff"]
        for i, o in enumerate(res["out"]):
            out[f"{name}/out{i}"] = o
        # predictors are big: keep a checksum per (block, frame) instead
        out[f"{name}/pred_sum"] = res["pred"].astype(np.uint64).sum(axis=2)
        w = np.arange(1, res["pred"].shape[2] + 1, dtype=np.uint64)
        out[f"{name}/pred_wsum"] = (res["pred"].astype(np.uint64) * w).sum(axis=2)
        print(name, "diff", res["diff"])
    for bd in (8, 10):
        for (sx, sy) in ((1, 1), (1, 0), (0, 0)):
            for ex in (None, 0, 1):
                key = f"apply/bd{bd}_ss{sx}{sy}_ex{ex}"
                r = apply_block_case(0xbaba, bd, sx, sy, ex)
                out[key + "/accum"], out[key + "/count"] = r["accum"], r["count"]
    np.savez_compressed(os.path.join(HERE, "golden_r01.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_r01.npz"), os.path.getsize(os.path.join(HERE, "golden_r01.npz")), "bytes")


if __name__ == "__main__":
    main()
