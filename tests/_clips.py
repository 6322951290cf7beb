"""Deterministic synthetic 'moving texture' clips (SURVEY.md 8d / BASELINE.md 3).

canvas (H+64)x(W+64) uniform ints in [0,2^bd) -> 5x5 box blur -> rescale to
[0.11,0.89]*(2^bd-1); frame n = canvas cropped at (dy,dx)=(n,2n) + N(0,sigma)
(sigma 3 for 8-bit, 12 for 10-bit), rounded, clipped; chroma = mid-grey +-
small uniform noise.  numpy default_rng, seeds 1234 (8-bit) / 77 (10-bit).
"""
import numpy as np


def _box5(a):
    a = a.astype(np.float64)
    c = np.cumsum(np.pad(a, ((5, 0), (0, 0))), axis=0)
    a = (c[5:] - c[:-5])
    c = np.cumsum(np.pad(a, ((0, 0), (5, 0))), axis=1)
    return (c[:, 5:] - c[:, :-5]) / 25.0


def moving_texture(width, height, num_frames, bit_depth=8, seed=None, ss_x=1, ss_y=1,
                   monochrome=False, motion=(1, 2), sigma=None, chroma_noise=None):
    """Returns list of (y,u,v) planes (u,v None if monochrome); dtype u8 or u16."""
    if seed is None:
        seed = 1234 if bit_depth == 8 else 77
    if sigma is None:
        sigma = 3.0 * (1 << (bit_depth - 8))
    if chroma_noise is None:
        chroma_noise = 2 << (bit_depth - 8)
    rng = np.random.default_rng(seed)
    maxv = (1 << bit_depth) - 1
    dt = np.uint8 if bit_depth == 8 else np.uint16
    pad = max(64, (abs(motion[0]) + abs(motion[1]) * 1) * num_frames + 8)
    canvas = rng.integers(0, 1 << bit_depth, size=(height + pad + 4, width + 2 * pad + 4))
    canvas = _box5(canvas)
    lo, hi = canvas.min(), canvas.max()
    canvas = (0.11 + 0.78 * (canvas - lo) / (hi - lo)) * maxv
    cw, ch = (width + ss_x) >> ss_x, (height + ss_y) >> ss_y
    frames = []
    for n in range(num_frames):
        dy, dx = motion[0] * n, motion[1] * n
        y = canvas[dy:dy + height, dx:dx + width] + rng.normal(0.0, sigma, size=(height, width))
        y = np.clip(np.rint(y), 0, maxv).astype(dt)
        if monochrome:
            frames.append((y, None, None))
            continue
        mid = 1 << (bit_depth - 1)
        u = (mid + rng.integers(-chroma_noise, chroma_noise + 1, size=(ch, cw))).astype(dt)
        v = (mid + rng.integers(-chroma_noise, chroma_noise + 1, size=(ch, cw))).astype(dt)
        frames.append((y, u, v))
    return frames


def random_frames(width, height, num_frames, bit_depth=8, seed=0, ss_x=1, ss_y=1, monochrome=False,
                  extreme=None):
    """Uniform-random or extreme-valued frames (test/temporal_filter_test.cc style)."""
    rng = np.random.default_rng(seed)
    maxv = (1 << bit_depth) - 1
    dt = np.uint8 if bit_depth == 8 else np.uint16
    cw, ch = (width + ss_x) >> ss_x, (height + ss_y) >> ss_y
    out = []
    for n in range(num_frames):
        def mk(h, w):
            if extreme is None:
                return rng.integers(0, maxv + 1, size=(h, w)).astype(dt)
            val = maxv if (extreme + n) % 2 == 0 else 0
            return np.full((h, w), val, dtype=dt)
        out.append((mk(height, width), None if monochrome else mk(ch, cw), None if monochrome else mk(ch, cw)))
    return out
