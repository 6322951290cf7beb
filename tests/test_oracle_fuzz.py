"""CPU: the same seeded random sweep, oracle vs the compiled reference (skipped without oracle/_ref)."""
import os

import pytest

import _clips
import _oracle
import _params
import _ref
from test_gpu_fuzz import _case

pytestmark = pytest.mark.skipif(not _ref.available(), reason="oracle/_ref/libtf_ref.so not built")


# TF_FUZZ_SEEDS=N widens the sweep (soak runs); 115/426/460: see tests/test_gpu_fuzz.py
@pytest.mark.parametrize("seed", sorted(set(range(int(os.environ.get("TF_FUZZ_SEEDS", "40")))) | {115, 426, 460}))
def test_random_configuration_oracle_vs_reference(seed):
    W, H, N, bd, kw, clip, random_frames = _case(seed)
    fk = dict(ss_x=kw["ss_x"], ss_y=kw["ss_y"], monochrome=kw["monochrome"])
    frames = (_clips.random_frames(W, H, N, bd, seed=clip["seed"], **fk) if random_frames
              else _clips.moving_texture(W, H, N, bd, **fk, **clip))
    p = _params.tf_params(W, H, N, bit_depth=bd, **kw)
    r, o = _ref.RefFilter(p, frames), _oracle.OracleFilter(p, frames)
    a, b = r.run(), o.run()
    for k in ("mvs", "mses", "pred", "diff"):
        assert (a[k] == b[k]).all(), (seed, k)
    for x, y in zip(a["out"], b["out"]):
        assert (x == y).all(), seed
    r.close()
    o.close()
