"""GPU: the configuration bench.py times -- several tf_gpu contexts in flight on one device through
tf_gpu_filter_resident_async / _result (2 at 4K, 4 at 1080p; SURVEY 8e config 5: independent ARF windows) -- must
give, for every context and every round of the staggered pipeline, exactly the frame the synchronous public call
tf_gpu_filter() returns for the same window on an otherwise idle device, and the oracle's rows."""
import numpy as np
import pytest

import _clips
import _oracle
import _params

pytestmark = pytest.mark.gpu

CASES = [  # name, width, height, bit depth, frames, strength, contexts in flight, windows per context
    ("4k10_n15_x2x4", 3840, 2160, 10, 15, 5, (2, 4), 1),
    ("1080p10_n11_x4", 1920, 1080, 10, 11, 5, (4,), 2),
    ("1080p8_n7_x4", 1920, 1080, 8, 7, 4, (2, 4), 2),
]


def _window(pkg, ctx, p, W, H, bd, N, seed, id_base):
    frames = _clips.moving_texture(W, H, N, bd, seed=seed)
    bufs = []
    for i, (y, u, v) in enumerate(frames):
        b = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"], frame_id=id_base + i)
        bufs.append(b.set_planes(y, u, v, extend=False))
        ctx.cache_frame(b)
    return frames, bufs


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_concurrent_contexts_equal_the_synchronous_call(pkg, case):
    name, W, H, bd, N, strength, concs, nwin = case
    p = _params.tf_params(W, H, N, bit_depth=bd, q_factor=32, filter_strength=strength, allow_hp=1)
    K = max(concs)
    ctxs = [pkg.TemporalFilterGpu(max_cached_frames=40) for _ in range(K)]
    wins = [[_window(pkg, ctxs[ci], p, W, H, bd, N, seed=500 + 10 * ci + w, id_base=1 + 100 * w) for w in range(nwin)]
            for ci in range(K)]
    fi = p["filter_frame_idx"]
    p["noise_levels"] = tuple(ctxs[0].estimate_noise_from_single_plane(wins[0][0][1][fi], pl, bd) for pl in range(3))
    mb_rows = (H + 31) // 32

    # the synchronous public call, one window at a time on an idle device: the expected frames
    want = {}
    for ci in range(K):
        sync = pkg.TemporalFilterGpu(max_cached_frames=24)  # per window set: the frame ids repeat across contexts
        for w in range(nwin):
            out = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"])
            r = sync.temporal_filter(p, [b for b in wins[ci][w][1]], out)
            want[ci, w] = ([out.full_blocks(pl).copy() for pl in range(3)], r["diff"].copy())
        sync.close()

    # the oracle on one block row of one window pins the expected frames themselves
    o = _oracle.OracleFilter(p, wins[K - 1][nwin - 1][0])
    mid = mb_rows // 2
    ref = o.run(record=False, rows=(mid, mid + 1))
    o.close()
    for pl in range(3):
        bh = 32 >> (1 if pl else 0)
        assert (ref["out"][pl][mid * bh:(mid + 1) * bh] == want[K - 1, nwin - 1][0][pl][mid * bh:(mid + 1) * bh]).all()

    def check(ci, w):
        out = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"])
        ctxs[ci].download_output(out)
        for pl in range(3):
            assert (out.full_blocks(pl) == want[ci, w][0][pl]).all(), (name, ci, w, pl)

    for conc in concs:
        # bench.py's staggered pipeline: a context gets its next window as soon as its previous one is harvested
        pending = [None] * conc
        rounds = 3 * nwin
        for k in range(rounds):
            for ci in range(conc):
                if pending[ci] is not None:
                    _, diff = ctxs[ci].filter_resident_result()
                    assert (diff == want[ci, pending[ci]][1]).all(), (name, conc, ci, k)
                    check(ci, pending[ci])
                w = k % nwin
                ctxs[ci].filter_resident_async(p, [b.frame_id for b in wins[ci][w][1]])
                pending[ci] = w
        for ci in range(conc):
            _, diff = ctxs[ci].filter_resident_result()
            assert (diff == want[ci, pending[ci]][1]).all()
            check(ci, pending[ci])
    for c in ctxs:
        c.close()


def test_pipelined_submits_on_several_contexts_match(pkg):
    """bench.py's end-to-end leg: host buffers through tf_gpu_submit / tf_gpu_wait, two windows queued per context,
    two contexts staggered (1080p 10-bit, 7 frames): every output equals the synchronous call's."""
    W, H, bd, N = 1920, 1080, 10, 7
    p = _params.tf_params(W, H, N, bit_depth=bd, allow_hp=1)
    ctxs = [pkg.TemporalFilterGpu(max_cached_frames=24) for _ in range(2)]
    wins = []
    for ci in range(2):
        frames = _clips.moving_texture(W, H, N, bd, seed=900 + ci)
        wins.append([pkg.Yv12Buffer(W, H, 1, 1, True, p["border"]).set_planes(y, u, v, extend=False) for (y, u, v) in frames])
    sync = pkg.TemporalFilterGpu()
    want = []
    for ci in range(2):
        out = pkg.Yv12Buffer(W, H, 1, 1, True, p["border"])
        r = sync.temporal_filter(p, wins[ci], out)
        want.append(([out.full_blocks(pl).copy() for pl in range(3)], r["diff"].copy()))
    sync.close()
    outs = [[pkg.Yv12Buffer(W, H, 1, 1, True, p["border"]) for _ in range(2)] for _ in range(2)]
    queues = [[], []]
    done = 0
    for k in range(6):
        for ci in range(2):
            if len(queues[ci]) == 2:
                t, diff, _keep, o = queues[ci].pop(0)
                ctxs[ci].wait(t)
                assert [diff[0], diff[1]] == list(want[ci][1])
                for pl in range(3):
                    assert (o.full_blocks(pl) == want[ci][0][pl]).all()
                done += 1
            o = outs[ci][k % 2]
            t, diff, keep = ctxs[ci].submit(p, wins[ci], o)
            queues[ci].append((t, diff, keep, o))
    for ci in range(2):
        for t, diff, _keep, o in queues[ci]:
            ctxs[ci].wait(t)
            assert [diff[0], diff[1]] == list(want[ci][1])
            for pl in range(3):
                assert (o.full_blocks(pl) == want[ci][0][pl]).all()
            done += 1
    assert done == 12
    for c in ctxs:
        c.close()
