"""Small 8-bit and 10-bit windows through the public filter call (uploads, search, filter, read-back,
pipelined submits), a row-range resident call and the batch search, for compute-sanitizer runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest, _clips, _params
pkg = conftest.load_package()
ctx = pkg.TemporalFilterGpu()
for (W, H, bd, N, kw) in [(176, 144, 8, 3, {}), (208, 120, 10, 3, {"use_downsampled_sad": 1}), (96, 64, 12, 3, {"speed": 0})]:
    frames = _clips.moving_texture(W, H, N, bd)
    p = _params.tf_params(W, H, N, bit_depth=bd, **kw)
    bufs = []
    for (y, u, v) in frames:
        b = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"])
        bufs.append(b.set_planes(y, u, v, extend=False))
    outs = [pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"]) for _ in range(3)]
    tk = [ctx.submit(p, bufs, o) for o in outs]
    for t in tk:
        ctx.wait(t[0])
    print(W, H, bd, "diff", list(tk[0][1]), list(tk[2][1]))
# the resident calls: a block-row slab (one search32 launch over all frames: the latency mode) and the batch search
W, H, bd, N = 208, 120, 10, 3
frames = _clips.moving_texture(W, H, N, bd)
p = _params.tf_params(W, H, N, bit_depth=bd, use_downsampled_sad=1)
bufs = [pkg.Yv12Buffer(W, H, 1, 1, True, p["border"], frame_id=7000 + i).set_planes(*f, extend=False) for i, f in enumerate(frames)]
for b in bufs:
    ctx.cache_frame(b)
ids = [b.frame_id for b in bufs]
print("slab rows [1,3)", ctx.filter_resident(dict(p, out_row_begin=1, out_row_end=3), ids)[1].tolist(),
      "full", ctx.filter_resident(p, ids)[1].tolist())
print("batch", ctx.fullpel_search_batch(p, bufs[0], bufs[1], 16, [(0, 0, 0, 0), (16, 16, 1, -2), (192, 104, 3, 3)]).tolist())
ctx.close()
