"""One resident temporal-filter call per step for profiling under ncu (no timing claims)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import _params

wl = sys.argv[1] if len(sys.argv) > 1 else "1080p8_n7"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
width, height, bd, n, strength = bench.WORKLOADS[wl]
pkg = bench.load_package()
ctx = pkg.TemporalFilterGpu(device=0, max_cached_frames=40)
p = bench.window_params(wl)  # the parameters of the bench line (q, allow_hp from the aomenc run)
frames = bench.make_window(width, height, bd, n, 77 if bd > 8 else 1234)
bufs = []
for i, (y, u, v) in enumerate(frames):
    b = pkg.Yv12Buffer(width, height, 1, 1, bd > 8, p["border"], frame_id=1 + i)
    b.set_planes(y, u, v, extend=False)
    bufs.append(b)
fi = p["filter_frame_idx"]
p["noise_levels"] = tuple(ctx.estimate_noise_from_single_plane(bufs[fi], pl, bd) for pl in range(3))
for b in bufs:
    ctx.cache_frame(b)
for k in range(steps):
    ms, diff = ctx.filter_resident(p, [b.frame_id for b in bufs])
    print("step", k, "kernel ms", ms, "split", ctx.last_kernel_times(), "diff", diff.tolist())
