"""Executed-work counters of one window for the library in TF_GPU_LIB (or the in-tree build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, _params
pkg = bench.load_package()
wl = sys.argv[1] if len(sys.argv) > 1 else "1080p10_n11"
width, height, bd, n, strength = bench.WORKLOADS[wl]
frames = bench.make_window(width, height, bd, n, 77)
p = _params.tf_params(width, height, n, bit_depth=bd, q_factor=32, filter_strength=strength)
ctx = pkg.TemporalFilterGpu(device=0, max_cached_frames=40)
bufs = []
for i, (y, u, v) in enumerate(frames):
    b = pkg.Yv12Buffer(width, height, 1, 1, bd > 8, p["border"], frame_id=1 + i)
    b.set_planes(y, u, v, extend=False)
    bufs.append(b)
    ctx.cache_frame(b)
ids = [b.frame_id for b in bufs]
ctx.collect_counters(1)
ctx.filter_resident(p, ids)
c = ctx.read_counters()
nbr = ((height + 31) // 32) * ((width + 31) // 32) * (n - 1)
print(os.environ.get("TF_GPU_LIB", "cur"), wl, "per block-ref: sad px", c[0] / nbr, "subpel evals(1024px)", c[1] / nbr / 1024, "var evals(1024px)", c[2] / nbr / 1024, "c3", c[3] / nbr)
