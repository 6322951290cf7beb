"""One-off: rate of the library's frame upload path (2-D pinned copies + device border extension), 4K 10-bit."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest
pkg = conftest.load_package()
W, H = 3840, 2160
ctx = pkg.TemporalFilterGpu()
bufs = []
for i in range(16):
    b = pkg.Yv12Buffer(W, H, 1, 1, True, 160, frame_id=1000 + i)
    for a in b.alloc:
        a[:] = (i * 7) % 1000
        ctx.host_register(a)
    bufs.append(b)
nid = 2000
for r in range(3):  # fill every cache slot once so that no allocation happens in the timed part
    for b in bufs:
        b.frame_id = nid; nid += 1
        ctx.cache_frame(b)
t = time.perf_counter()
nid = 5000
reps = 3
for r in range(reps):
    for b in bufs:
        b.frame_id = nid; nid += 1
        ctx.cache_frame(b)
dt = time.perf_counter() - t
mb = W * H * 1.5 * 2 / 1e6
print(f"upload path: {reps * len(bufs)} frames of {mb:.1f} MB in {dt * 1e3:.1f} ms = {reps * len(bufs) * mb / dt / 1e3:.1f} GB/s (serialised, one sync per frame)")
