"""One-off: where the host thread spends its time in the e2e loop (4K 10-bit, 2 contexts, depth 2)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest, bench, _params
pkg = conftest.load_package()
W, H, BD, N = 3840, 2160, 10, 15
frames = bench.make_window(W, H, BD, N, seed=77)
p = _params.tf_params(W, H, N, bit_depth=BD, q_factor=bench.Q_FACTOR, filter_strength=5)
conc, depth = 2, 2
ctxs = [pkg.TemporalFilterGpu() for _ in range(conc)]
wins, outs = [], []
for ci in range(conc):
    bufs = []
    for (y, u, v) in frames:
        b = pkg.Yv12Buffer(W, H, 1, 1, True, p["border"])
        b.set_planes(y, u, v, extend=False)
        for a in b.alloc:
            ctxs[ci].host_register(a)
        bufs.append(b)
    wins.append(bufs)
    row = []
    for d in range(depth):
        o = pkg.Yv12Buffer(W, H, 1, 1, True, p["border"])
        for a in o.alloc:
            ctxs[ci].host_register(a)
        row.append(o)
    outs.append(row)
p["noise_levels"] = tuple(ctxs[0].estimate_noise_from_single_plane(wins[0][N // 2], pl, BD) for pl in range(3))
cp = pkg.make_params(p)
def run(nsteps, acc):
    q = [[] for _ in range(conc)]
    for k in range(nsteps):
        for ci in range(conc):
            if len(q[ci]) == depth:
                t = time.perf_counter(); ctxs[ci].wait(q[ci].pop(0)[0]); acc[1] += time.perf_counter() - t
            t = time.perf_counter(); q[ci].append(ctxs[ci].submit(cp, wins[ci], outs[ci][k % depth])); acc[0] += time.perf_counter() - t
    for ci in range(conc):
        for t_ in q[ci]:
            t = time.perf_counter(); ctxs[ci].wait(t_[0]); acc[1] += time.perf_counter() - t
run(3, [0, 0])
acc = [0.0, 0.0]
t0 = time.perf_counter(); K = 8
run(K, acc)
dt = time.perf_counter() - t0
print(f"{K * conc} windows in {dt * 1e3:.1f} ms = {K * conc / dt:.2f} fps; host time in submit {acc[0] * 1e3:.1f} ms ({acc[0] / (K * conc) * 1e3:.2f} ms each), in wait {acc[1] * 1e3:.1f} ms")
