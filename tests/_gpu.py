"""Shared helpers for the GPU parity tests: run the CUDA path through the C ABI
and the oracle on the same inputs, compare bit-exactly."""
import numpy as np

import _oracle
import _params


def run_gpu(pkg, ctx, p, frames, dump=True, frame_id_base=0):
    bufs = []
    for i, (y, u, v) in enumerate(frames):
        b = pkg.Yv12Buffer(p["width"], p["height"], p["ss_x"], p["ss_y"], p["use_hbd"], p["border"],
                           p["monochrome"], frame_id=(frame_id_base + i + 1) if frame_id_base else 0)
        b.set_planes(y, u, v)
        bufs.append(b)
    out = pkg.Yv12Buffer(p["width"], p["height"], p["ss_x"], p["ss_y"], p["use_hbd"], p["border"], p["monochrome"])
    res = ctx.temporal_filter(p, bufs, out, dump=dump)
    res["out"] = [out.full_blocks(pl).astype(np.uint16).copy() for pl in range(out.num_planes)]
    res["bufs"] = bufs
    return res


def compare(g, o, p, tol_out=0):
    """Returns dict of mismatch counts. MVs/MSEs/pred must be exact; out within tol_out."""
    rep = {}
    nf, fi = p["num_frames"], p["filter_frame_idx"]
    sel = [f for f in range(nf) if f != fi]
    for k in ("mvs", "mses", "pred"):
        rep[k] = int((g[k][:, sel] != o[k][:, sel]).sum())
    for k in ("accum", "count"):
        if k in g and o.get(k) is not None:
            rep[k] = int((g[k] != o[k]).sum())
    rep["out_bad"] = 0
    rep["out_maxdiff"] = 0
    rep["out_total"] = 0
    for a, b in zip(g["out"], o["out"]):
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        rep["out_bad"] += int((d > tol_out).sum())
        rep["out_mismatch"] = rep.get("out_mismatch", 0) + int((d > 0).sum())
        rep["out_maxdiff"] = max(rep["out_maxdiff"], int(d.max()))
        rep["out_total"] += d.size
    rep["diff_equal"] = bool((g["diff"] == o["diff"]).all())
    return rep


def oracle_run(p, frames):
    o = _oracle.OracleFilter(p, frames)
    r = o.run()
    o.close()
    return r
