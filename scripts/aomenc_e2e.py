#!/usr/bin/env python3
"""End-to-end aomenc comparison: the stock reference build against the CONFIG_TF_GPU=1 build
(integration/_build/, made by scripts/build_aomenc.sh) on synthetic moving-texture Y4M clips.

For every clip the unmodified aomenc, the timing-only build and the tf_gpu build encode the same
file with the same command line; the report holds bitrate, PSNR (overall / Y / U / V), the md5 of the
IVF and the wall time spent in the temporal filter (TF_SEAM_TIMING line of the patched builds), plus
the parameters every ARF window was filtered with (TF_SEAM_WINDOW lines of the tf_gpu build).
Identical md5 = the filtered frames entering the encoder were identical.

    python scripts/aomenc_e2e.py [--out profiles/aomenc_e2e_r02.json] [--clips cif8,cif10,hd8]
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _clips  # noqa: E402

BUILD = os.path.join(ROOT, "integration", "_build")

CLIPS = {
    # name: (width, height, bit depth, frames, aomenc arguments)     BASELINE.json configs 1-3
    "cif8": (352, 288, 8, 17, ["--good", "--cpu-used=4", "--arnr-maxframes=7"]),
    "cif10": (352, 288, 10, 17, ["--good", "--cpu-used=4", "--bit-depth=10", "--arnr-maxframes=11"]),
    "hd8": (1920, 1080, 8, 12, ["--good", "--cpu-used=4", "--arnr-maxframes=7", "--arnr-strength=4", "--threads=16"]),
    "hd10": (1920, 1080, 10, 12, ["--good", "--cpu-used=4", "--bit-depth=10", "--arnr-maxframes=11", "--threads=16"]),
    # the bench workload (BASELINE.json config 4): 24 frames so that the ARF at frame 16 sees a full 15-frame window
    "uhd10": (3840, 2160, 10, 24, ["--good", "--cpu-used=4", "--bit-depth=10", "--arnr-maxframes=15", "--end-usage=q",
                                   "--cq-level=32", "--threads=16"]),
}


def write_y4m(path, frames, width, height, bd):
    cs = "C420jpeg" if bd == 8 else f"C420p{bd} XYSCSS=420P{bd}"
    with open(path, "wb") as f:
        f.write(f"YUV4MPEG2 W{width} H{height} F30:1 Ip A1:1 {cs}\n".encode())
        for (y, u, v) in frames:
            f.write(b"FRAME\n")
            for pl in (y, u, v):
                f.write(np.ascontiguousarray(pl.astype("<u2" if bd > 8 else np.uint8)).tobytes())


def encode(binary, clip_path, out_path, args, nframes):
    cmd = [os.path.join(BUILD, binary)] + args + [f"--limit={nframes}", "--psnr", "-o", out_path, clip_path]
    env = dict(os.environ, TF_SEAM_TIMING="1")
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    text = r.stdout + r.stderr
    if r.returncode != 0:
        return {"error": text[-2000:], "cmd": " ".join(cmd)}
    res = {"wall_s": round(wall, 2), "cmd": " ".join(os.path.basename(c) if c.startswith("/") else c for c in cmd)}
    m = re.findall(r"PSNR \(Overall/Avg/Y/U/V\)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", text)
    if m:
        o, a, y, u, v = (float(x) for x in m[-1])
        res["psnr"] = {"overall": o, "avg": a, "y": y, "u": u, "v": v}
    data = open(out_path, "rb").read()
    res["ivf_bytes"] = len(data)
    res["ivf_md5"] = hashlib.md5(data).hexdigest()
    res["bitrate_kbps"] = round(len(data) * 8 / (nframes / 30.0) / 1e3, 3)
    t = re.search(r"TF_SEAM_TIMING impl=(\w+) windows=(\d+) filter_ms=([\d.]+)(?: pushes=(\d+) push_ms=(-?[\d.]+) create_ms=([\d.]+))?", text)
    if t:
        res["tf"] = {"impl": t.group(1), "windows": int(t.group(2)), "filter_wall_ms": float(t.group(3))}
        if t.group(4):
            res["tf"].update(pushes=int(t.group(4)), push_wall_ms=float(t.group(5)), cuda_create_ms=float(t.group(6)))
    wins = re.findall(r"TF_SEAM_WINDOW (.*)", text)
    if wins:
        res["windows"] = [dict(kv.split("=") for kv in w.split()) for w in wins]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "aomenc_e2e.json"))
    ap.add_argument("--clips", default="cif8,cif10,hd8")
    ap.add_argument("--tmp", default="/tmp/aomenc_e2e")
    args = ap.parse_args()
    os.makedirs(args.tmp, exist_ok=True)
    report = {"builds": {"stock": "aomenc_stock (unmodified reference, generic C target)",
                         "stock_timed": "aomenc_stock_timed (+ integration/tf_timing_only.patch)",
                         "tfgpu": "aomenc_tfgpu (+ integration/tf_gpu_seam.patch, CONFIG_TF_GPU=1, libtf_gpu.so)"},
              "clips": {}}
    for name in args.clips.split(","):
        w, h, bd, n, eargs = CLIPS[name]
        frames = _clips.moving_texture(w, h, n, bd)
        clip = os.path.join(args.tmp, f"{name}.y4m")
        write_y4m(clip, frames, w, h, bd)
        entry = {"width": w, "height": h, "bit_depth": bd, "frames": n, "clip": "tests/_clips.moving_texture (seeded)"}
        # the tf_gpu build first: its first call pays CUDA context creation (reported inside wall_s)
        for key, binary in (("tfgpu", "aomenc_tfgpu"), ("stock_timed", "aomenc_stock_timed"), ("stock", "aomenc_stock")):
            entry[key] = encode(binary, clip, os.path.join(args.tmp, f"{name}_{key}.ivf"), eargs, n)
            print(name, key, json.dumps({k: v for k, v in entry[key].items() if k != "windows"})[:400], file=sys.stderr)
        ok = all("ivf_md5" in entry[k] for k in ("tfgpu", "stock", "stock_timed"))
        entry["identical_bitstream"] = ok and entry["tfgpu"]["ivf_md5"] == entry["stock"]["ivf_md5"] == entry["stock_timed"]["ivf_md5"]
        if ok and "psnr" in entry["tfgpu"] and "psnr" in entry["stock"]:
            entry["psnr_delta_db"] = {k: round(entry["tfgpu"]["psnr"][k] - entry["stock"]["psnr"][k], 4) for k in ("overall", "y", "u", "v")}
            entry["bitrate_delta_pct"] = round(100.0 * (entry["tfgpu"]["ivf_bytes"] - entry["stock"]["ivf_bytes"]) / entry["stock"]["ivf_bytes"], 5)
        if ok and "tf" in entry["tfgpu"] and "tf" in entry["stock_timed"]:
            entry["tf_wall_speedup"] = round(entry["stock_timed"]["tf"]["filter_wall_ms"] / max(entry["tfgpu"]["tf"]["filter_wall_ms"], 1e-9), 2)
        report["clips"][name] = entry
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(report, open(args.out, "w"), indent=1)
    print(json.dumps({k: {"identical": v["identical_bitstream"], "tf_speedup": v.get("tf_wall_speedup")} for k, v in report["clips"].items()}))


if __name__ == "__main__":
    main()
