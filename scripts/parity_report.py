"""Parity report (SURVEY 8d "Parity report"): CUDA path through the C ABI vs the oracle over the fixed parity
cases and a seeded random sweep -- counts of compared / mismatching MVs, block errors, predictor samples,
accumulators, filtered pixels (and their maximum difference) and FRAME_DIFF.  Writes profiles/parity_<R>.json.
usage (GPU box): python scripts/parity_report.py [R] [number of random configurations]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest
import _clips, _gpu, _params
from test_gpu_parity import CASES
from test_gpu_fuzz import _case

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
NFUZZ = int(sys.argv[2]) if len(sys.argv) > 2 else 300
pkg = conftest.load_package()
ctx = pkg.TemporalFilterGpu()
tot = dict(configurations=0, block_frames=0, mv_components=0, mv_mismatch=0, block_errors=0, block_error_mismatch=0,
           predictor_samples=0, predictor_mismatch=0, accumulator_entries=0, accumulator_mismatch=0,
           filtered_pixels=0, filtered_pixel_mismatch=0, filtered_pixel_max_abs_diff=0, frame_diff_compared=0,
           frame_diff_mismatch=0)


def account(g, o, p):
    rep = _gpu.compare(g, o, p, tol_out=1)
    nf, fi = p["num_frames"], p["filter_frame_idx"]
    sel = [f for f in range(nf) if f != fi]
    tot["configurations"] += 1
    tot["block_frames"] += g["mvs"].shape[0] * len(sel)
    tot["mv_components"] += int(g["mvs"][:, sel].size)
    tot["mv_mismatch"] += rep["mvs"]
    tot["block_errors"] += int(g["mses"][:, sel].size)
    tot["block_error_mismatch"] += rep["mses"]
    tot["predictor_samples"] += int(g["pred"][:, sel].size)
    tot["predictor_mismatch"] += rep["pred"]
    tot["accumulator_entries"] += int(g["accum"].size + g["count"].size)
    tot["accumulator_mismatch"] += rep.get("accum", 0) + rep.get("count", 0)
    tot["filtered_pixels"] += rep["out_total"]
    tot["filtered_pixel_mismatch"] += rep.get("out_mismatch", 0)
    tot["filtered_pixel_max_abs_diff"] = max(tot["filtered_pixel_max_abs_diff"], rep["out_maxdiff"])
    tot["frame_diff_compared"] += 1
    tot["frame_diff_mismatch"] += 0 if rep["diff_equal"] else 1


for name, W, H, N, bd, ckw, pkw in CASES:
    frames = _clips.moving_texture(W, H, N, bd, ss_x=pkw.get("ss_x", 1), ss_y=pkw.get("ss_y", 1),
                                   monochrome=pkw.get("monochrome", 0), **ckw)
    p = _params.tf_params(W, H, N, bit_depth=bd, **pkw)
    account(_gpu.run_gpu(pkg, ctx, p, frames), _gpu.oracle_run(p, frames), p)
for seed in range(NFUZZ):
    W, H, N, bd, kw, clip, random_frames = _case(seed)
    fk = dict(ss_x=kw["ss_x"], ss_y=kw["ss_y"], monochrome=kw["monochrome"])
    frames = (_clips.random_frames(W, H, N, bd, seed=clip["seed"], **fk) if random_frames
              else _clips.moving_texture(W, H, N, bd, **fk, **clip))
    p = _params.tf_params(W, H, N, bit_depth=bd, **kw)
    account(_gpu.run_gpu(pkg, ctx, p, frames), _gpu.oracle_run(p, frames), p)
tot["filtered_pixel_mismatch_fraction"] = tot["filtered_pixel_mismatch"] / max(tot["filtered_pixels"], 1)
tot["cases"] = f"{len(CASES)} fixed (tests/test_gpu_parity.py CASES) + {NFUZZ} seeded random configurations (tests/test_gpu_fuzz.py)"
tot["oracle"] = "oracle/libtf_oracle.so, itself bit-identical to the compiled reference (tests/test_oracle_vs_ref.py, test_oracle_fuzz.py)"
tot["end_to_end"] = ("aomenc built with CONFIG_TF_GPU=1 against the stock build, same command lines, five clips (CIF 8/10-bit, "
                     "1080p 8/10-bit, 4K 10-bit): IVF files byte-identical, bitrate delta 0, PSNR delta 0.0 dB "
                     "(profiles/aomenc_e2e_r02.json, scripts/aomenc_e2e.py)")
for d in ("profiles", "gpurun_out"):
    os.makedirs(os.path.join(ROOT, d), exist_ok=True)
    json.dump(tot, open(os.path.join(ROOT, d, f"parity_{R}.json"), "w"), indent=1)
print(json.dumps(tot, indent=1))
