"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm prints exactly one
JSON line on stdout with the agreed keys (everything else goes to stderr), non-zero ranks stay silent."""
import json
import os
import subprocess
import sys

import pytest

import _ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not _ref.available(), reason="oracle/_ref/libtf_ref.so not built")


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                           "cif8_n7", "--steps", "1", "--warmup", "1", *args],
                          capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "arf_filtered_frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "cif8_n7" and d["dtype"] == "u8" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["simd"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_generic_c_flavour_is_selectable():
    d = json.loads(_run(None, "--ref-simd", "c").stdout.strip())
    assert d["cpu_baseline"]["simd"].startswith("none")


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
