"""CPU: the C-ABI library loads and exports every symbol include/tf_gpu.h declares; error
behaviour without a device (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "tf_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tf_gpu_\w+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    assert declared_functions() == sorted(pkg.EXPORTS)


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.tf_gpu_abi_version() == 2


def test_struct_layouts(pkg):
    # POD layout the C side compiles to (x86-64 SysV): catches drift between header and binding
    assert C.sizeof(pkg.Frame) == 88
    assert C.sizeof(pkg.Params) == 160  # extend_output_borders took one reserved slot
    assert C.sizeof(pkg.Dump) == 40
    assert C.sizeof(pkg.DeviceCfg) == 32


def test_no_device_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.TfGpuError) as e:
        pkg.TemporalFilterGpu()
    assert e.value.code == -4  # TF_GPU_ERR_NO_DEVICE; no CPU fallback


def test_product_does_not_reference_oracle():
    """The product path must never import, link or call anything under oracle/."""
    for rel in ("aom-av1-psy_b200/__init__.py", "aom-av1-psy_b200/sharding.py", "aom-av1-psy_b200/csrc/tf_gpu.cu",
                "aom-av1-psy_b200/csrc/tf_kernels.cuh", "aom-av1-psy_b200/csrc/Makefile", "include/tf_gpu.h"):
        text = open(os.path.join(ROOT, rel)).read()
        assert "tf_oracle" not in text and "libtf_ref" not in text and "_oracle" not in text, rel
