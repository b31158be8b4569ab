"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/eslam_gpu.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from slam_eslam_b200 import abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "eslam_gpu.h")).read()
    return sorted(set(re.findall(r"\b(eslam_gpu_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert sorted(abi.EXPORTED_SYMBOLS) == declared_symbols()


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_default_config_matches_reference_defaults(lib_path):
    """src/Configuration.hpp:53-59,85-111"""
    lib = abi.load_library(lib_path)
    c = abi.Config()
    lib.eslam_gpu_default_config(ctypes.byref(c))
    d = abi.default_config()
    for name, _ in abi.Config._fields_:
        assert getattr(c, name) == getattr(d, name), name
    assert (c.seed, c.min_effective, c.min_contacts) == (42, 50, 3)
    assert (c.measurement_error, c.discount_factor, c.slip_factor) == (0.1, 0.9, 0.05)
    assert (c.contact_likelihood_correction, c.contact_point_radius) == (0.33, 0.01)


def test_struct_sizes_match_the_header():
    assert ctypes.sizeof(abi.Contact) == 40
    assert ctypes.sizeof(abi.Step) == 72
    assert ctypes.sizeof(abi.UpdateResult) == 56
    assert ctypes.sizeof(abi.Config) == 144


def test_no_cpu_fallback(lib_path):
    """Without a CUDA device the product refuses to create a context instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = abi.load_library(lib_path)
    ctx = ctypes.c_void_p()
    cfg = abi.default_config()
    assert lib.eslam_gpu_create(ctypes.byref(cfg), 0, None, ctypes.byref(ctx)) == abi.ERR_CUDA
    with pytest.raises(RuntimeError):
        from slam_eslam_b200 import GpuPoseEstimator
        GpuPoseEstimator()


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing shipped may import, include or link it."""
    bad = []
    for base in ("slam-eslam_b200", "slam_eslam_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".inc")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"#include\s*[\"<][^\">]*oracle|import\s+oracle|from\s+oracle|libeslam_oracle|libeslam_ref|ref_shim", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
