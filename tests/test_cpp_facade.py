"""The C++ facade (include/eslam/*.hpp: the reference's class names over the C ABI) builds with g++ on the CPU box and,
on the GPU box, passes the reference's own test cases rewritten against it (tests/cpp/test_facade.cpp)."""
import os
import subprocess

import pytest

from slam_eslam_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_facade")


def compile_facade_test():
    build.build()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"), "-L" + os.path.join(ROOT, "slam-eslam_b200"), "-leslam_gpu",
           "-Wl,-rpath,$ORIGIN/../../slam-eslam_b200", "-o", BIN]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    return BIN


def test_facade_compiles_and_links():
    assert os.path.exists(compile_facade_test())


@pytest.mark.gpu
def test_facade_runs_reference_test_cases():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(BIN):
        compile_facade_test()
    res = subprocess.run([BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0 and "facade tests ok" in res.stdout, res.stdout[-3000:]
