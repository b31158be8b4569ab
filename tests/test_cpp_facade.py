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


def test_libeslam_and_pkgconfig_like_the_reference_installs_them(tmp_path):
    """The reference installs `eslam` as a library with a pkg-config file (src/CMakeLists.txt:16-19, src/eslam.pc.in:6-10): a
    consumer that asks pkg-config for `eslam` compiles and links against the installed headers, libeslam.so and libeslam_gpu.so."""
    import shutil
    build.build()
    prefix = tmp_path / "prefix"
    res = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "packaging"), "install", "PREFIX=%s" % prefix],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    for f in ("lib/libeslam.so", "lib/libeslam_gpu.so", "lib/pkgconfig/eslam.pc", "include/eslam/EmbodiedSlamFilter.hpp", "include/eslam_gpu.h"):
        assert (prefix / f).exists(), f
    if shutil.which("pkg-config"):
        env = dict(os.environ, PKG_CONFIG_PATH=str(prefix / "lib" / "pkgconfig"))
        flags = subprocess.run(["pkg-config", "--cflags", "--libs", "eslam"], env=env, stdout=subprocess.PIPE, text=True).stdout.split()
    else:
        flags = ["-I%s" % (prefix / "include"), "-L%s" % (prefix / "lib"), "-leslam", "-leslam_gpu"]
    assert "-leslam" in flags
    out = tmp_path / "consumer"
    res = subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"), "-o", str(out)] + flags +
                         ["-Wl,-rpath,%s" % (prefix / "lib")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    import ctypes
    lib = ctypes.CDLL(str(prefix / "lib" / "libeslam.so"))
    lib.eslam_version.restype = ctypes.c_char_p
    assert lib.eslam_version().startswith(b"2.")


@pytest.mark.gpu
def test_facade_runs_reference_test_cases():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(BIN):
        compile_facade_test()
    res = subprocess.run([BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0 and "facade tests ok" in res.stdout, res.stdout[-3000:]
