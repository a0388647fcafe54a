"""CMake packaging: the repository installs a `libsais` package -- same package name, target name and
headers as the reference's CMakeLists.txt:65-94 -- so an existing `find_package(libsais)` consumer re-links against
the GPU library unchanged.  Packages the prebuilt in-tree .so (seconds), configures and builds tests/cmake_consumer
against the installed prefix, and runs it (on a box without a GPU the library answers -2: no CPU fallback)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cmake():
    exe = shutil.which("cmake") or os.path.join(os.path.dirname(sys.executable), "cmake")
    return exe if exe and os.path.exists(exe) else None


def test_find_package_consumer_links_the_gpu_library(tmp_path):
    cmake = _cmake()
    if cmake is None:
        pytest.skip("cmake not available")
    import libsais_b200
    libsais_b200.load_library()
    so = libsais_b200.LIB_PATH
    prefix, b1, b2 = tmp_path / "prefix", tmp_path / "b1", tmp_path / "b2"
    run = lambda *a: subprocess.run(list(a), check=True, capture_output=True, text=True)
    run(cmake, "-S", ROOT, "-B", str(b1), "-DLIBSAIS_CUDA_PREBUILT=" + so)
    run(cmake, "--install", str(b1), "--prefix", str(prefix))
    for rel in ("include/libsais.h", "include/libsais64.h"):
        assert (prefix / rel).exists()
    cfg = [p for p in prefix.rglob("libsaisConfig.cmake")]
    assert cfg, "libsaisConfig.cmake was not installed"
    run(cmake, "-S", os.path.join(ROOT, "tests", "cmake_consumer"), "-B", str(b2), "-DCMAKE_PREFIX_PATH=" + str(prefix))
    run(cmake, "--build", str(b2))
    exe = b2 / "consumer"
    ldd = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
    assert "libsais_cuda.so" in ldd
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "libsais rc=" in r.stdout
