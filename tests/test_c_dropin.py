"""The drop-in claim, literally: one C99 program written against the libsais API is compiled twice --
against the unmodified reference (oracle/_ref) and against libsais_cuda.so -- and both binaries must
print the same lines.  Fast paths / validation run everywhere; the computing part needs a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "dropin_demo.c")
OUT = os.path.join(ROOT, "tests", "c", "_build")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsais_ref.so")
CUDA_DIR = os.path.join(ROOT, "libsais_b200")


def _build(which):
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "demo_" + which)
    if which == "ref":
        # the reference's prototypes are the same; link by path
        cmd = ["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"), SRC, REF_SO, "-Wl,-rpath," + os.path.dirname(REF_SO), "-fopenmp", "-o", exe]
    else:
        import libsais_b200
        libsais_b200.load_library()
        cmd = ["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-L", CUDA_DIR, "-lsais_cuda", "-Wl,-rpath," + CUDA_DIR, "-o", exe]
    subprocess.check_call(cmd)
    return exe


def _run(exe, *args):
    return subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=300).stdout.splitlines()


def test_fast_paths_and_validation_match_reference_binary():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built")
    a, b = _run(_build("ref"), 1, 1, 1, "fast"), _run(_build("cuda"), 1, 1, 1, "fast")
    assert a and a == b, (a, b)


@pytest.mark.gpu
def test_same_c_program_same_output_on_gpu():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built")
    ref, cu = _build("ref"), _build("cuda")
    for n, sigma, seed in ((6, 3, 1), (1000, 2, 2), (100000, 4, 3), (3000000, 26, 4), (50000, 1, 5)):
        a, b = _run(ref, n, sigma, seed), _run(cu, n, sigma, seed)
        assert a and a == b, (n, sigma, seed, a, b)
