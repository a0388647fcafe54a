"""CPU tests of the drop-in boundary: libsais_cuda.so builds, loads, exports every symbol the
headers declare, and mirrors the reference's argument validation and n <= 1 fast paths -- none
of which needs a GPU.  Computing calls must fail loudly (-2) when no CUDA device exists."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

import _libs
import checks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import libsais_b200
    return libsais_b200.device_count() > 0


def test_library_builds_and_exports_every_declared_symbol():
    import libsais_b200
    lib = libsais_b200.load_library()
    with open(os.path.join(ROOT, "tests", "golden", "api_symbols.json")) as f:
        table = json.load(f)
    declared = []
    for hdr in ("libsais.h", "libsais64.h", "libsais_cuda.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # declarations only, no comments
        declared += re.findall(r"\b(libsais\w*)\s*\(", text)
    declared = sorted(set(declared))
    # the headers declare exactly the reference's interface (34 + 20 symbols) plus the extras
    assert set(table["libsais.h"]) <= set(declared) and len(table["libsais.h"]) == 34
    assert set(table["libsais64.h"]) <= set(declared) and len(table["libsais64.h"]) == 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert missing == []


def test_only_api_symbols_are_exported():
    import libsais_b200
    libsais_b200.load_library()
    out = subprocess.run(["nm", "-D", "--defined-only", libsais_b200.LIB_PATH], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if l.strip()]
    assert names and all(n.startswith("libsais") for n in names), [n for n in names if not n.startswith("libsais")][:5]


def test_headers_compile_as_c99_and_cxx():
    src = '#include "libsais.h"\n#include "libsais64.h"\n#include "libsais_cuda.h"\nint main(void){return 0;}\n'
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        r = subprocess.run([cc, std, "-DLIBSAIS_OPENMP", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                            "-x", "c" if cc == "gcc" else "c++", "-"], input=src, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_argument_validation_and_fast_paths():
    checks.check_errors(_libs.cuda())


def test_gsa_entry_points_validate_like_the_reference():
    """GSA needs a trailing separator (reference src/libsais.c:7035); n <= 1 needs no GPU."""
    cu = _libs.cuda()
    rc, _ = cu.gsa(np.frombuffer(b"ab\0ab", dtype=np.uint8).copy())
    assert rc == -1
    rc, SA = cu.gsa(np.zeros(1, dtype=np.uint8))
    assert rc == 0 and SA[0] == 0
    rc, _ = cu.gsa(np.zeros(0, dtype=np.uint8))
    assert rc == 0


def test_no_cpu_fallback_without_gpu():
    """The product never computes on the CPU: with no CUDA device the call fails with -2."""
    if _has_gpu():
        pytest.skip("a GPU is present")
    import libsais_b200
    rc, SA = _libs.cuda().sa(np.frombuffer(b"banana", dtype=np.uint8).copy())
    assert rc == -2
    with pytest.raises(RuntimeError):
        libsais_b200.Context(0)


def test_product_does_not_reference_the_oracle():
    """Nothing under libsais_b200/ may import, link or open oracle/."""
    pkg = os.path.join(ROOT, "libsais_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".map")):
                text = open(os.path.join(dp, f)).read()
                assert "liboracle" not in text and "oracle/" not in text and "_ref" not in text, os.path.join(dp, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libsais_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libsais_ref" not in out


def test_python_mirrors_of_the_stats_structs_match_the_header(tmp_path):
    """libsais_b200.Stats / Round are ctypes mirrors of libsais_cuda_stats / libsais_cuda_round (include/libsais_cuda.h):
    a C99 program prints the compiler's sizes and field offsets, which must equal the mirrors'."""
    import libsais_b200
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "libsais_cuda.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(libsais_cuda_round), offsetof(libsais_cuda_round, key_bits),\n'
                   '    offsetof(libsais_cuda_round, device_ms), offsetof(libsais_cuda_round, bytes), sizeof(libsais_cuda_stats)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    R = libsais_b200.Round
    assert got == [C.sizeof(R), R.key_bits.offset, R.device_ms.offset, R.bytes.offset, C.sizeof(libsais_b200.Stats)]
