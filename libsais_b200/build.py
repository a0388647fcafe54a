"""Build libsais_b200/libsais_cuda.so (the C-ABI shared library) with nvcc for sm_100a.

In-tree build: the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
Usage: python -m libsais_b200.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsais_cuda.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["ctx.cu", "hostcopy.cu", "sa_core.cu", "post.cu", "gsa.cu", "dist64.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("LSC_NVCC_EXTRA", "").split()
FLAGS = EXTRA + ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v" if os.environ.get("LSC_PTXAS_V") else "-O3",
         "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", f) for f in ("libsais.h", "libsais64.h", "libsais_cuda.h")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in _deps())


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    srcp = os.path.join(CSRC, src)
    deps = _deps()
    if os.path.exists(obj) and all(os.path.getmtime(p) <= os.path.getmtime(obj) for p in deps):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=4) as ex:
        res = list(ex.map(_compile, SOURCES))
    if verbose:
        for _, log in res:
            if log:
                print(log)
    # API symbols are exported explicitly via the version script; everything else stays hidden
    cmd = [NVCC, "-shared", "-o", OUT] + [o for o, _ in res] + ["-lpthread", "-Xlinker", "--version-script=" + os.path.join(CSRC, "exports.map")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
