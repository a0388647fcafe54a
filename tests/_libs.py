"""ctypes loaders for the three implementations the tests compare:
   * cuda   -- the product: libsais_b200/libsais_cuda.so (through libsais_b200's own loader)
   * oracle -- oracle/liboracle.so, the CPU restatement (test infrastructure)
   * ref    -- oracle/_ref/libsais_ref.so, the unmodified reference compiled by oracle/Makefile
Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may touch oracle/."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsais_ref.so")


def _ensure_oracle():
    src = os.path.join(ROOT, "oracle", "oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Impl:
    """Uniform numpy-level view of one implementation of the libsais C API.
    `prefix` is prepended to the reference symbol names (``oracle_`` for the oracle)."""

    def __init__(self, lib, prefix="", omp=False):
        self.lib, self.prefix, self.omp = lib, prefix, omp

    def _f(self, name, bits):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = C.c_int32 if bits == 32 else C.c_int64
        return fn

    @staticmethod
    def _it(bits):
        return (np.int32, C.c_int32) if bits == 32 else (np.int64, C.c_int64)

    def sa(self, T, bits=32, fs=0, want_freq=False):
        dt, ct = self._it(bits)
        n = len(T)
        SA = np.full(n + fs, -7, dtype=dt)
        freq = np.full(256, -1, dtype=dt) if want_freq else None
        name = "libsais" if bits == 32 else "libsais64"
        rc = self._f(name, bits)(ptr(T), ptr(SA), ct(n), ct(fs), ptr(freq))
        return (rc, SA[:n], freq) if want_freq else (rc, SA[:n])

    def gsa(self, T, bits=32, want_freq=False):
        dt, ct = self._it(bits)
        n = len(T)
        SA = np.full(max(n, 1), -7, dtype=dt)
        freq = np.full(256, -1, dtype=dt) if want_freq else None
        name = "libsais_gsa" if bits == 32 else "libsais64_gsa"
        rc = self._f(name, bits)(ptr(T), ptr(SA), ct(n), ct(0), ptr(freq))
        return (rc, SA[:n], freq) if want_freq else (rc, SA[:n])

    def plcp_gsa(self, T, SA, bits=32):
        dt, ct = self._it(bits)
        n = len(T)
        SA = np.ascontiguousarray(SA, dtype=dt)
        P = np.full(max(n, 1), -7, dtype=dt)
        name = "libsais_plcp_gsa" if bits == 32 else "libsais64_plcp_gsa"
        rc = self._f(name, bits)(ptr(T), ptr(SA), ptr(P), ct(n))
        return rc, P[:n]

    def sa_int(self, T, k, bits=32, fs=0):
        dt, ct = self._it(bits)
        n = len(T)
        T = np.ascontiguousarray(T, dtype=dt)
        SA = np.full(n + fs, -7, dtype=dt)
        name = "libsais_int" if bits == 32 else "libsais64_long"
        rc = self._f(name, bits)(ptr(T), ptr(SA), ct(n), ct(k), ct(fs))
        return rc, SA[:n], T

    def bwt(self, T, bits=32, want_freq=False, inplace=False):
        dt, ct = self._it(bits)
        n = len(T)
        U = T if inplace else np.zeros(n, dtype=np.uint8)
        A = np.zeros(max(n, 1), dtype=dt)
        freq = np.full(256, -1, dtype=dt) if want_freq else None
        name = "libsais_bwt" if bits == 32 else "libsais64_bwt"
        rc = self._f(name, bits)(ptr(T), ptr(U), ptr(A), ct(n), ct(0), ptr(freq))
        return (rc, U, freq) if want_freq else (rc, U)

    def bwt_aux(self, T, r, bits=32):
        dt, ct = self._it(bits)
        n = len(T)
        U = np.zeros(n, dtype=np.uint8)
        A = np.zeros(max(n, 1), dtype=dt)
        I = np.full((max(n, 1) - 1) // max(r, 1) + 1 if r > 0 else 1, -1, dtype=dt)
        name = "libsais_bwt_aux" if bits == 32 else "libsais64_bwt_aux"
        rc = self._f(name, bits)(ptr(T), ptr(U), ptr(A), ct(n), ct(0), None, ct(r), ptr(I))
        return rc, U, I

    def unbwt(self, B, primary, bits=32, inplace=False):
        dt, ct = self._it(bits)
        n = len(B)
        U = B if inplace else np.zeros(n, dtype=np.uint8)
        A = np.zeros(n + 1, dtype=dt)
        name = "libsais_unbwt" if bits == 32 else "libsais64_unbwt"
        rc = self._f(name, bits)(ptr(B), ptr(U), ptr(A), ct(n), None, ct(primary))
        return rc, U

    def unbwt_aux(self, B, r, I, bits=32):
        dt, ct = self._it(bits)
        n = len(B)
        U = np.zeros(n, dtype=np.uint8)
        A = np.zeros(n + 1, dtype=dt)
        I = np.ascontiguousarray(I, dtype=dt)
        name = "libsais_unbwt_aux" if bits == 32 else "libsais64_unbwt_aux"
        rc = self._f(name, bits)(ptr(B), ptr(U), ptr(A), ct(n), None, ct(r), ptr(I))
        return rc, U

    def plcp(self, T, SA, bits=32):
        dt, ct = self._it(bits)
        n = len(T)
        SA = np.ascontiguousarray(SA, dtype=dt)
        P = np.full(max(n, 1), -7, dtype=dt)
        if T.dtype == np.uint8:
            name = "libsais_plcp" if bits == 32 else "libsais64_plcp"
        else:
            name = "libsais_plcp_int"
            T = np.ascontiguousarray(T, dtype=np.int32)
        rc = self._f(name, bits)(ptr(T), ptr(SA), ptr(P), ct(n))
        return rc, P[:n]

    def lcp(self, PLCP, SA, bits=32):
        dt, ct = self._it(bits)
        n = len(SA)
        SA = np.ascontiguousarray(SA, dtype=dt)
        PLCP = np.ascontiguousarray(PLCP, dtype=dt)
        L = np.full(max(n, 1), -7, dtype=dt)
        name = "libsais_lcp" if bits == 32 else "libsais64_lcp"
        rc = self._f(name, bits)(ptr(PLCP), ptr(SA), ptr(L), ct(n))
        return rc, L[:n]


_cache = {}


def oracle():
    if "oracle" not in _cache:
        _ensure_oracle()
        _cache["oracle"] = Impl(C.CDLL(ORACLE_SO), "oracle_")
    return _cache["oracle"]


def ref():
    """The compiled reference, or None when oracle/_ref was not built (no /root/reference)."""
    if "ref" not in _cache:
        _ensure_oracle()
        _cache["ref"] = Impl(C.CDLL(REF_SO), "") if os.path.exists(REF_SO) else None
    return _cache["ref"]


def cuda():
    if "cuda" not in _cache:
        import libsais_b200
        _cache["cuda"] = Impl(libsais_b200.load_library(), "")
    return _cache["cuda"]
