"""libsais_b200 -- Python host-side mirror of the libsais C API over libsais_cuda.so.

The product is the C-ABI shared library built from ``csrc/`` (hand-written CUDA for sm_100a
behind the reference's own C99 interface, ``include/libsais.h`` / ``include/libsais64.h``).
This module only loads it with ctypes and offers numpy-level wrappers with the reference's
function names, argument meaning and return codes, so tests and the benchmark read like a
C caller.  There is no CPU fallback: if the library is missing it is built with nvcc, and if
it cannot be loaded the import of any compute entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsais_cuda.so")
MAX_KERNEL_CLASSES = 32
_lib = None


class Stats(C.Structure):
    _fields_ = [("n_classes", C.c_int32), ("n_rounds", C.c_int32), ("total_launches", C.c_uint64),
                ("launches", C.c_uint64 * MAX_KERNEL_CLASSES), ("ms", C.c_double * MAX_KERNEL_CLASSES),
                ("bytes", C.c_double * MAX_KERNEL_CLASSES), ("device_ms", C.c_double),
                ("workspace_bytes", C.c_uint64)]


class Round(C.Structure):
    _fields_ = [("h", C.c_uint64), ("n_active", C.c_uint64), ("n_groups", C.c_uint64),
                ("passes", C.c_int32), ("key_bits", C.c_int32), ("device_ms", C.c_double), ("bytes", C.c_double)]


def load_library(build_if_missing=True):
    """Load (building it first if needed) libsais_cuda.so and declare the extra entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        # The prebuilt in-tree library is used as is (it travels to the GPU box with the snapshot, where file
        # times say nothing); it is (re)built only when missing, or when LIBSAIS_CUDA_AUTOBUILD=1 asks for the
        # source-time check.  __graft_entry__.build() / `python -m libsais_b200.build` do the real build.
        from . import build as _build
        if not os.path.exists(LIB_PATH) or (os.environ.get("LIBSAIS_CUDA_AUTOBUILD") == "1" and _build.needs_build()):
            _build.build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libsais_cuda.so is missing: run `python -m libsais_b200.build` (needs nvcc)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.libsais_create_ctx.restype = vp
    lib.libsais_cuda_create_ctx.restype = vp
    lib.libsais_cuda_create_ctx.argtypes = [i32]
    lib.libsais_free_ctx.argtypes = [vp]
    lib.libsais_free_ctx.restype = None
    lib.libsais_cuda_stream.restype = vp
    lib.libsais_cuda_stream.argtypes = [vp]
    lib.libsais_cuda_set_profiling.argtypes = [vp, i32]
    lib.libsais_cuda_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.libsais_cuda_get_round.argtypes = [vp, i32, C.POINTER(Round)]
    lib.libsais_cuda_kernel_class_name.restype = C.c_char_p
    lib.libsais_cuda_kernel_class_name.argtypes = [i32]
    lib.libsais_cuda_last_error.argtypes = [vp]
    for name in ("sa_dev", "bwt_dev", "plcp_dev", "lcp_dev", "unbwt_dev"):
        getattr(lib, "libsais_cuda_" + name).restype = i64
    lib.libsais_cuda_sa_dev.argtypes = [vp, vp, vp, i64]
    lib.libsais_cuda_bwt_dev.argtypes = [vp, vp, vp, i64]
    lib.libsais_cuda_plcp_dev.argtypes = [vp, vp, vp, vp, i64]
    lib.libsais_cuda_lcp_dev.argtypes = [vp, vp, vp, vp, i64]
    lib.libsais_cuda_unbwt_dev.argtypes = [vp, vp, vp, i64, i64]
    for name in ("libsais_bwt_ctx",):
        getattr(lib, name).restype = i32
    lib.libsais_bwt_ctx.argtypes = [vp, vp, vp, vp, i32, i32, vp]
    lib.libsais_ctx.argtypes = [vp, vp, vp, i32, i32, vp]
    _lib = lib
    return lib


def device_count():
    return int(load_library().libsais_cuda_device_count())


class Context:
    """A libsais_cuda context bound to one GPU (reference: libsais_create_ctx, include/libsais.h:57)."""

    def __init__(self, device=-1):
        self.lib = load_library()
        self.handle = self.lib.libsais_cuda_create_ctx(int(device))
        if not self.handle:
            raise RuntimeError("libsais_cuda: could not create a context on device %d (no usable GPU?)" % device)

    def close(self):
        if self.handle:
            self.lib.libsais_free_ctx(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return int(self.lib.libsais_cuda_stream(self.handle) or 0)

    def set_profiling(self, on):
        self.lib.libsais_cuda_set_profiling(self.handle, 1 if on else 0)

    def stats(self):
        s = Stats()
        self.lib.libsais_cuda_get_stats(self.handle, C.byref(s))
        names = [self.lib.libsais_cuda_kernel_class_name(i).decode() for i in range(s.n_classes)]
        per = {names[i]: {"launches": int(s.launches[i]), "ms": float(s.ms[i]), "bytes": float(s.bytes[i])}
               for i in range(s.n_classes) if s.launches[i]}
        rounds = []
        for r in range(s.n_rounds):
            rd = Round()
            self.lib.libsais_cuda_get_round(self.handle, r, C.byref(rd))
            rounds.append({"h": int(rd.h), "n_active": int(rd.n_active), "n_groups": int(rd.n_groups),
                           "passes": int(rd.passes), "key_bits": int(rd.key_bits), "ms": float(rd.device_ms), "bytes": float(rd.bytes)})
        return {"total_launches": int(s.total_launches), "device_ms": float(s.device_ms),
                "workspace_bytes": int(s.workspace_bytes), "kernels": per, "rounds": rounds}

    def release_workspace(self):
        """Give the context's device workspace back (it is re-grown by the next call)."""
        self.lib.libsais_cuda_release_workspace.argtypes = [C.c_void_p]
        return int(self.lib.libsais_cuda_release_workspace(self.handle))

    def last_error(self):
        return int(self.lib.libsais_cuda_last_error(self.handle))

    # ---- host-pointer calls (numpy arrays), reference semantics
    def libsais(self, T, SA, fs=0, freq=None):
        """libsais_ctx (include/libsais.h:119)."""
        n = len(T)
        return int(self.lib.libsais_ctx(self.handle, T.ctypes.data, SA.ctypes.data, n, fs,
                                        None if freq is None else freq.ctypes.data))

    def libsais_bwt(self, T, U, A, fs=0, freq=None):
        """libsais_bwt_ctx (include/libsais.h:209): returns the primary index."""
        n = len(T)
        return int(self.lib.libsais_bwt_ctx(self.handle, T.ctypes.data, U.ctypes.data, A.ctypes.data, n, fs,
                                            None if freq is None else freq.ctypes.data))

    def bwt_ptr(self, T_ptr, U_ptr, A_ptr, n):
        """libsais_bwt_ctx on raw host addresses (pinned buffers owned by the caller)."""
        return int(self.lib.libsais_bwt_ctx(self.handle, T_ptr, U_ptr, A_ptr, n, 0, None))

    # ---- device-pointer calls (addresses of device memory on this context's GPU)
    def sa_dev(self, d_T, d_SA, n):
        return int(self.lib.libsais_cuda_sa_dev(self.handle, d_T, d_SA, n))

    def bwt_dev(self, d_T, d_U, n):
        return int(self.lib.libsais_cuda_bwt_dev(self.handle, d_T, d_U, n))

    def plcp_dev(self, d_T, d_SA, d_PLCP, n):
        return int(self.lib.libsais_cuda_plcp_dev(self.handle, d_T, d_SA, d_PLCP, n))

    def lcp_dev(self, d_PLCP, d_SA, d_LCP, n):
        return int(self.lib.libsais_cuda_lcp_dev(self.handle, d_PLCP, d_SA, d_LCP, n))

    def unbwt_dev(self, d_B, d_U, n, primary):
        return int(self.lib.libsais_cuda_unbwt_dev(self.handle, d_B, d_U, n, primary))


def _i32(n):
    return np.empty(max(int(n), 1), dtype=np.int32)


def libsais(T, fs=0, want_freq=False):
    """Suffix array of a uint8 numpy array (reference libsais(), include/libsais.h:84).
    Returns SA (int32[n]); raises on a non-zero return code."""
    lib = load_library()
    T = np.ascontiguousarray(T, dtype=np.uint8)
    SA = _i32(len(T) + fs)
    freq = np.zeros(256, dtype=np.int32) if want_freq else None
    lib.libsais.restype = C.c_int32
    rc = lib.libsais(T.ctypes.data_as(C.c_void_p), SA.ctypes.data_as(C.c_void_p), C.c_int32(len(T)),
                     C.c_int32(fs), None if freq is None else freq.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("libsais returned %d" % rc)
    return (SA[:len(T)], freq) if want_freq else SA[:len(T)]


def libsais_bwt(T):
    """BWT + primary index (reference libsais_bwt(), include/libsais.h:182)."""
    lib = load_library()
    T = np.ascontiguousarray(T, dtype=np.uint8)
    U = np.empty(len(T), dtype=np.uint8)
    A = _i32(len(T))
    lib.libsais_bwt.restype = C.c_int32
    rc = lib.libsais_bwt(T.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p),
                         C.c_int32(len(T)), C.c_int32(0), None)
    if rc < 0:
        raise RuntimeError("libsais_bwt returned %d" % rc)
    return U, int(rc)


def libsais_unbwt(B, primary):
    """Inverse BWT (reference libsais_unbwt(), include/libsais.h:289)."""
    lib = load_library()
    B = np.ascontiguousarray(B, dtype=np.uint8)
    U = np.empty(len(B), dtype=np.uint8)
    A = _i32(len(B) + 1)
    lib.libsais_unbwt.restype = C.c_int32
    rc = lib.libsais_unbwt(B.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p),
                           C.c_int32(len(B)), None, C.c_int32(primary))
    if rc != 0:
        raise RuntimeError("libsais_unbwt returned %d" % rc)
    return U
