"""16-bit symbol entry points (include/libsais16.h: 32 symbols, include/libsais16x64.h: 20): every function is
called on the GPU library exactly as on the compiled reference (oracle/_ref, built from src/libsais16.c + src/libsais16x64.c)
and must return identical outputs.  CPU part: the symbols are exported and the n <= 1 fast paths / argument checks match."""
import ctypes as C

import numpy as np
import pytest

import _libs

VP = C.c_void_p


def p(a):
    return None if a is None else a.ctypes.data_as(VP)


def texts16():
    rng = np.random.default_rng(16)
    return {
        "wide": rng.integers(0, 65536, 40_000).astype(np.uint16),
        "mid": rng.integers(0, 300, 60_000).astype(np.uint16),
        "binary_high": (rng.integers(0, 2, 30_000) * 40_000 + 7).astype(np.uint16),
        "repeats": np.tile(rng.integers(0, 1000, 500).astype(np.uint16), 40),
        "runs": np.repeat(rng.integers(0, 5, 300).astype(np.uint16) + 60_000, rng.integers(1, 150, 300)),
    }


def gsa16():
    rng = np.random.default_rng(17)
    parts = []
    for _ in range(400):
        parts.append(rng.integers(1, 2000, int(rng.integers(1, 80))).astype(np.uint16))
        parts.append(np.zeros(1, dtype=np.uint16))
    return np.concatenate(parts)


def call16(lib, pre, name, T, w64, threads=None, ctx=None):
    """Run libsais16<name> (or libsais16x64<name>); returns (rc, outputs)."""
    it, ct = (np.int64, C.c_int64) if w64 else (np.int32, C.c_int32)
    f = getattr(lib, pre + name)
    f.restype = ct
    n = len(T)
    pre_args = [VP(ctx)] if ctx is not None else []
    post = [ct(threads)] if threads is not None else []
    base = name.replace("_omp", "").replace("_ctx", "")
    o = _libs.oracle()
    bits = 64 if w64 else 32
    Tw = T.astype(np.int64)
    if base in ("", "_gsa"):
        SA = np.full(n + 3, -7, dtype=it); freq = np.full(65536, -1, dtype=it)
        rc = f(*pre_args, p(T), p(SA), ct(n), ct(3), p(freq), *post)
        return rc, [SA[:n], freq]
    if base == "_bwt":
        U = np.zeros(n, dtype=np.uint16); A = np.zeros(n + 1, dtype=it); freq = np.full(65536, -1, dtype=it)
        rc = f(*pre_args, p(T), p(U), p(A), ct(n), ct(0), p(freq), *post)
        return rc, [U, freq]
    if base == "_bwt_aux":
        U = np.zeros(n, dtype=np.uint16); A = np.zeros(n + 1, dtype=it); I = np.full((n - 1) // 128 + 1, -1, dtype=it)
        rc = f(*pre_args, p(T), p(U), p(A), ct(n), ct(0), None, ct(128), p(I), *post)
        return rc, [U, I]
    SA = np.ascontiguousarray(o.sa_int(Tw, 65536, 64)[1], dtype=it)           # SA of the widened text = SA of the 16-bit text
    if base in ("_unbwt", "_unbwt_aux"):
        ISA = np.empty(n, dtype=np.int64); ISA[SA] = np.arange(n)
        p0 = int(ISA[0])
        rows = np.where(SA > 0, T[np.maximum(SA - 1, 0)], 0).astype(np.uint16)
        B = np.concatenate([T[n - 1:], rows[:p0], rows[p0 + 1:]]).astype(np.uint16)
        U = np.zeros(n, dtype=np.uint16); A = np.zeros(n + 1, dtype=it)
        if base == "_unbwt":
            rc = f(*pre_args, p(B), p(U), p(A), ct(n), None, ct(p0 + 1), *post)
        else:
            r = 64
            I = np.ascontiguousarray(ISA[::r] + 1, dtype=it)
            rc = f(*pre_args, p(B), p(U), p(A), ct(n), None, ct(r), p(I), *post)
        return rc, [U]
    if base == "_plcp":
        P = np.full(n, -7, dtype=it)
        rc = f(p(T), p(SA), p(P), ct(n), *post)
        return rc, [P]
    if base == "_lcp":
        P = np.ascontiguousarray(o.plcp(T.astype(np.int32), SA.astype(np.int32))[1], dtype=it)
        L = np.full(n, -7, dtype=it)
        rc = f(p(P), p(SA), p(L), ct(n), *post)
        return rc, [L]
    raise AssertionError(name)


NAMES = ["", "_omp", "_bwt", "_bwt_omp", "_bwt_aux", "_bwt_aux_omp", "_unbwt", "_unbwt_omp", "_unbwt_aux", "_unbwt_aux_omp",
         "_plcp", "_plcp_omp", "_lcp", "_lcp_omp"]
CTX_NAMES = ["_ctx", "_bwt_ctx", "_bwt_aux_ctx", "_unbwt_ctx", "_unbwt_aux_ctx"]


def test_16bit_symbols_are_exported_and_validate():
    import libsais_b200
    lib = libsais_b200.load_library()
    for pre, names in (("libsais16", NAMES + CTX_NAMES + ["_gsa", "_gsa_omp", "_gsa_ctx", "_plcp_gsa", "_plcp_gsa_omp", "_int", "_int_omp",
                                                          "_create_ctx", "_create_ctx_omp", "_free_ctx", "_unbwt_create_ctx", "_unbwt_create_ctx_omp", "_unbwt_free_ctx"]),
                       ("libsais16x64", NAMES + ["_gsa", "_gsa_omp", "_plcp_gsa", "_plcp_gsa_omp", "_long", "_long_omp"])):
        for nm in names:
            assert hasattr(lib, pre + nm), pre + nm
    lib.libsais16.restype = C.c_int32
    T = np.array([5], dtype=np.uint16); SA = np.full(1, -1, dtype=np.int32); freq = np.full(65536, -1, dtype=np.int32)
    assert lib.libsais16(p(T), p(SA), C.c_int32(1), C.c_int32(0), p(freq)) == 0 and SA[0] == 0 and freq[5] == 1 and int(freq.sum()) == 1
    assert lib.libsais16(None, p(SA), C.c_int32(1), C.c_int32(0), None) == -1
    assert lib.libsais16(p(T), p(SA), C.c_int32(-1), C.c_int32(0), None) == -1
    lib.libsais16_gsa.restype = C.c_int32
    assert lib.libsais16_gsa(p(np.array([3, 4], dtype=np.uint16)), p(np.zeros(2, dtype=np.int32)), C.c_int32(2), C.c_int32(0), None) == -1
    lib.libsais16_bwt.restype = C.c_int32
    U = np.zeros(1, dtype=np.uint16)
    assert lib.libsais16_bwt(p(T), p(U), p(SA), C.c_int32(1), C.c_int32(0), None) == 1 and U[0] == 5
    lib.libsais16_bwt_aux.restype = C.c_int32
    assert lib.libsais16_bwt_aux(p(T), p(U), p(SA), C.c_int32(1), C.c_int32(0), None, C.c_int32(3), p(SA)) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("w64", [False, True])
def test_16bit_functions_match_the_reference(w64):
    import libsais_b200
    assert libsais_b200.device_count() > 0
    r = _libs.ref()
    if r is None or not hasattr(r.lib, "libsais16"):
        pytest.skip("the compiled reference (with libsais16) did not travel to this box")
    cu = libsais_b200.load_library()
    pre = "libsais16x64" if w64 else "libsais16"
    names = NAMES + ([] if w64 else CTX_NAMES)
    for label, T in texts16().items():
        for nm in names:
            cctx = rctx = None
            if nm.endswith("_ctx"):
                mk = "libsais16_unbwt_create_ctx" if "unbwt" in nm else "libsais16_create_ctx"
                for L in (cu, r.lib):
                    getattr(L, mk).restype = VP
                cctx, rctx = getattr(cu, mk)(), getattr(r.lib, mk)()
            th = 2 if nm.endswith("_omp") else None
            a = call16(cu, pre, nm, T, w64, th, cctx)
            b = call16(r.lib, pre, nm, T, w64, th, rctx)
            assert a[0] == b[0], (pre + nm, label, a[0], b[0])
            for x, y in zip(a[1], b[1]):
                assert np.array_equal(x, y), (pre + nm, label, int(np.argmax(x != y)))
            if cctx:
                fr = "libsais16_unbwt_free_ctx" if "unbwt" in nm else "libsais16_free_ctx"
                getattr(cu, fr).restype = None; getattr(r.lib, fr).restype = None
                getattr(cu, fr)(VP(cctx)); getattr(r.lib, fr)(VP(rctx))
    # generalized suffix arrays + their PLCP
    Tg = gsa16()
    it, ct = (np.int64, C.c_int64) if w64 else (np.int32, C.c_int32)
    outs = []
    for L in (cu, r.lib):
        f = getattr(L, pre + "_gsa"); f.restype = ct
        SA = np.full(len(Tg), -7, dtype=it); freq = np.full(65536, -1, dtype=it)
        rc = f(p(Tg), p(SA), ct(len(Tg)), ct(0), p(freq))
        g = getattr(L, pre + "_plcp_gsa"); g.restype = ct
        P = np.full(len(Tg), -7, dtype=it)
        rc2 = g(p(Tg), p(SA), p(P), ct(len(Tg)))
        outs.append((rc, rc2, SA, freq, P))
    assert outs[0][0] == outs[1][0] == 0 and outs[0][1] == outs[1][1] == 0
    for x, y in zip(outs[0][2:], outs[1][2:]):
        assert np.array_equal(x, y)
    # integer-alphabet twins
    Ti = np.random.default_rng(3).integers(0, 70_000, 20_000)
    nm = "libsais16x64_long" if w64 else "libsais16_int"
    res = []
    for L in (cu, r.lib):
        f = getattr(L, nm); f.restype = ct
        Tc = np.ascontiguousarray(Ti, dtype=it); SA = np.full(len(Ti), -7, dtype=it)
        rc = f(p(Tc), p(SA), ct(len(Ti)), ct(70_000), ct(0))
        res.append((rc, SA, Tc))
    assert res[0][0] == res[1][0] == 0 and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], Ti)
