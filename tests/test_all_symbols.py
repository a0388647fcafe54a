"""Every one of the 54 exported reference symbols (include/libsais.h: 34, include/libsais64.h: 20) is called on
the GPU library through the C-ABI exactly as on the compiled reference (oracle/_ref, or the oracle port when the
reference did not travel), with the same arguments, and every output buffer and return code must be identical:
all `_omp`, `_ctx`, `create_ctx*` / `unbwt_create_ctx*` variants included."""
import ctypes as C

import numpy as np
import pytest

import _libs
from libsais_b200 import gen

pytestmark = pytest.mark.gpu

VP = C.c_void_p


def p(a):
    return None if a is None else a.ctypes.data_as(VP)


def text_cases():
    rng = np.random.default_rng(123)
    return {
        "dna": gen.dna(11, 30_000),
        "bytes": gen.rand_bytes(12, 20_000),
        "runs": np.repeat(rng.integers(97, 101, 400).astype(np.uint8), rng.integers(1, 200, 400)),
    }


def gsa_text():
    rng = np.random.default_rng(5)
    parts = []
    for _ in range(300):
        parts.append((rng.integers(0, 4, int(rng.integers(1, 90))) + 65).astype(np.uint8))
        parts.append(np.zeros(1, dtype=np.uint8))
    return np.concatenate(parts)


class Api:
    """One implementation (a ctypes library + symbol prefix) with lazily created contexts."""

    def __init__(self, lib, prefix=""):
        self.lib, self.prefix = lib, prefix
        self._ctx = {}

    def fn(self, name, restype):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    def ctx(self, maker, arg=None):
        key = (maker, arg)
        if key not in self._ctx:
            f = self.fn(maker, VP)
            self._ctx[key] = f() if arg is None else f(C.c_int32(arg))
            assert self._ctx[key], "%s returned NULL" % maker
        return self._ctx[key]

    def close(self):
        for (maker, _), h in self._ctx.items():
            self.fn("libsais_unbwt_free_ctx" if "unbwt" in maker else "libsais_free_ctx", None)(VP(h))
        self._ctx.clear()


def run_symbol(api, name, T, Tg):
    """Call `name` on implementation `api`; returns (return code, list of output arrays)."""
    w64 = name.startswith("libsais64")
    it, ct = (np.int64, C.c_int64) if w64 else (np.int32, C.c_int32)
    base = name[len("libsais64" if w64 else "libsais"):]           # "", "_omp", "_bwt_aux_ctx", ...
    omp = base.endswith("_omp")
    ctx = base.endswith("_ctx")
    core = base[:-4] if (omp or ctx) else base
    n = len(T)
    f = api.fn(name, ct)
    pre = []
    if ctx:
        pre = [VP(api.ctx("libsais_unbwt_create_ctx_omp" if core.startswith("_unbwt") else "libsais_create_ctx_omp", 2)
                  if core in ("_bwt_aux", "_unbwt_aux") else
                  api.ctx("libsais_unbwt_create_ctx" if core.startswith("_unbwt") else "libsais_create_ctx"))]
    post = [ct(3)] if omp else []
    # inputs the later stages need come from the reference definition computed once with numpy-level helpers
    if core == "":
        SA = np.full(n + 5, -7, dtype=it); freq = np.full(256, -1, dtype=it)
        rc = f(*pre, p(T), p(SA), ct(n), ct(5), p(freq), *post)
        return rc, [SA[:n], freq]
    if core in ("_int", "_long"):
        Ti = (T.astype(it) * 3 + 1)
        Tc = Ti.copy()
        SA = np.full(n, -7, dtype=it)
        rc = f(p(Ti), p(SA), ct(n), ct(int(Ti.max()) + 1), ct(0), *post)
        return rc, [SA, np.array([int((Ti == Tc).all())])]
    if core == "_gsa":
        m = len(Tg)
        SA = np.full(m, -7, dtype=it); freq = np.full(256, -1, dtype=it)
        rc = f(*pre, p(Tg), p(SA), ct(m), ct(0), p(freq), *post)
        return rc, [SA, freq]
    if core == "_plcp_gsa":
        m = len(Tg)
        SA = np.ascontiguousarray(_libs.oracle().gsa(Tg, 64 if w64 else 32)[1], dtype=it)
        P = np.full(m, -7, dtype=it)
        rc = f(p(Tg), p(SA), p(P), ct(m), *post)
        return rc, [P]
    if core == "_bwt":
        U = np.zeros(n, dtype=np.uint8); A = np.zeros(n, dtype=it); freq = np.full(256, -1, dtype=it)
        rc = f(*pre, p(T), p(U), p(A), ct(n), ct(0), p(freq), *post)
        return rc, [U, freq]
    if core == "_bwt_aux":
        r = 64
        U = np.zeros(n, dtype=np.uint8); A = np.zeros(n, dtype=it); I = np.full((n - 1) // r + 1, -1, dtype=it)
        rc = f(*pre, p(T), p(U), p(A), ct(n), ct(0), None, ct(r), p(I), *post)
        return rc, [U, I]
    if core in ("_unbwt", "_unbwt_aux"):
        o = _libs.oracle()
        bits = 64 if w64 else 32
        U = np.zeros(n, dtype=np.uint8); A = np.zeros(n + 1, dtype=it)
        if core == "_unbwt":
            primary, B = o.bwt(T, bits)
            rc = f(*pre, p(B), p(U), p(A), ct(n), None, ct(primary), *post)
        else:
            _, B, I = o.bwt_aux(T, 128, bits)
            I = np.ascontiguousarray(I, dtype=it)
            rc = f(*pre, p(B), p(U), p(A), ct(n), None, ct(128), p(I), *post)
        return rc, [U]
    if core in ("_plcp", "_plcp_int", "_lcp"):
        o = _libs.oracle()
        bits = 64 if w64 else 32
        SA = np.ascontiguousarray(o.sa(T, bits)[1], dtype=it)
        if core == "_plcp":
            P = np.full(n, -7, dtype=it)
            rc = f(p(T), p(SA), p(P), ct(n), *post)
            return rc, [P]
        if core == "_plcp_int":
            Ti = T.astype(np.int32)
            P = np.full(n, -7, dtype=it)
            rc = f(p(Ti), p(SA), p(P), ct(n), *post)
            return rc, [P]
        P = np.ascontiguousarray(o.plcp(T, SA, bits)[1], dtype=it)
        L = np.full(n, -7, dtype=it)
        rc = f(p(P), p(SA), p(L), ct(n), *post)
        return rc, [L]
    raise AssertionError("unhandled symbol " + name)


SYMBOLS_32 = ["libsais", "libsais_ctx", "libsais_omp", "libsais_int", "libsais_int_omp",
              "libsais_gsa", "libsais_gsa_ctx", "libsais_gsa_omp", "libsais_plcp_gsa", "libsais_plcp_gsa_omp",
              "libsais_bwt", "libsais_bwt_aux", "libsais_bwt_ctx", "libsais_bwt_aux_ctx", "libsais_bwt_omp", "libsais_bwt_aux_omp",
              "libsais_unbwt", "libsais_unbwt_ctx", "libsais_unbwt_aux", "libsais_unbwt_aux_ctx", "libsais_unbwt_omp", "libsais_unbwt_aux_omp",
              "libsais_plcp", "libsais_plcp_int", "libsais_lcp", "libsais_plcp_omp", "libsais_plcp_int_omp", "libsais_lcp_omp"]
SYMBOLS_64 = ["libsais64", "libsais64_omp", "libsais64_long", "libsais64_long_omp",
              "libsais64_gsa", "libsais64_gsa_omp", "libsais64_plcp_gsa", "libsais64_plcp_gsa_omp",
              "libsais64_bwt", "libsais64_bwt_aux", "libsais64_bwt_omp", "libsais64_bwt_aux_omp",
              "libsais64_unbwt", "libsais64_unbwt_aux", "libsais64_unbwt_omp", "libsais64_unbwt_aux_omp",
              "libsais64_plcp", "libsais64_lcp", "libsais64_plcp_omp", "libsais64_lcp_omp"]
CTX_SYMBOLS = ["libsais_create_ctx", "libsais_create_ctx_omp", "libsais_free_ctx",
               "libsais_unbwt_create_ctx", "libsais_unbwt_create_ctx_omp", "libsais_unbwt_free_ctx"]


@pytest.fixture(scope="module")
def apis():
    import libsais_b200
    assert libsais_b200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    cu = Api(libsais_b200.load_library())
    r = _libs.ref()
    other = Api(r.lib) if r is not None else Api(_libs.oracle().lib, "oracle_")
    yield cu, other
    cu.close(); other.close()


def test_symbol_count():
    assert len(SYMBOLS_32) + len(CTX_SYMBOLS) == 34 and len(SYMBOLS_64) == 20


@pytest.mark.parametrize("name", SYMBOLS_32 + SYMBOLS_64)
def test_symbol_matches_reference(apis, name):
    cu, other = apis
    if other.prefix and name.endswith(("_omp", "_ctx")):
        pytest.skip("the oracle port has no _omp/_ctx twins; the compiled reference did not travel to this box")
    Tg = gsa_text()
    for label, T in text_cases().items():
        a = run_symbol(cu, name, T, Tg)
        b = run_symbol(other, name, T, Tg)
        assert a[0] == b[0], "%s on %s: return code %d vs %d" % (name, label, a[0], b[0])
        for x, y in zip(a[1], b[1]):
            assert np.array_equal(x, y), "%s on %s: output differs at %d" % (name, label, int(np.argmax(x != y)))


def test_context_lifecycle(apis):
    """create_ctx / create_ctx_omp / free_ctx (+ unbwt twins): NULL on threads < 0, free(NULL) is a no-op, a
    context survives many calls of different kinds and sizes."""
    cu, _ = apis
    lib = cu.lib
    for maker, freer in (("libsais_create_ctx", "libsais_free_ctx"), ("libsais_unbwt_create_ctx", "libsais_unbwt_free_ctx")):
        getattr(lib, maker).restype = VP
        getattr(lib, maker + "_omp").restype = VP
        getattr(lib, freer).restype = None
        getattr(lib, freer)(VP(None))
        assert getattr(lib, maker + "_omp")(C.c_int32(-1)) is None
        for h in (getattr(lib, maker)(), getattr(lib, maker + "_omp")(C.c_int32(0)), getattr(lib, maker + "_omp")(C.c_int32(4))):
            assert h
            for n in (1000, 50_000, 10):
                T = gen.dna(n, n)
                SA = np.empty(n, dtype=np.int32)
                lib.libsais_ctx.restype = C.c_int32
                assert lib.libsais_ctx(VP(h), p(T), p(SA), C.c_int32(n), C.c_int32(0), None) == 0
                assert (SA == _libs.oracle().sa(T)[1]).all()
            getattr(lib, freer)(VP(h))
