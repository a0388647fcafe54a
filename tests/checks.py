"""Implementation-agnostic checks: run one implementation of the libsais C API (an
_libs.Impl) against the committed golden vectors and against another implementation."""
import hashlib
import json
import os

import numpy as np

import cases

HERE = os.path.dirname(os.path.abspath(__file__))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        return json.load(f)


def kat():
    with open(os.path.join(HERE, "golden", "kat.json")) as f:
        return json.load(f)


def _eq(got, want, full):
    if full:
        return [int(x) for x in got] == want
    return digest(got) == want


def check_case(impl, T, g, full, bits=32):
    """All outputs of one text against its golden record; returns list of failed fields."""
    bad = []
    dt = np.int32 if bits == 32 else np.int64
    assert digest(T) == g["input_sha"], "test input drifted from the golden generator"
    rc, SA, freq = impl.sa(T, bits, want_freq=True)
    if rc != 0 or not _eq(SA.astype(np.int32), g["sa"], full):
        bad.append("sa")
    if not _eq(freq.astype(np.int32), g["freq"], full):
        bad.append("freq")
    rc, U = impl.bwt(T, bits)
    if rc != g["primary"] or not _eq(U, g["bwt"], full):
        bad.append("bwt")
    if len(T) == 0:
        return bad
    SAg = SA if "sa" not in bad else None
    if SAg is not None:
        rc, P = impl.plcp(T, SAg.astype(dt), bits)
        if rc != 0 or not _eq(P.astype(np.int32), g["plcp"], full):
            bad.append("plcp")
        else:
            rc, L = impl.lcp(P.astype(dt), SAg.astype(dt), bits)
            if rc != 0 or not _eq(L.astype(np.int32), g["lcp"], full):
                bad.append("lcp")
    for r in (2, 8, 128):
        rc, U2, I = impl.bwt_aux(T, r, bits)
        if rc != 0 or not _eq(U2, g["bwt"], full) or not _eq(I.astype(np.int32), g["aux_r%d" % r], full):
            bad.append("aux_r%d" % r)
        elif r == 8:
            rc, back = impl.unbwt_aux(U2, r, I.astype(dt), bits)
            if rc != 0 or not (back == T).all():
                bad.append("unbwt_aux")
    if "bwt" not in bad:
        rc, back = impl.unbwt(U, g["primary"], bits)
        if rc != 0 or not (back == T).all():
            bad.append("unbwt")
    return bad


def check_golden(impl, which=("small", "medium", "int"), bits=32):
    g = golden()
    failures = {}
    if "small" in which:
        for name, T in cases.small_cases().items():
            bad = check_case(impl, T, g["small"][name], True, bits)
            if bad:
                failures[name] = bad
    if "medium" in which:
        for name, T in cases.medium_cases().items():
            bad = check_case(impl, T, g["medium"][name], False, bits)
            if bad:
                failures[name] = bad
    if "int" in which:
        for name, (T, k) in cases.int_cases().items():
            gi = g["int"][name]
            T32 = np.ascontiguousarray(T, dtype=np.int32 if bits == 32 else np.int64)
            before = T32.copy()
            rc, SA, Tafter = impl.sa_int(T32, k, bits)
            bad = []
            if rc != 0 or digest(SA.astype(np.int32)) != gi["sa"]:
                bad.append("sa_int")
            if not (Tafter == before).all():
                bad.append("T_modified")
            if bits == 32 and not bad:
                rc, P = impl.plcp(np.ascontiguousarray(T, dtype=np.int32), SA)
                if rc != 0 or digest(P) != gi["plcp"]:
                    bad.append("plcp_int")
            if bad:
                failures["int:" + name] = bad
    return failures


def check_gsa(impl, bits=32):
    """Generalized SA + PLCP-GSA against the golden vectors (generated from the compiled reference)."""
    g = golden()["gsa"]
    dt = np.int32 if bits == 32 else np.int64
    failures = {}
    for name, T in cases.gsa_cases().items():
        gi = g[name]
        assert digest(T) == gi["input_sha"]
        bad = []
        rc, SA, freq = impl.gsa(T, bits, want_freq=True)
        enc = (lambda a: [int(x) for x in a]) if gi["full"] else (lambda a: digest(a.astype(np.int32)))
        if rc != 0 or enc(SA) != gi["sa"]:
            bad.append("gsa")
        elif digest(freq.astype(np.int32)) != gi["freq"]:
            bad.append("freq")
        else:
            rc, P = impl.plcp_gsa(T, SA.astype(dt), bits)
            if rc != 0 or enc(P) != gi["plcp"]:
                bad.append("plcp_gsa")
        if bad:
            failures[name] = bad
    # a collection must end with a separator (reference :7035) and have no empty member (:6886-6889)
    for bad_text in (b"ab\0ab", b"\0a\0", b"a\0\0", b"ab\0\0cd\0", b"\0\0"):
        rc, _ = impl.gsa(np.frombuffer(bad_text, dtype=np.uint8).copy(), bits)
        if rc != -1:
            failures["rejects:" + repr(bad_text)] = [rc]
    return failures


def check_kat(impl):
    k = kat()
    T = np.frombuffer(k["banana"]["text"].encode(), dtype=np.uint8).copy()
    rc, SA = impl.sa(T)
    assert rc == 0 and list(SA) == k["banana"]["sa"]
    rc, U = impl.bwt(T)
    assert rc == k["banana"]["primary"] and U.tobytes().decode() == k["banana"]["bwt"]
    rc, P = impl.plcp(T, SA)
    assert rc == 0 and list(P) == k["banana"]["plcp"]
    rc, L = impl.lcp(P, SA)
    assert rc == 0 and list(L) == k["banana"]["lcp"]
    rc, U, I = impl.bwt_aux(T, 4)
    assert rc == 0 and list(I) == k["banana"]["aux_r4"]
    T = np.frombuffer(k["a8"]["text"].encode(), dtype=np.uint8).copy()
    rc, SA = impl.sa(T)
    assert rc == 0 and list(SA) == k["a8"]["sa"]
    rc, U = impl.bwt(T)
    assert rc == k["a8"]["primary"]
    T = np.frombuffer(b"ab\0ab\0b\0", dtype=np.uint8).copy()
    rc, SA = impl.gsa(T)
    assert rc == 0 and list(SA) == k["gsa"]["sa"]
    rc, P = impl.plcp_gsa(T, SA)
    assert rc == 0 and list(P) == k["gsa"]["plcp"]
    Ti = np.array(k["int"]["text"], dtype=np.int32)
    rc, SA, Tafter = impl.sa_int(Ti, k["int"]["k"])
    assert rc == 0 and list(SA) == k["int"]["sa"] and list(Tafter) == k["int"]["text"]
    rc, P = impl.plcp(Ti, SA)
    assert rc == 0 and list(P) == k["int"]["plcp"]


def check_errors(impl):
    """Argument validation and fast paths (reference src/libsais.c:7020-7027, :7103-7108, :7125,
    :8042-8053).  None of these needs a GPU."""
    import ctypes as C
    from _libs import ptr
    T = np.frombuffer(b"banana", dtype=np.uint8).copy()
    SA = np.zeros(8, dtype=np.int32)
    U = np.zeros(8, dtype=np.uint8)
    f = impl._f
    i32 = C.c_int32
    assert f("libsais", 32)(None, ptr(SA), i32(6), i32(0), None) == -1
    assert f("libsais", 32)(ptr(T), None, i32(6), i32(0), None) == -1
    assert f("libsais", 32)(ptr(T), ptr(SA), i32(-1), i32(0), None) == -1
    assert f("libsais", 32)(ptr(T), ptr(SA), i32(6), i32(-1), None) == -1
    assert f("libsais_bwt", 32)(ptr(T), None, ptr(SA), i32(6), i32(0), None) == -1
    assert f("libsais_bwt", 32)(ptr(T), ptr(U), None, i32(6), i32(0), None) == -1
    I = np.zeros(4, dtype=np.int32)
    for r in (0, 1, 3, 6, -2):
        assert f("libsais_bwt_aux", 32)(ptr(T), ptr(U), ptr(SA), i32(6), i32(0), None, i32(r), ptr(I)) == -1
    assert f("libsais_bwt_aux", 32)(ptr(T), ptr(U), ptr(SA), i32(6), i32(0), None, i32(4), None) == -1
    for i in (0, 7, -1):
        assert f("libsais_unbwt", 32)(ptr(T), ptr(U), ptr(SA), i32(6), None, i32(i)) == -1
    bad_I = np.array([4, 9], dtype=np.int32)
    assert f("libsais_unbwt_aux", 32)(ptr(T), ptr(U), ptr(SA), i32(6), None, i32(4), ptr(bad_I)) == -1
    assert f("libsais_unbwt_aux", 32)(ptr(T), ptr(U), ptr(SA), i32(6), None, i32(3), ptr(I)) == -1
    assert f("libsais_plcp", 32)(ptr(T), None, ptr(SA), i32(6)) == -1
    assert f("libsais_lcp", 32)(ptr(SA), ptr(SA), None, i32(6)) == -1
    # n = 0: returns 0, SA untouched, freq zeroed
    SA[:] = -7
    freq = np.full(256, -1, dtype=np.int32)
    assert f("libsais", 32)(ptr(T), ptr(SA), i32(0), i32(0), ptr(freq)) == 0
    assert (SA == -7).all() and (freq == 0).all()
    assert f("libsais_bwt", 32)(ptr(T), ptr(U), ptr(SA), i32(0), i32(0), None) == 0
    # n = 1
    assert f("libsais", 32)(ptr(T), ptr(SA), i32(1), i32(0), ptr(freq)) == 0
    assert SA[0] == 0 and freq[ord("b")] == 1 and freq.sum() == 1
    assert f("libsais_bwt", 32)(ptr(T), ptr(U), ptr(SA), i32(1), i32(0), None) == 1 and U[0] == ord("b")
    assert f("libsais_bwt_aux", 32)(ptr(T), ptr(U), ptr(SA), i32(1), i32(0), None, i32(2), ptr(I)) == 0 and I[0] == 1
    assert f("libsais_unbwt", 32)(ptr(T), ptr(U), ptr(SA), i32(1), None, i32(1)) == 0 and U[0] == ord("b")
    assert f("libsais_unbwt", 32)(ptr(T), ptr(U), ptr(SA), i32(1), None, i32(0)) == -1
    P = np.full(2, -7, dtype=np.int32)
    SA[0] = 0
    assert f("libsais_plcp", 32)(ptr(T), ptr(SA), ptr(P), i32(1)) == 0 and P[0] == 0
    # 64-bit flavour of the same rules
    i64 = C.c_int64
    SA64 = np.zeros(8, dtype=np.int64)
    assert f("libsais64", 64)(None, ptr(SA64), i64(6), i64(0), None) == -1
    assert f("libsais64", 64)(ptr(T), ptr(SA64), i64(-1), i64(0), None) == -1
    assert f("libsais64_bwt", 64)(ptr(T), ptr(U), ptr(SA64), i64(1), i64(0), None) == 1
    assert f("libsais64_unbwt", 64)(ptr(T), ptr(U), ptr(SA64), i64(6), None, i64(7)) == -1
