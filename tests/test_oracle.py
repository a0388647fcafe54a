"""CPU tests: the oracle (oracle/oracle.c) is pinned against the committed golden vectors
(generated from the unmodified reference), the SURVEY §8c known answers, the brute-force
definition, and -- where oracle/_ref is present -- the compiled reference itself."""
import numpy as np
import pytest

import _libs
import cases
import checks


def test_oracle_known_answers():
    checks.check_kat(_libs.oracle())


def test_oracle_matches_golden_vectors_32():
    assert checks.check_golden(_libs.oracle(), bits=32) == {}


def test_oracle_matches_golden_vectors_64():
    assert checks.check_golden(_libs.oracle(), which=("small", "int"), bits=64) == {}


def test_oracle_gsa_matches_golden_vectors():
    assert checks.check_gsa(_libs.oracle(), 32) == {}
    assert checks.check_gsa(_libs.oracle(), 64) == {}


def test_reference_gsa_reproduces_golden_vectors():
    if _libs.ref() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    assert checks.check_gsa(_libs.ref(), 32) == {}


def test_oracle_error_codes_and_fast_paths():
    checks.check_errors(_libs.oracle())


def test_reference_error_codes_and_fast_paths():
    if _libs.ref() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    checks.check_errors(_libs.ref())


def test_reference_reproduces_golden_vectors():
    """The golden file really is what the compiled reference produces (guards against drift)."""
    if _libs.ref() is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    assert checks.check_golden(_libs.ref(), bits=32) == {}


def test_oracle_vs_bruteforce_definition():
    import ctypes as C
    o = _libs.oracle()
    rng = np.random.default_rng(7)
    fn = o.lib.oracle_check_sa_bruteforce
    fn.restype = C.c_int64
    for _ in range(300):
        n = int(rng.integers(1, 120))
        T = rng.integers(0, int(rng.choice([1, 2, 3, 256])), n).astype(np.uint8)
        rc, SA = o.sa(T, 64)
        assert rc == 0 and sorted(SA.tolist()) == list(range(n))
        assert fn(_libs.ptr(T), _libs.ptr(SA), C.c_int64(n)) == 0


def test_oracle_vs_reference_randomised():
    r = _libs.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    o = _libs.oracle()
    rng = np.random.default_rng(3)
    for _ in range(400):
        n = int(rng.integers(0, 400))
        sigma = int(rng.choice([1, 2, 4, 16, 256]))
        T = rng.integers(0, sigma, n).astype(np.uint8)
        for bits in (32, 64):
            a, b = o.sa(T, bits, want_freq=True), r.sa(T, bits, want_freq=True)
            assert a[0] == b[0] == 0 and (a[1] == b[1]).all() and (a[2] == b[2]).all()
            a, b = o.bwt(T, bits), r.bwt(T, bits)
            assert a[0] == b[0] and (a[1] == b[1]).all()
            if n:
                rr = int(rng.choice([2, 4, 32]))
                a2, b2 = o.bwt_aux(T, rr, bits), r.bwt_aux(T, rr, bits)
                assert a2[0] == b2[0] == 0 and (a2[1] == b2[1]).all() and (a2[2] == b2[2]).all()
                u = o.unbwt(a[1], a[0], bits)
                assert u[0] == 0 and (u[1] == T).all()
            SA = r.sa(T, bits)[1]
            p1, p2 = o.plcp(T, SA, bits), r.plcp(T, SA, bits)
            assert p1[0] == p2[0] == 0 and (p1[1] == p2[1]).all()
            l1, l2 = o.lcp(p1[1], SA, bits), r.lcp(p2[1], SA, bits)
            assert (l1[1] == l2[1]).all()


def test_oracle_medium_inputs_vs_reference():
    r = _libs.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    o = _libs.oracle()
    for name, T in cases.medium_cases().items():
        a, b = o.sa(T), r.sa(T)
        assert (a[1] == b[1]).all(), name
