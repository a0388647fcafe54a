"""GPU parity tests: the CUDA path, called through the C-ABI of libsais_cuda.so exactly as a C
program would call the reference, must be bit-exact with the committed golden vectors, with
the oracle, and (where oracle/_ref travelled to the box) with the unmodified reference."""
import os

import numpy as np
import pytest

import _libs
import cases
import checks
from libsais_b200 import gen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    import libsais_b200
    assert libsais_b200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return _libs.cuda()


def test_known_answers(cu):
    checks.check_kat(cu)


def test_error_codes_and_fast_paths(cu):
    checks.check_errors(cu)


def test_golden_small_32(cu):
    assert checks.check_golden(cu, which=("small",), bits=32) == {}


def test_golden_small_64(cu):
    assert checks.check_golden(cu, which=("small",), bits=64) == {}


def test_golden_medium(cu):
    assert checks.check_golden(cu, which=("medium",), bits=32) == {}


def test_golden_int_alphabets(cu):
    assert checks.check_golden(cu, which=("int",), bits=32) == {}
    assert checks.check_golden(cu, which=("int",), bits=64) == {}


def test_generalized_suffix_arrays(cu):
    """libsais_gsa / libsais_plcp_gsa (+64-bit): golden vectors and a larger random collection vs the oracle."""
    assert checks.check_gsa(cu, 32) == {}
    assert checks.check_gsa(cu, 64) == {}
    rng = np.random.default_rng(8)
    parts = []
    for _ in range(20000):
        parts.append((rng.integers(0, 4, int(rng.integers(1, 120))) + 65).astype(np.uint8))
        parts.append(np.zeros(1, dtype=np.uint8))
    T = np.concatenate(parts)
    o = _libs.oracle()
    a, b = cu.gsa(T), o.gsa(T)
    assert a[0] == b[0] == 0 and (a[1] == b[1]).all()
    p1, p2 = cu.plcp_gsa(T, b[1]), o.plcp_gsa(T, b[1])
    assert p1[0] == p2[0] == 0 and (p1[1] == p2[1]).all()


def _full_compare(cu, other, T, bits=32, aux_r=64):
    a, b = cu.sa(T, bits, want_freq=True), other.sa(T, bits, want_freq=True)
    assert a[0] == b[0] == 0
    assert (a[1] == b[1]).all(), "SA differs at %d" % int(np.argmax(a[1] != b[1]))
    assert (a[2] == b[2]).all()
    SA = b[1]
    a, b = cu.bwt(T, bits), other.bwt(T, bits)
    assert a[0] == b[0] and (a[1] == b[1]).all()
    U, primary = b[1], b[0]
    a2, b2 = cu.bwt_aux(T, aux_r, bits), other.bwt_aux(T, aux_r, bits)
    assert a2[0] == b2[0] == 0 and (a2[1] == b2[1]).all() and (a2[2] == b2[2]).all()
    u = cu.unbwt(U, primary, bits)
    assert u[0] == 0 and (u[1] == T).all()
    u = cu.unbwt_aux(U, aux_r, b2[2], bits)
    assert u[0] == 0 and (u[1] == T).all()
    p1, p2 = cu.plcp(T, SA, bits), other.plcp(T, SA, bits)
    assert p1[0] == p2[0] == 0 and (p1[1] == p2[1]).all()
    l1, l2 = cu.lcp(p2[1], SA, bits), other.lcp(p2[1], SA, bits)
    assert l1[0] == l2[0] == 0 and (l1[1] == l2[1]).all()


def test_random_differential_vs_oracle(cu):
    o = _libs.oracle()
    rng = np.random.default_rng(2024)
    for it in range(150):
        n = int(rng.integers(2, 3000))
        sigma = int(rng.choice([1, 2, 3, 4, 5, 16, 17, 64, 128, 129, 256]))
        T = rng.integers(0, sigma, n).astype(np.uint8)
        if it % 3 == 0:                          # long repeats: several doubling rounds
            period = int(rng.integers(1, 40))
            T = np.resize(T[:period], n).copy()
            if n > 10:
                T[int(rng.integers(0, n))] ^= 1
        _full_compare(cu, o, T, bits=32 if it % 2 else 64, aux_r=int(rng.choice([2, 16, 1024])))


def test_config1_dna_1mib_vs_oracle_and_reference(cu):
    """BASELINE config 1: libsais SA of a 1 MiB iid ACGT string (seed 1)."""
    T = gen.dna(1, 1 << 20)
    _full_compare(cu, _libs.oracle(), T)
    if _libs.ref() is not None:
        _full_compare(cu, _libs.ref(), T)


def test_random_bytes_4mib_vs_oracle(cu):
    """BASELINE config 2 at a size the oracle finishes in seconds."""
    T = gen.rand_bytes(2, 1 << 22)
    _full_compare(cu, _libs.oracle(), T)


def test_repetitive_dna_vs_oracle(cu):
    """BASELINE config 3 at 1/1000 scale: 100 mutated copies of a 19 000-base genome."""
    T = gen.repetitive_dna(19000, 100)
    _full_compare(cu, _libs.oracle(), T)


def test_repetitive_dna_19m_partitioned_scatter_vs_oracle(cu):
    """BASELINE config 3 at 1/100 scale (19 M symbols): large enough that the ISA updates and the
    phi array go through the locality-partitioned scatter and most suffixes stay active for ~10 rounds."""
    T = gen.repetitive_dna(190000, 100)
    o = _libs.oracle()
    a, b = cu.sa(T), o.sa(T)
    assert a[0] == b[0] == 0 and (a[1] == b[1]).all()
    SA = b[1]
    a, b = cu.bwt(T), o.bwt(T)
    assert a[0] == b[0] and (a[1] == b[1]).all()
    p1, p2 = cu.plcp(T, SA), o.plcp(T, SA)
    assert p1[0] == 0 and (p1[1] == p2[1]).all()
    l1 = cu.lcp(p2[1], SA)
    assert l1[0] == 0 and (l1[1] == p2[1][SA]).all()
    u = cu.unbwt(a[1], a[0])
    assert u[0] == 0 and (u[1] == T).all()


def test_forced_isa_modes_agree(cu):
    """Lazy ISA (binary-search fallback) and full ISA (partitioned scatter) are two routes to the same SA."""
    import os
    o = _libs.oracle()
    T = np.concatenate([gen.dna(5, 300000), gen.dna(5, 300000), gen.rand_bytes(6, 200000)])
    want = o.sa(T)[1]
    try:
        for mode in ("0", "1"):
            os.environ["LIBSAIS_CUDA_LAZY_ISA"] = mode
            rc, SA = cu.sa(T)
            assert rc == 0 and (SA == want).all(), mode
            rcb, U = cu.bwt(T)
            assert (rcb, U.tobytes()) == (o.bwt(T)[0], o.bwt(T)[1].tobytes()), mode
    finally:
        os.environ.pop("LIBSAIS_CUDA_LAZY_ISA", None)


def test_local_and_global_round_sorts_agree(cu):
    """Rounds >= 1 sort small groups inside shared memory (local_sort.cuh) or with the global onesweep: both
    must give the oracle's SA.  Inputs: many small groups (mutated copies), a mix where one long repeat makes a
    large group for a few rounds (global rounds first, local rounds later), and groups just around the tile size."""
    import os
    o = _libs.oracle()
    rep = gen.repetitive_dna(40000, 60)                                    # 2.4 M, groups of <= 60
    big = np.concatenate([gen.dna(11, 700000), np.zeros(5000, dtype=np.uint8) + 65, gen.dna(11, 700000), gen.dna(12, 300000)])
    blocky = np.tile(gen.dna(13, 1100), 2049)[: 2200000]                   # ~2049 copies: groups just above the tile size, shrinking at the end
    try:
        for T in (rep, big, blocky):
            want = o.sa(T)[1]
            for mode in ("1", "0"):
                os.environ["LIBSAIS_CUDA_LOCAL_SORT"] = mode
                os.environ["LIBSAIS_CUDA_LAZY_ISA"] = "0"
                rc, SA = cu.sa(T)
                assert rc == 0 and (SA == want).all(), (len(T), mode)
    finally:
        os.environ.pop("LIBSAIS_CUDA_LOCAL_SORT", None)
        os.environ.pop("LIBSAIS_CUDA_LAZY_ISA", None)


def test_seeded_stress_differential(cu, capsys):
    """15 s of tools/stress.py: random structured texts (copies, runs, periodic, skewed, exact repeats) with the
    route-forcing knobs flipped at random, SA / BWT (+ PLCP / LCP / unBWT on a sample) against the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "stress.py"), "15", "7"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and '"stress": "ok"' in r.stdout, (r.stdout[-500:], r.stderr[-500:])


def test_adversarial_periodic_inputs_vs_oracle(cu):
    o = _libs.oracle()
    for T in (np.zeros(1 << 17, dtype=np.uint8), np.resize(np.frombuffer(b"ab", dtype=np.uint8), (1 << 17) + 1),
              cases.fibonacci_string(24), cases.thue_morse(17),
              np.concatenate([gen.dna(7, 50000), gen.dna(7, 50000), gen.dna(7, 50000)])):
        _full_compare(cu, o, T.copy())


def test_int_alphabet_vs_oracle(cu):
    o = _libs.oracle()
    rng = np.random.default_rng(5)
    for k in (2, 3, 255, 256, 257, 65536, 1 << 20, (1 << 31) - 1):
        n = int(rng.integers(2, 20000))
        T = rng.integers(0, k, n)
        for bits in (32, 64):
            a, b = cu.sa_int(T, k, bits), o.sa_int(T, k, bits)
            assert a[0] == b[0] == 0 and (a[1] == b[1]).all(), (k, bits)
            assert (a[2] == np.asarray(T)).all()
        p1, p2 = cu.plcp(T.astype(np.int32), b[1]), o.plcp(T.astype(np.int32), b[1])
        assert (p1[1] == p2[1]).all()


def test_inplace_aliasing(cu):
    """U may alias T for bwt / unbwt; LCP may alias SA (include/libsais.h:175, :282, :394)."""
    T = gen.dna(3, 100000)
    rc, U = cu.bwt(T.copy())
    buf = T.copy()
    rc2, U2 = cu.bwt(buf, inplace=True)
    assert rc == rc2 and (U2 == U).all() and U2 is buf
    rc3, back = cu.unbwt(buf, rc, inplace=True)
    assert rc3 == 0 and (back == T).all()


def test_concurrent_host_threads_use_independent_contexts(cu):
    """Re-entrancy: ctx-less calls from different host threads use per-thread contexts and may run
    concurrently (reference contract: one context per thread, include/libsais.h:53-55)."""
    import threading
    o = _libs.oracle()
    inputs = [gen.dna(40 + i, 300000 + 1111 * i) for i in range(4)] + [gen.rand_bytes(50 + i, 200000 + 77 * i) for i in range(4)]
    want = [(o.sa(T)[1], o.bwt(T)) for T in inputs]
    results = [None] * len(inputs)

    def work(i):
        for _ in range(3):
            rc, SA = cu.sa(inputs[i])
            rcb, U = cu.bwt(inputs[i])
            results[i] = (rc, SA, rcb, U)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(inputs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i, (rc, SA, rcb, U) in enumerate(results):
        assert rc == 0 and (SA == want[i][0]).all()
        assert rcb == want[i][1][0] and (U == want[i][1][1]).all()


def test_device_pointer_entry_points_roundtrip(cu):
    import torch
    import libsais_b200
    ctx = libsais_b200.Context(0)
    T = gen.rand_bytes(9, 3_000_001)
    o = _libs.oracle()
    dT = torch.from_numpy(T).cuda()
    dSA = torch.empty(len(T), dtype=torch.int32, device="cuda")
    dU = torch.empty(len(T), dtype=torch.uint8, device="cuda")
    dP = torch.empty(len(T), dtype=torch.int32, device="cuda")
    dL = torch.empty(len(T), dtype=torch.int32, device="cuda")
    dBack = torch.empty(len(T), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert ctx.sa_dev(dT.data_ptr(), dSA.data_ptr(), len(T)) == 0
    primary = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), len(T))
    assert ctx.plcp_dev(dT.data_ptr(), dSA.data_ptr(), dP.data_ptr(), len(T)) == 0
    assert ctx.lcp_dev(dP.data_ptr(), dSA.data_ptr(), dL.data_ptr(), len(T)) == 0
    assert ctx.unbwt_dev(dU.data_ptr(), dBack.data_ptr(), len(T), primary) == 0
    rc, SA = o.sa(T)
    rcb, U = o.bwt(T)
    assert (dSA.cpu().numpy() == SA).all()
    assert primary == rcb and (dU.cpu().numpy() == U).all()
    assert (dP.cpu().numpy() == o.plcp(T, SA)[1]).all()
    assert (dL.cpu().numpy() == o.lcp(o.plcp(T, SA)[1], SA)[1]).all()
    assert (dBack.cpu().numpy() == T).all()
    st = ctx.stats()
    assert st["total_launches"] > 0
    ctx.close()


def _verify_sa_on_gpu(dT, dSA, n, chunk=1 << 27):
    """Linear-time SA check with torch ops on the GPU: permutation + Burkhardt-Kaerkkaeinen neighbour order."""
    import torch
    seen = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for lo in range(0, n, chunk):
        seen[dSA[lo:lo + chunk].long()] = 1
    assert bool(seen.all()), "SA is not a permutation"
    del seen
    ISA = torch.empty(n, dtype=torch.int32, device="cuda")
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        ISA[dSA[lo:hi].long()] = torch.arange(lo, hi, dtype=torch.int32, device="cuda")
    for lo in range(1, n, chunk):
        hi = min(n, lo + chunk)
        a = dSA[lo - 1:hi - 1].long(); b = dSA[lo:hi].long()
        ta, tb = dT[a], dT[b]
        ra = torch.where(a + 1 < n, ISA[torch.clamp(a + 1, max=n - 1)], torch.full_like(ISA[:1], -1))
        rb = torch.where(b + 1 < n, ISA[torch.clamp(b + 1, max=n - 1)], torch.full_like(ISA[:1], -1))
        assert bool(((ta < tb) | ((ta == tb) & (ra < rb))).all()), "suffix order violated"
    return ISA


def test_n_above_2pow30_device_api_properties(cu):
    """n = 2^30 + 12345 iid ACGT through the device-pointer API: exercises the 64-bit tile status
    words of the onesweep pass (n >= 2^30); checked by the linear-time SA verifier on the GPU and
    by unbwt(bwt(T)) == T."""
    import torch
    import libsais_b200
    n = (1 << 30) + 12345
    T = gen.dna(5, n)
    ctx = libsais_b200.Context(0)
    dT = torch.from_numpy(T).cuda()
    dSA = torch.empty(n, dtype=torch.int32, device="cuda")
    assert ctx.sa_dev(dT.data_ptr(), dSA.data_ptr(), n) == 0
    ISA = _verify_sa_on_gpu(dT, dSA, n)
    dU = torch.empty(n, dtype=torch.uint8, device="cuda")
    primary = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n)
    assert primary == int(ISA[0]) + 1
    del ISA, dSA
    dB = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert ctx.unbwt_dev(dU.data_ptr(), dB.data_ptr(), n, primary) == 0
    assert torch.equal(dB, dT)
    ctx.close()


def test_large_bwt_roundtrip_and_sa_properties(cu):
    """Size-independent properties at a size the oracle is too slow for (64 MiB random bytes):
    SA is a permutation, adjacent suffixes are ordered (checked on the device-produced ISA by the
    Burkhardt-Kaerkkaeinen neighbour rule on a sample), unbwt(bwt(T)) == T."""
    n = 1 << 26
    T = gen.rand_bytes(2, n)
    rc, SA = cu.sa(T)
    assert rc == 0
    seen = np.zeros(n, dtype=np.bool_)
    seen[SA] = True
    assert seen.all()
    ISA = np.empty(n, dtype=np.int64)
    ISA[SA] = np.arange(n)
    rng = np.random.default_rng(1)
    idx = rng.integers(1, n, 2_000_000)
    a, b = SA[idx - 1].astype(np.int64), SA[idx].astype(np.int64)
    ta, tb = T[a], T[b]
    ra = np.where(a + 1 < n, ISA[np.minimum(a + 1, n - 1)], -1)
    rb = np.where(b + 1 < n, ISA[np.minimum(b + 1, n - 1)], -1)
    assert ((ta < tb) | ((ta == tb) & (ra < rb))).all()
    rcb, U = cu.bwt(T)
    assert rcb == int(ISA[0]) + 1
    rcu, back = cu.unbwt(U, rcb)
    assert rcu == 0 and (back == T).all()
