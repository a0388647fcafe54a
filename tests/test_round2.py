"""GPU parity tests added in round 2: the round-0 MSD path (partition.cuh), PLCP on texts with very long
matches (the level-seeded compare), the batch entry point (BASELINE config 4) and the libsais64 host API
beyond INT32_MAX.  Everything goes through the C-ABI and is compared with the oracle / compiled reference."""
import ctypes as C
import os

import numpy as np
import pytest

import _libs
from libsais_b200 import gen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    import libsais_b200
    assert libsais_b200.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return _libs.cuda()


def _best_cpu():
    return _libs.ref() or _libs.oracle()


def test_msd_and_lsd_round0_paths_agree(cu):
    """Round 0 either partitions by the top 16 key bits and finishes the buckets in shared memory (MSD) or runs
    the stable LSD passes: both must produce the oracle's SA and BWT.  Sizes straddle the tile (4608 / 3840)
    and chunk boundaries; alphabets cover every code width the MSD path accepts (1, 2, 4, 8 bits)."""
    o = _best_cpu()
    rng = np.random.default_rng(77)
    texts = [gen.rand_bytes(3, n) for n in (64, 100, 4607, 4608, 4609, 3840 * 3 + 1, 70_001, 1 << 20)]
    texts += [gen.dna(7, n) for n in (4609, 300_000)]
    texts += [(rng.integers(0, 16, 200_000) + 97).astype(np.uint8), (rng.integers(0, 2, 100_000) + 48).astype(np.uint8)]
    texts += [np.concatenate([gen.rand_bytes(8, 50_000), np.zeros(37, dtype=np.uint8)]),       # zero tail: end-of-text rule
              np.concatenate([np.zeros(9, dtype=np.uint8), gen.rand_bytes(9, 50_000)]),
              np.tile(gen.rand_bytes(9, 20_000), 5)]                                            # exact repeats: many ties per bucket
    try:
        for T in texts:
            rs, SAr = o.sa(T)
            rb, Ur = o.bwt(T)
            for mode in ("2", "0"):
                os.environ["LIBSAIS_CUDA_MSD"] = mode
                rc, SA = cu.sa(T)
                assert rc == 0 and (SA == SAr).all(), (len(T), mode)
                rcb, U = cu.bwt(T)
                assert rcb == rb and (U == Ur).all(), (len(T), mode)
                _, U2, I = cu.bwt_aux(T, 64)
                assert (U2 == Ur).all() and (I == o.bwt_aux(T, 64)[2]).all(), (len(T), mode)
    finally:
        os.environ.pop("LIBSAIS_CUDA_MSD", None)


def test_plcp_lcp_on_long_matches(cu):
    """PLCP / LCP where neighbouring suffixes share thousands to millions of symbols (a^n, (ab)^n, a long period,
    two identical halves, mutated copies): the compare must stay bounded (level seeding) and bit-exact."""
    o = _best_cpu()
    n = 1 << 20
    texts = {"zeros": np.zeros(n, dtype=np.uint8), "abab": np.resize(np.frombuffer(b"ab", dtype=np.uint8), n + 1),
             "period_1000": np.resize(gen.dna(3, 1000), n), "two_copies": np.concatenate([gen.dna(4, n // 2), gen.dna(4, n // 2)]),
             "mutated_copies": gen.repetitive_dna(n // 64, 64), "tiny": np.frombuffer(b"banana", dtype=np.uint8).copy(),
             "n33": np.resize(np.frombuffer(b"abc", dtype=np.uint8), 33), "n32": np.zeros(32, dtype=np.uint8)}
    for name, T in texts.items():
        for bits in (32, 64):
            SA = o.sa(T, bits)[1]
            p1, p2 = cu.plcp(T, SA, bits), o.plcp(T, SA, bits)
            assert p1[0] == p2[0] == 0 and (p1[1] == p2[1]).all(), (name, bits)
            l1, l2 = cu.lcp(p2[1], SA, bits), o.lcp(p2[1], SA, bits)
            assert l1[0] == l2[0] == 0 and (l1[1] == l2[1]).all(), (name, bits)
        Ti = T.astype(np.int32) * 1000 + 7
        p1, p2 = cu.plcp(Ti, SA), _libs.oracle().plcp(Ti, SA)
        assert (p1[1] == p2[1]).all(), name


def test_bwt_batch_entry_point(cu):
    """libsais_cuda_bwt_batch (BASELINE config 4's mechanism): 10 blocks of different sizes and alphabets through
    the pooled contexts (3 host threads on the GPU), every block compared with the oracle; then the same blocks
    with every visible GPU in the device list."""
    import libsais_b200
    lib = libsais_b200.load_library()
    o = _best_cpu()
    blocks = [gen.dna(1000 + b, 200_000 + 4099 * b) for b in range(6)] + [gen.rand_bytes(2000 + b, 150_000 + 777 * b) for b in range(4)]
    blocks.append(np.frombuffer(b"x", dtype=np.uint8).copy())          # n = 1 fast path inside a batch
    k = len(blocks)
    outs = [np.zeros(len(b), dtype=np.uint8) for b in blocks]
    Tp = (C.c_void_p * k)(*[b.ctypes.data for b in blocks])
    Up = (C.c_void_p * k)(*[u.ctypes.data for u in outs])
    ns = (C.c_int32 * k)(*[len(b) for b in blocks])
    pr = (C.c_int32 * k)()
    ms = (C.c_float * k)()
    lib.libsais_cuda_bwt_batch.restype = C.c_int32
    ndev = libsais_b200.device_count()
    for devs in ([0], list(range(ndev)), None):
        for u in outs:
            u[:] = 0
        dv = None if devs is None else (C.c_int32 * len(devs))(*devs)
        rc = lib.libsais_cuda_bwt_batch(Tp, Up, ns, pr, ms, C.c_int32(k), dv, C.c_int32(0 if devs is None else len(devs)), C.c_int32(3))
        assert rc == 0
        for i, b in enumerate(blocks):
            want = o.bwt(b)
            assert int(pr[i]) == want[0] and (outs[i] == want[1]).all(), (devs, i)
    assert lib.libsais_cuda_bwt_batch(None, Up, ns, pr, ms, C.c_int32(k), None, C.c_int32(0), C.c_int32(0)) == -1
    bad = (C.c_int32 * 1)(ndev + 5)
    assert lib.libsais_cuda_bwt_batch(Tp, Up, ns, pr, ms, C.c_int32(k), bad, C.c_int32(1), C.c_int32(0)) == -1
    lib.libsais_cuda_batch_release()


def test_libsais64_host_api_above_int32_max(cu):
    """libsais64 through the HOST API with n = 2^31 + 1024 > INT32_MAX (iid ACGT, seed 5): the reference takes its
    native 64-bit path (src/libsais64.c:7085); compared element by element with libsais64_omp of the compiled
    reference when it travelled to this box (a few minutes of CPU), else checked by the linear-time verifier."""
    import torch
    import libsais_b200
    from libsais_b200 import check
    free, _ = torch.cuda.mem_get_info()
    if free < 150e9:
        pytest.skip("needs ~130 GB of HBM")
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 0
    if avail < 60e9:
        pytest.skip("needs ~45 GB of host memory")
    n = (1 << 31) + 1024
    T = gen.dna_torch(5, n, device="cuda").cpu().numpy()
    lib = libsais_b200.load_library()
    lib.libsais64.restype = C.c_int64
    SA = np.empty(n, dtype=np.int64)
    freq = np.zeros(256, dtype=np.int64)
    rc = lib.libsais64(T.ctypes.data_as(C.c_void_p), SA.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int64(0), freq.ctypes.data_as(C.c_void_p))
    assert rc == 0 and int(freq.sum()) == n
    lib.libsais_cuda_release_workspace(None)
    r = _libs.ref()
    if r is not None and avail > 90e9:
        SAr = np.empty(n, dtype=np.int64)
        f = r.lib.libsais64_omp
        f.restype = C.c_int64
        assert f(T.ctypes.data_as(C.c_void_p), SAr.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int64(0), None, C.c_int64(0)) == 0
        assert np.array_equal(SA, SAr)
    else:
        dT = torch.from_numpy(T).cuda()
        d32 = torch.empty(n, dtype=torch.int32, device="cuda")
        ch = 1 << 27
        for lo in range(0, n, ch):
            part = torch.from_numpy(SA[lo:lo + ch]).cuda()
            assert int(part.min()) >= 0 and int(part.max()) < n
            d32[lo:lo + ch] = part.to(torch.int32)                     # positions >= 2^31 wrap; the checker masks them back
        assert check.verify_sa_u32(dT, d32, n) == "ok"


class _DistStats(C.Structure):
    _fields_ = [("n_gpus", C.c_int32), ("rounds", C.c_int32), ("key_symbols", C.c_int32), ("key_bits", C.c_int32),
                ("slice_max", C.c_uint64), ("active_after_round0", C.c_uint64), ("exchanged_bytes", C.c_uint64),
                ("seconds_total", C.c_double), ("seconds_device", C.c_double), ("verify", C.c_int32), ("reserved", C.c_int32),
                ("verify_violations", C.c_uint64), ("phase_seconds", C.c_double * 8)]


def test_distributed_prefix_doubling_in_library(cu):
    """dist64.cu (BASELINE config 5's mechanism, what libsais64 runs beyond the single-GPU limit): G ranks -- one host
    thread each, sharing GPUs when the box has fewer -- must return the reference's suffix array, and the distributed
    checker must accept it.  Texts: iid (one doubling round), mutated copies / periodic / a^n (every suffix stays
    unresolved for many rounds), tiny.  Also through libsais64() itself with $LIBSAIS_CUDA_DIST."""
    import libsais_b200
    lib = libsais_b200.load_library()
    lib.libsais_cuda_sa64_multi.restype = C.c_int64
    o = _best_cpu()
    texts = [gen.dna(5, 1 << 19), gen.rand_bytes(2, 200_003), gen.repetitive_dna(10_000, 40), np.zeros(20_000, dtype=np.uint8),
             np.resize(np.frombuffer(b"abracadabra", dtype=np.uint8), 100_003), np.frombuffer(b"mississippi", dtype=np.uint8).copy()]
    for T in texts:
        want = o.sa(T, 64)[1]
        for G in (1, 2, 3, 5):
            SA = np.full(len(T), -1, dtype=np.int64)
            freq = np.zeros(256, dtype=np.int64)
            st = _DistStats(); st.verify = 1
            rc = lib.libsais_cuda_sa64_multi(T.ctypes.data_as(C.c_void_p), SA.ctypes.data_as(C.c_void_p), C.c_int64(len(T)),
                                             freq.ctypes.data_as(C.c_void_p), None, C.c_int32(G), C.byref(st))
            assert rc == 0 and (SA == want).all(), (len(T), G)
            assert st.verify == 1 and st.n_gpus == G
            assert (freq == np.bincount(T, minlength=256)).all()
    try:
        os.environ["LIBSAIS_CUDA_DIST"] = "2"
        T = texts[0]
        rc, SA = cu.sa(T, 64)
        assert rc == 0 and (SA == o.sa(T, 64)[1]).all()
    finally:
        os.environ.pop("LIBSAIS_CUDA_DIST", None)
    assert lib.libsais_cuda_sa64_multi(None, None, C.c_int64(5), None, None, C.c_int32(2), None) == -1
    assert lib.libsais_cuda_sa64_multi(texts[0].ctypes.data_as(C.c_void_p), None, C.c_int64(len(texts[0])), None, None, C.c_int32(0), None) == -1


def test_distributed_bwt_through_libsais64_bwt(cu):
    """libsais64_bwt / libsais64_bwt_aux on the multi-GPU path ($LIBSAIS_CUDA_DIST forces it at small n): BWT rows come from
    the distributed SA slices and the replicated packed text, aux samples from the ISA slices; in place (U == T) too."""
    o = _best_cpu()
    texts = [gen.dna(6, 300_001), gen.rand_bytes(3, 100_000), gen.repetitive_dna(5_000, 30), np.zeros(5_000, dtype=np.uint8)]
    try:
        for G in ("1", "2", "3"):
            os.environ["LIBSAIS_CUDA_DIST"] = G
            for T in texts:
                want = o.bwt(T, 64)
                got = cu.bwt(T, 64, want_freq=True)
                assert got[0] == want[0] and (got[1] == want[1]).all(), (G, len(T))
                assert (got[2] == np.bincount(T, minlength=256)).all()
                wa = o.bwt_aux(T, 256, 64)
                ga = cu.bwt_aux(T, 256, 64)
                assert ga[0] == 0 and (ga[1] == wa[1]).all() and (ga[2] == wa[2]).all(), (G, len(T))
                buf = T.copy()
                gi = cu.bwt(buf, 64, inplace=True)
                assert gi[0] == want[0] and (gi[1] == want[1]).all()
    finally:
        os.environ.pop("LIBSAIS_CUDA_DIST", None)


def test_unbwt_aux_uses_the_samples(cu):
    """libsais_unbwt_aux decodes every block of r symbols as an independent chain that starts at the sampled row I[j]
    (reference src/libsais.c:7943-7973 decodes the blocks independently too).  Parity on several r, sizes that are not
    multiples of r, both index widths; and a proof that the samples are really what drives the decoding: with I[1] and I[2]
    exchanged the output has exactly those two blocks exchanged."""
    o = _best_cpu()
    for T in (gen.dna(21, 100_000), gen.rand_bytes(22, 65_537), gen.repetitive_dna(2_000, 30), np.zeros(5_000, dtype=np.uint8)):
        for r in (2, 64, 256, 4096):
            for bits in (32, 64):
                rc, U, I = o.bwt_aux(T, r, bits)
                got = cu.unbwt_aux(U, r, I, bits)
                assert got[0] == 0 and (got[1] == T).all(), (len(T), r, bits)
    T = gen.dna(23, 10 * 256)
    r = 256
    rc, U, I = o.bwt_aux(T, r)
    J = I.copy(); J[1], J[2] = I[2], I[1]
    got = cu.unbwt_aux(U, r, J)
    want = T.copy(); want[r:2 * r], want[2 * r:3 * r] = T[2 * r:3 * r], T[r:2 * r]
    assert got[0] == 0 and (got[1] == want).all()


def test_position_ordered_rounds(cu):
    """Rounds >= 1 on the position-ordered active list (po_rounds.cuh) against the slot-ordered rounds of round 1
    and the oracle: mutated copies (groups of `copies` suffixes that take many rounds), exact repeats (ties that
    only the end of the text breaks), runs, texts whose groups straddle the 128 / 512 / 1024 limits, small lists
    (LIBSAIS_CUDA_LOCAL_MIN lowers the threshold so short texts take the path), several update-bin widths.
    SA, BWT + primary index and the aux samples must be bit-exact."""
    o = _best_cpu()
    rng = np.random.default_rng(2024)
    texts = {"copies100": gen.repetitive_dna(20_000, 100), "copies7": gen.repetitive_dna(150_000, 7, rate_num=300),
             "copies130": gen.repetitive_dna(3_000, 130), "copies600": gen.repetitive_dna(1_500, 600, rate_num=4000),
             "copies1000": gen.repetitive_dna(700, 1000, rate_num=8000),
             "exact_repeats": np.tile(gen.dna(5, 30_000), 9), "tail_repeat": np.concatenate([gen.dna(6, 100_000), gen.dna(6, 100_000)[:70_000]]),
             "bytes_copies": np.tile(gen.rand_bytes(4, 50_000), 4), "runs": np.repeat(gen.dna(8, 40_000), 5),
             "small": gen.repetitive_dna(300, 20), "binary": (rng.integers(0, 2, 300_000) + 48).astype(np.uint8)}
    # groups of 3000 after round 0 (a 40-symbol motif) that split into groups of ~300 (ten 300-symbol continuations): slot-ordered
    # rounds first, the switch to the position-ordered rounds once no group exceeds 1024
    motif, conts = gen.dna(21, 40), [gen.dna(30 + i, 300) for i in range(10)]
    texts["late_switch"] = np.concatenate([np.concatenate([motif, conts[int(rng.integers(0, 10))], gen.dna(100 + i, int(rng.integers(5, 60)))]) for i in range(3000)])
    texts["mutated_bytes"] = texts["bytes_copies"].copy()
    texts["mutated_bytes"][rng.integers(0, len(texts["mutated_bytes"]), 300)] = 7
    knobs = ("LIBSAIS_CUDA_PO", "LIBSAIS_CUDA_LOCAL_MIN", "LIBSAIS_CUDA_PO_BIN")
    try:
        for name, T in texts.items():
            rs, SAr = o.sa(T)
            rb, Ur = o.bwt(T)
            Ir = o.bwt_aux(T, 128)[2]
            for po, lmin, drop in (("1", "1", None), ("1", "1", "0"), ("1", "4096", "9"), ("0", "1", None)):
                for k in knobs:
                    os.environ.pop(k, None)
                os.environ["LIBSAIS_CUDA_PO"] = po; os.environ["LIBSAIS_CUDA_LOCAL_MIN"] = lmin
                if drop is not None:
                    os.environ["LIBSAIS_CUDA_PO_BIN"] = drop
                rc, SA = cu.sa(T)
                assert rc == 0 and (SA == SAr).all(), (name, po, lmin, drop)
                rcb, U = cu.bwt(T)
                assert rcb == rb and (U == Ur).all(), (name, po, lmin, drop)
                _, U2, I = cu.bwt_aux(T, 128)
                assert (U2 == Ur).all() and (I == Ir).all(), (name, po, lmin, drop)
    finally:
        for k in knobs:
            os.environ.pop(k, None)


def test_bwt_streamed_rows_into_pinned_buffers(cu):
    """libsais_bwt / libsais_bwt_aux with PINNED caller buffers: the rows of the slots that round 0 settles leave for the
    host while the remaining buckets are sorted, the bucket of suffix 0 and the late rows follow (api.cu bwt_body,
    sa_core.cu streamed rows).  Same bytes and primary index as the oracle -- for texts whose suffix 0 is settled in
    round 0, texts where it is not (a prefix that occurs twice), texts that fall back to the one-piece copy (everything
    repeated), and with the streaming switched off."""
    import torch
    import libsais_b200
    lib = libsais_b200.load_library()
    o = _best_cpu()
    rng = np.random.default_rng(5)
    R = gen.rand_bytes(11, 3_000_000)
    twice = R.copy(); twice[1_500_000:1_500_040] = twice[:40]                 # suffix 0 shares 40 bytes with suffix 1.5e6
    texts = {"bytes_3M": R, "dna_6M": gen.dna(12, 6_000_000), "prefix_twice": twice,
             "sigma16": (rng.integers(0, 16, 2_000_000) + 65).astype(np.uint8), "all_repeated": np.tile(gen.rand_bytes(13, 700_000), 2),
             "zero_head": np.concatenate([np.zeros(5, dtype=np.uint8), gen.rand_bytes(14, 1_000_000)])}
    lib.libsais_bwt.restype = C.c_int32
    lib.libsais_bwt_aux.restype = C.c_int32
    try:
        for name, T in texts.items():
            n = len(T)
            rb, Ur = o.bwt(T)
            Ir = o.bwt_aux(T, 256)[2]
            for stream in ("1", "0"):
                os.environ["LIBSAIS_CUDA_STREAM_ROWS"] = stream
                Tp = torch.from_numpy(T).pin_memory()
                Up = torch.full((n,), 0xEE, dtype=torch.uint8).pin_memory()
                Ap = torch.empty(n, dtype=torch.int32).pin_memory()
                rc = lib.libsais_bwt(C.c_void_p(Tp.data_ptr()), C.c_void_p(Up.data_ptr()), C.c_void_p(Ap.data_ptr()), C.c_int32(n), C.c_int32(0), None)
                assert rc == rb and (Up.numpy() == Ur).all(), (name, stream, rc, rb)
                Up.fill_(0xEE)
                Ip = torch.zeros((n - 1) // 256 + 1, dtype=torch.int32).pin_memory()
                rc = lib.libsais_bwt_aux(C.c_void_p(Tp.data_ptr()), C.c_void_p(Up.data_ptr()), C.c_void_p(Ap.data_ptr()), C.c_int32(n), C.c_int32(0), None,
                                         C.c_int32(256), C.c_void_p(Ip.data_ptr()))
                assert rc == 0 and (Up.numpy() == Ur).all() and (Ip.numpy() == Ir).all(), (name, stream)
                # in place: U is T (allowed by include/libsais.h)
                Tq = torch.from_numpy(T.copy()).pin_memory()
                rc = lib.libsais_bwt(C.c_void_p(Tq.data_ptr()), C.c_void_p(Tq.data_ptr()), C.c_void_p(Ap.data_ptr()), C.c_int32(n), C.c_int32(0), None)
                assert rc == rb and (Tq.numpy() == Ur).all(), (name, stream, "in place")
    finally:
        os.environ.pop("LIBSAIS_CUDA_STREAM_ROWS", None)
