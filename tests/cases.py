"""Seeded test inputs shared by the golden-vector generator, the oracle tests and the GPU parity tests."""
import numpy as np

from libsais_b200 import gen


def _b(s):
    return np.frombuffer(s, dtype=np.uint8).copy()


def fibonacci_string(k):
    a, b = b"b", b"a"
    for _ in range(k):
        a, b = b, b + a
    return _b(b)


def thue_morse(bits):
    t = np.zeros(1, dtype=np.uint8)
    for _ in range(bits):
        t = np.concatenate([t, 1 - t])
    return (t + ord("a")).astype(np.uint8)


def rnd(seed, n, sigma):
    rng = np.random.default_rng(seed)
    return rng.integers(0, sigma, n, dtype=np.int64).astype(np.uint8)


def small_cases():
    """name -> uint8 array; every edge the reference's API treats specially plus tie-heavy strings."""
    c = {}
    c["empty"] = _b(b"")
    c["one"] = _b(b"x")
    c["two_same"] = _b(b"aa")
    c["two_diff"] = _b(b"ba")
    c["banana"] = _b(b"banana")
    c["mississippi"] = _b(b"mississippi")
    c["a8"] = _b(b"aaaaaaaa")
    c["abracadabra"] = _b(b"abracadabra")
    c["zeros9"] = np.zeros(9, dtype=np.uint8)
    c["zeros_then_one"] = np.array([0] * 17 + [1], dtype=np.uint8)
    c["one_then_zeros"] = np.array([1] + [0] * 17, dtype=np.uint8)
    c["ff_run"] = np.full(33, 255, dtype=np.uint8)
    c["abab"] = _b(b"ab" * 37)
    c["aab_runs"] = _b(b"aab" * 50 + b"aa")
    c["fib12"] = fibonacci_string(12)
    c["tm8"] = thue_morse(8)
    for n in (3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 257):
        c["r2_%d" % n] = rnd(100 + n, n, 2)
        c["r4_%d" % n] = rnd(200 + n, n, 4)
        c["r256_%d" % n] = rnd(300 + n, n, 256)
    c["r1_100"] = rnd(1, 100, 1)
    c["r3_1000"] = rnd(2, 1000, 3)
    c["r256_1000"] = rnd(3, 1000, 256)
    c["zero_tail"] = np.concatenate([rnd(4, 300, 3), np.zeros(70, dtype=np.uint8)])
    c["zero_mid"] = np.concatenate([rnd(5, 100, 2), np.zeros(100, dtype=np.uint8), rnd(6, 100, 2)])
    return c


def medium_cases():
    """Inputs of a few 10^4..10^6 symbols; golden file keeps only digests of their outputs."""
    c = {}
    c["dna_64k"] = gen.dna(1, 1 << 16)
    c["bytes_64k"] = gen.rand_bytes(2, 1 << 16)
    c["rep_dna_19k_x20"] = gen.repetitive_dna(19000, 20)
    c["a_5000"] = np.full(5000, ord("a"), dtype=np.uint8)
    c["abab_9001"] = np.resize(_b(b"ab"), 9001)
    c["fib22"] = fibonacci_string(22)
    c["tm14"] = thue_morse(14)
    c["zeros_40000"] = np.zeros(40000, dtype=np.uint8)
    c["period7_30000"] = np.resize(rnd(7, 7, 4), 30000)
    c["r2_50000"] = rnd(8, 50000, 2)
    c["text_like"] = (rnd(9, 60000, 27) + 96).astype(np.uint8)
    return c


def int_cases():
    """name -> (int array, k) for libsais_int / libsais_plcp_int."""
    rng = np.random.default_rng(11)
    c = {}
    c["kat"] = (np.array([2, 1, 3, 1, 3, 1], dtype=np.int64), 4)
    c["k2_500"] = (rng.integers(0, 2, 500), 2)
    c["k5_3000"] = (rng.integers(0, 5, 3000), 5)
    c["k1000_4000"] = (rng.integers(0, 1000, 4000), 1000)
    c["k70000_20000"] = (rng.integers(0, 70000, 20000), 70000)
    c["k2e9_5000"] = (rng.integers(0, 2_000_000_000, 5000), 2_000_000_000)
    c["const_300"] = (np.full(300, 7), 8)
    c["ramp_down"] = (np.arange(2000, 0, -1), 2001)
    c["period3"] = (np.resize(np.array([5, 5, 9]), 4001), 10)
    return c


def gsa_cases():
    """name -> uint8 array ending in 0: collections of 0-separated strings (libsais_gsa)."""
    rng = np.random.default_rng(21)
    c = {}
    c["kat"] = _b(b"ab\0ab\0b\0")
    c["only_sep"] = np.zeros(1, dtype=np.uint8)
    c["one_string"] = _b(b"banana\0")
    c["dup_strings"] = _b(b"abc\0abc\0abc\0")

    def coll(count, maxlen, sigma):
        parts = []
        for _ in range(count):
            ln = int(rng.integers(1, maxlen))       # the reference rejects empty members
            parts.append((rng.integers(0, sigma, ln) + 1).astype(np.uint8))
            parts.append(np.zeros(1, dtype=np.uint8))
        return np.concatenate(parts)
    c["coll_small"] = coll(20, 12, 2)
    c["coll_dna"] = coll(300, 200, 4)
    c["coll_bytes"] = coll(100, 500, 255)
    c["coll_many_short"] = coll(5000, 6, 3)
    c["coll_repeats"] = np.concatenate([np.resize(np.array([1, 2, 1, 3, 0], dtype=np.uint8), 20000)])
    return c
