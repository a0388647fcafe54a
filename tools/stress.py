"""Seeded differential stress of the CUDA library against the oracle (run on a GPU box):
   python tools/stress.py [seconds] [seed] [big]
("big": texts of 2^20 .. 2^24 symbols made of copies / runs / repeats only -- many tiles per local-sort round)
Random texts with structure that exercises every route through the prefix doubling -- many small groups (local
counting sort), groups of hundreds to thousands (in-tile radix), giant groups (global onesweep), few active
suffixes (lazy ISA) -- with the route-forcing environment knobs flipped at random.  Compares SA, BWT + primary
and, for a sample of cases, PLCP / LCP / unBWT.  Prints one JSON line; exit code 1 on the first mismatch."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs  # noqa: E402


def make_text(rng, big=False):
    kind = rng.integers(0, 7)
    n = int(2 ** rng.uniform(4, 21.5))
    if big:
        kind = int(rng.choice([1, 1, 1, 3, 4]))
        n = int(2 ** rng.uniform(20, 24))
    sigma = int(rng.choice([1, 2, 3, 4, 5, 16, 26, 64, 200, 256]))
    if kind == 0:                                   # iid
        T = rng.integers(0, sigma, n)
    elif kind == 1:                                 # mutated copies of a base (many groups of `copies` elements)
        copies = int(rng.choice([2, 3, 8, 50, 100, 130, 300, 700, 2100]))
        base = rng.integers(0, sigma, max(n // copies, 4))
        T = np.tile(base, copies)
        mut = rng.random(len(T)) < 10 ** rng.uniform(-4, -1.5)
        T = np.where(mut, rng.integers(0, max(sigma, 2), len(T)), T)
    elif kind == 2:                                 # periodic
        T = np.resize(rng.integers(0, sigma, int(rng.integers(1, 50))), n)
    elif kind == 3:                                 # long runs inside random text
        T = rng.integers(0, sigma, n)
        for _ in range(int(rng.integers(1, 5))):
            a = int(rng.integers(0, n)); ln = int(rng.integers(1, max(n // 3, 2)))
            T[a:a + ln] = rng.integers(0, sigma)
    elif kind == 4:                                 # exact repeats of a long block
        blk = rng.integers(0, sigma, max(n // int(rng.integers(2, 6)), 2))
        T = np.concatenate([blk, rng.integers(0, sigma, int(rng.integers(0, 100))), blk, blk[: len(blk) // 2]])
    elif kind == 5:                                 # skewed (Zipf-like) symbols
        T = np.minimum(rng.zipf(1.5, n) - 1, sigma - 1)
    else:                                           # trailing / leading zeros around random text
        T = np.concatenate([np.zeros(int(rng.integers(0, 40)), dtype=np.int64), rng.integers(0, sigma, n),
                            np.zeros(int(rng.integers(0, 40)), dtype=np.int64)])
    T = np.ascontiguousarray(T, dtype=np.uint8)
    if rng.random() < 0.3 and sigma < 200:
        T = (T + int(rng.integers(0, 256 - sigma))).astype(np.uint8)      # shift the alphabet
    return kind, T


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    big = len(sys.argv) > 3 and sys.argv[3] == "big"
    rng = np.random.default_rng(seed)
    cu, o = _libs.cuda(), _libs.oracle()
    t0, cases, symbols = time.time(), 0, 0
    while time.time() - t0 < budget:
        kind, T = make_text(rng, big)
        env = {}
        if rng.random() < 0.5: env["LIBSAIS_CUDA_LOCAL_SORT"] = str(int(rng.integers(0, 2)))
        if rng.random() < 0.5: env["LIBSAIS_CUDA_LAZY_ISA"] = str(int(rng.integers(0, 2)))
        if rng.random() < 0.2: env["LIBSAIS_CUDA_KEY_SYMBOLS"] = str(int(rng.integers(1, 9)))
        if rng.random() < 0.3: env["LIBSAIS_CUDA_LAZY_LOCAL"] = str(int(rng.integers(0, 2)))
        if rng.random() < 0.3: env["LIBSAIS_CUDA_PO"] = str(int(rng.integers(0, 2)))
        if rng.random() < 0.6: env["LIBSAIS_CUDA_LOCAL_MIN"] = str(int(rng.choice([1, 64, 4096])))
        if rng.random() < 0.3: env["LIBSAIS_CUDA_PO_BIN"] = str(int(rng.integers(0, 12)))
        for k in ("LIBSAIS_CUDA_LOCAL_SORT", "LIBSAIS_CUDA_LAZY_ISA", "LIBSAIS_CUDA_KEY_SYMBOLS", "LIBSAIS_CUDA_PO", "LIBSAIS_CUDA_LOCAL_MIN", "LIBSAIS_CUDA_PO_BIN", "LIBSAIS_CUDA_LAZY_LOCAL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        what = None
        rc, SA = cu.sa(T)
        rco, SAo = o.sa(T)
        if rc != rco or not (SA == SAo).all(): what = "sa"
        if what is None:
            a, b = cu.bwt(T), o.bwt(T)
            if a[0] != b[0] or not (a[1] == b[1]).all(): what = "bwt"
        if what is None and cases % 4 == 0 and len(T) > 1:
            p1, p2 = cu.plcp(T, SAo), o.plcp(T, SAo)
            if p1[0] != p2[0] or not (p1[1] == p2[1]).all(): what = "plcp"
            l1 = cu.lcp(p2[1], SAo)
            if what is None and (l1[0] != 0 or not (l1[1] == p2[1][SAo]).all()): what = "lcp"
            u = cu.unbwt(b[1], b[0])
            if what is None and (u[0] != 0 or not (u[1] == T).all()): what = "unbwt"
        if what is not None:
            np.save(os.path.join(ROOT, "gpurun_out", "stress_fail.npy"), T)
            print(json.dumps({"stress": "FAIL", "what": what, "case": cases, "kind": int(kind), "n": len(T), "env": env, "seed": seed}))
            return 1
        cases += 1; symbols += len(T)
    print(json.dumps({"stress": "ok", "cases": cases, "symbols": symbols, "seconds": round(time.time() - t0, 1), "seed": seed}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
