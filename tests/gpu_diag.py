"""Diagnostic sweep for the first GPU runs: prints, per case and per API, pass/fail plus the
first mismatch -- more information per gpurun call than `pytest -x`.  Not a pytest module."""
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import _libs  # noqa: E402
import cases  # noqa: E402
from libsais_b200 import gen  # noqa: E402
import libsais_b200  # noqa: E402


def first_diff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return "shape %s vs %s" % (a.shape, b.shape)
    d = np.nonzero(a != b)[0]
    return None if len(d) == 0 else "first@%d got=%s want=%s ndiff=%d" % (d[0], a[d[0]:d[0] + 4], b[d[0]:d[0] + 4], len(d))


def run_case(name, T, cu, o, verbose=False):
    res = []
    try:
        t0 = time.time()
        rc, SA = cu.sa(T)
        dt = time.time() - t0
        rco, SAo = o.sa(T)
        d = first_diff(SA, SAo)
        res.append(("sa", rc == 0 and d is None, "rc=%d %s" % (rc, d)))
        rc, U = cu.bwt(T)
        rco, Uo = o.bwt(T)
        d = first_diff(U, Uo)
        res.append(("bwt", rc == rco and d is None, "rc=%d/%d %s" % (rc, rco, d)))
        if len(T):
            rc, U2, I = cu.bwt_aux(T, 8)
            rco, U2o, Io = o.bwt_aux(T, 8)
            d = first_diff(I, Io)
            res.append(("aux", rc == 0 and d is None, "rc=%d %s" % (rc, d)))
            rc, back = cu.unbwt(Uo, rco if False else o.bwt(T)[0])
            d = first_diff(back, T)
            res.append(("unbwt", rc == 0 and d is None, "rc=%d %s" % (rc, d)))
            rc, P = cu.plcp(T, SAo)
            rco, Po = o.plcp(T, SAo)
            d = first_diff(P, Po)
            res.append(("plcp", rc == 0 and d is None, "rc=%d %s" % (rc, d)))
            rc, L = cu.lcp(Po, SAo)
            rco, Lo = o.lcp(Po, SAo)
            d = first_diff(L, Lo)
            res.append(("lcp", rc == 0 and d is None, "rc=%d %s" % (rc, d)))
    except Exception:
        traceback.print_exc()
        res.append(("exception", False, ""))
    ok = all(r[1] for r in res)
    if not ok or verbose:
        print("%-22s n=%-9d %s" % (name, len(T), " ".join("%s:%s" % (r[0], "ok" if r[1] else "FAIL[" + r[2] + "]") for r in res)), flush=True)
    return ok


def main():
    print("devices:", libsais_b200.device_count(), flush=True)
    cu, o = _libs.cuda(), _libs.oracle()
    nbad = 0
    allc = list(cases.small_cases().items()) + list(cases.medium_cases().items())
    allc += [("dna_1m", gen.dna(1, 1 << 20)), ("bytes_4m", gen.rand_bytes(2, 1 << 22)), ("rep_19k_x100", gen.repetitive_dna(19000, 100))]
    for name, T in allc:
        if not run_case(name, T, cu, o, verbose=len(T) > 30000):
            nbad += 1
    for name, (T, k) in cases.int_cases().items():
        a, b = cu.sa_int(T, k), o.sa_int(T, k)
        d = first_diff(a[1], b[1])
        if a[0] != 0 or d:
            nbad += 1
            print("int:%s rc=%d %s" % (name, a[0], d), flush=True)
    print("cases failed:", nbad, "of", len(allc) + len(cases.int_cases()))
    import os
    for lazy in ("0", "1"):
        os.environ["LIBSAIS_CUDA_LAZY_ISA"] = lazy
        nb = 0
        for name, T in allc:
            if len(T) > 1 and not run_case(name + "/lazy" + lazy, T, cu, o):
                nb += 1
        print("forced lazy=%s: cases failed: %d" % (lazy, nb), flush=True)
    del os.environ["LIBSAIS_CUDA_LAZY_ISA"]
    # stats of one call
    ctx = libsais_b200.Context(0)
    T = gen.rand_bytes(2, 1 << 24)
    SA = np.empty(len(T), dtype=np.int32)
    ctx.set_profiling(True)
    for _ in range(2):
        t0 = time.time()
        rc = ctx.libsais(T, SA)
        print("libsais 16 MiB bytes rc", rc, "wall %.1f ms" % ((time.time() - t0) * 1e3), flush=True)
    st = ctx.stats()
    print("device_ms", st["device_ms"], "launches", st["total_launches"])
    for k, v in st["kernels"].items():
        print("  %-12s x%-4d %8.3f ms  %8.1f GB/s algo" % (k, v["launches"], v["ms"], v["bytes"] / max(v["ms"], 1e-9) / 1e6))
    for r in st["rounds"]:
        print("  round", r)


if __name__ == "__main__":
    main()
