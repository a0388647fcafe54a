"""Run the BASELINE.json configs that fit one GPU through the device-pointer API with per-kernel
profiling, verify the results with size-independent properties on the GPU (torch), and print one
JSON object per config.  usage: python tools/config_bench.py c1 c2 c3s c3 c4b ..."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import libsais_b200
from libsais_b200 import gen


def verify_sa(dT, dSA, n, chunk=1 << 27):
    """Linear-time SA check on the GPU: permutation + Burkhardt-Kaerkkaeinen neighbour order."""
    seen = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for lo in range(0, n, chunk):
        seen[dSA[lo:lo + chunk].long()] = 1
    if not bool(seen.all()):
        return "not a permutation"
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
        ok = (ta < tb) | ((ta == tb) & (ra < rb))
        if not bool(ok.all()):
            return "order violated near slot %d" % (lo + int((~ok).nonzero()[0]))
    return "ok"


def verify_plcp_sample(T, SA_host_fn, dSA, dP, n, samples=20000, seed=1):
    """PLCP spot check against a direct byte comparison on the CPU."""
    rng = np.random.default_rng(seed)
    slots = rng.integers(1, n, samples)
    sl = torch.from_numpy(slots).cuda()
    cur = dSA[sl].cpu().numpy().astype(np.int64); prev = dSA[sl - 1].cpu().numpy().astype(np.int64)
    got = dP[torch.from_numpy(cur).cuda()].cpu().numpy()
    for c, p, g in zip(cur, prev, got):
        l = 0
        m = n - max(c, p)
        while l < m:
            step = min(4096, m - l)
            x = T[c + l:c + l + step] != T[p + l:p + l + step]
            nz = np.flatnonzero(x)
            if len(nz):
                l += int(nz[0]); break
            l += step
        if l != g:
            return "PLCP[%d] = %d, expected %d" % (c, g, l)
    return "ok"


def kern_table(st):
    return {k: {"x": v["launches"], "ms": round(v["ms"], 3), "algo_gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in st["kernels"].items()}


def run(name):
    ctx = libsais_b200.Context(0)
    ctx.set_profiling(True)
    out = {"config": name}
    t0 = time.time()
    if name == "c1":
        T = gen.dna(1, 1 << 20); what = ("sa", "plcp", "lcp")
    elif name == "c2":
        T = gen.rand_bytes(2, 1 << 28); what = ("bwt",)
    elif name == "c2sa":
        T = gen.rand_bytes(2, 1 << 28); what = ("sa",)
    elif name == "c3s":
        T = gen.repetitive_dna(1_900_000, 100); what = ("sa", "plcp", "lcp", "bwt")
    elif name == "c3":
        T = gen.repetitive_dna(19_000_000, 100); what = ("sa", "plcp", "lcp")
    elif name == "c3d":                                      # config 3 at full size, generated on the device (SA only: the rounds)
        T = None; what = ("sa",)
    elif name == "c5lite":
        return run_c5lite()
    elif name == "c4b":
        T = gen.dna(1000, 1 << 27); what = ("bwt",)
    elif name == "text":
        rng = np.random.default_rng(5)
        words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(5000)]
        idx = rng.zipf(1.3, 12_000_000) % 5000
        T = np.frombuffer(b" ".join(words[i] for i in idx), dtype=np.uint8).copy(); what = ("sa", "plcp", "bwt")
    else:
        raise SystemExit("unknown config " + name)
    if T is None:
        dT = gen.repetitive_dna_torch(19_000_000, 100)
        torch.cuda.synchronize(); torch.cuda.empty_cache()
    else:
        dT = torch.from_numpy(T).cuda()
    n = dT.numel()
    out["n"] = n; out["gen_s"] = round(time.time() - t0, 1)
    dSA = None
    for w in what:
        if w == "sa":
            dSA = torch.empty(n, dtype=torch.int32, device="cuda")
            for _ in range(2):
                rc = ctx.sa_dev(dT.data_ptr(), dSA.data_ptr(), n)
            st = ctx.stats()
            out["sa"] = {"rc": rc, "ms": round(st["device_ms"], 3), "mbs": round(n / 1e6 / (st["device_ms"] / 1e3), 1), "launches": st["total_launches"],
                         "kernels": kern_table(st), "rounds": [(r["h"], r["n_active"], r["passes"]) for r in st["rounds"]]}
            out["sa"]["verify"] = verify_sa(dT, dSA, n)
        elif w == "bwt":
            dU = torch.empty(n, dtype=torch.uint8, device="cuda")
            for _ in range(2):
                rc = ctx.bwt_dev(dT.data_ptr(), dU.data_ptr(), n)
            st = ctx.stats()
            out["bwt"] = {"primary": rc, "ms": round(st["device_ms"], 3), "mbs": round(n / 1e6 / (st["device_ms"] / 1e3), 1), "launches": st["total_launches"],
                          "kernels": kern_table(st), "rounds": [(r["h"], r["n_active"], r["passes"]) for r in st["rounds"]]}
            dB = torch.empty(n, dtype=torch.uint8, device="cuda")
            for _ in range(2):
                rcu = ctx.unbwt_dev(dU.data_ptr(), dB.data_ptr(), n, rc)
            st = ctx.stats()
            out["unbwt"] = {"rc": rcu, "ms": round(st["device_ms"], 3), "mbs": round(n / 1e6 / (st["device_ms"] / 1e3), 1), "kernels": kern_table(st),
                            "roundtrip": bool(torch.equal(dB, dT))}
            del dU, dB
        elif w == "plcp":
            dP = torch.empty(n, dtype=torch.int32, device="cuda")
            for _ in range(2):
                rc = ctx.plcp_dev(dT.data_ptr(), dSA.data_ptr(), dP.data_ptr(), n)
            st = ctx.stats()
            out["plcp"] = {"rc": rc, "ms": round(st["device_ms"], 3), "mbs": round(n / 1e6 / (st["device_ms"] / 1e3), 1), "kernels": kern_table(st),
                           "max": int(dP.max()), "mean": round(float(dP.double().mean()), 2)}
            out["plcp"]["verify_sample"] = verify_plcp_sample(T, None, dSA, dP, n)
        elif w == "lcp":
            dL = torch.empty(n, dtype=torch.int32, device="cuda")
            for _ in range(2):
                rc = ctx.lcp_dev(dP.data_ptr(), dSA.data_ptr(), dL.data_ptr(), n)
            st = ctx.stats()
            ok = True
            for lo in range(0, n, 1 << 27):
                ok = ok and bool(torch.equal(dL[lo:lo + (1 << 27)], dP[dSA[lo:lo + (1 << 27)].long()]))
            out["lcp"] = {"rc": rc, "ms": round(st["device_ms"], 3), "mbs": round(n / 1e6 / (st["device_ms"] / 1e3), 1), "kernels": kern_table(st), "verify": ok}
            del dL
    print(json.dumps(out), flush=True)
    ctx.close()


def run_c5lite():
    """libsais64 through the HOST API on n = 2^31 + 1024 iid ACGT (seed 5): the n > INT32_MAX entry
    point on one GPU (u32 device indexes, int64 at the boundary).  Verified with the linear-time checker."""
    import ctypes as C
    n = (1 << 31) + 1024
    t0 = time.time()
    T = gen.dna(5, n)
    out = {"config": "c5lite", "n": n, "gen_s": round(time.time() - t0, 1)}
    lib = libsais_b200.load_library()
    lib.libsais64.restype = C.c_int64
    SA = np.empty(n, dtype=np.int64)
    freq = np.zeros(256, dtype=np.int64)
    t0 = time.time()
    rc = lib.libsais64(T.ctypes.data_as(C.c_void_p), SA.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int64(0), freq.ctypes.data_as(C.c_void_p))
    out["libsais64"] = {"rc": int(rc), "wall_s": round(time.time() - t0, 2), "freq_ok": bool(freq.sum() == n), "sa0": int(SA[0]), "sa0_expected_tail": bool(SA[0] >= n - 64)}
    st = libsais_b200.Stats()
    lib.libsais_cuda_get_stats(None, C.byref(st))
    out["libsais64"]["device_ms"] = round(st.device_ms, 1)
    out["libsais64"]["mbs_device"] = round(n / 1e6 / (st.device_ms / 1e3), 1)
    # free the library workspace before the torch-side check: drop the thread's default context by exiting later;
    # verify in chunks with int32 SA on the device
    dT = torch.from_numpy(T).cuda()
    dSA = torch.empty(n, dtype=torch.int32, device="cuda")
    ch = 1 << 27
    for lo in range(0, n, ch):
        dSA[lo:lo + ch] = torch.from_numpy(SA[lo:lo + ch]).cuda().to(torch.int32)
    # positions >= 2^31 wrap negative in int32: reinterpret as uint32 via long() & mask inside the checker
    dSA_l = None
    out["libsais64"]["verify"] = verify_sa_u32(dT, dSA, n)
    print(json.dumps(out), flush=True)


def verify_sa_u32(dT, dSA32, n, chunk=1 << 26):
    """verify_sa for SA stored as the low 32 bits (positions may exceed 2^31)."""
    M = 0xFFFFFFFF
    seen = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for lo in range(0, n, chunk):
        seen[dSA32[lo:lo + chunk].long() & M] = 1
    if not bool(seen.all()):
        return "not a permutation"
    del seen
    ISA = torch.empty(n, dtype=torch.int32, device="cuda")
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        ISA[dSA32[lo:hi].long() & M] = torch.arange(lo, hi, dtype=torch.int64, device="cuda").to(torch.int32)
    for lo in range(1, n, chunk):
        hi = min(n, lo + chunk)
        a = dSA32[lo - 1:hi - 1].long() & M; b = dSA32[lo:hi].long() & M
        ta, tb = dT[a], dT[b]
        ra = torch.where(a + 1 < n, ISA[torch.clamp(a + 1, max=n - 1)].long() & M, torch.full_like(a[:1], -1))
        rb = torch.where(b + 1 < n, ISA[torch.clamp(b + 1, max=n - 1)].long() & M, torch.full_like(a[:1], -1))
        ok = (ta < tb) | ((ta == tb) & (ra < rb))
        if not bool(ok.all()):
            return "order violated near slot %d" % (lo + int((~ok).nonzero()[0]))
    return "ok"


if __name__ == "__main__":
    for name in sys.argv[1:]:
        run(name)
        torch.cuda.empty_cache()
