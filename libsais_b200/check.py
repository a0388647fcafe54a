"""Size-independent result checkers that run on the GPU with torch (test/bench infrastructure, not the
product): used where the CPU oracle is too slow (full-size BASELINE configs).  A suffix array is correct
iff it is a permutation of [0, n) and every neighbouring pair is in order, which the Burkhardt-Kaerkkaeinen
test decides from one symbol and the rank of the following suffix."""
import numpy as np


def verify_sa(dT, dSA, n, chunk=1 << 27):
    """dT: uint8[n] device tensor, dSA: int32[n] device tensor (positions < 2^31). Returns "ok" or a reason."""
    import torch
    dev = dT.device
    seen = torch.zeros(n, dtype=torch.uint8, device=dev)
    for lo in range(0, n, chunk):
        seen[dSA[lo:lo + chunk].long()] = 1
    if not bool(seen.all()):
        return "not a permutation"
    del seen
    ISA = torch.empty(n, dtype=torch.int32, device=dev)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        ISA[dSA[lo:hi].long()] = torch.arange(lo, hi, dtype=torch.int32, device=dev)
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


def verify_plcp_sample(T, dSA, dP, n, samples=20000, seed=1):
    """PLCP spot check: T is the HOST text (numpy), PLCP[SA[i]] must equal lcp(suffix SA[i-1], suffix SA[i])
    by direct byte comparison, at `samples` random slots."""
    import torch
    rng = np.random.default_rng(seed)
    slots = rng.integers(1, n, samples)
    sl = torch.from_numpy(slots).to(dSA.device)
    cur = dSA[sl].cpu().numpy().astype(np.int64); prev = dSA[sl - 1].cpu().numpy().astype(np.int64)
    got = dP[torch.from_numpy(cur).to(dSA.device)].cpu().numpy()
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


def verify_lcp(dP, dSA, dL, n, chunk=1 << 27):
    """LCP[i] == PLCP[SA[i]] for every i."""
    import torch
    for lo in range(0, n, chunk):
        if not bool(torch.equal(dL[lo:lo + chunk], dP[dSA[lo:lo + chunk].long()])):
            return "LCP differs in [%d, %d)" % (lo, lo + chunk)
    return "ok"


def verify_sa_u32(dT, dSA32, n, chunk=1 << 26):
    """verify_sa for an SA stored as the low 32 bits of every entry (positions may exceed 2^31)."""
    import torch
    dev = dT.device
    M = 0xFFFFFFFF
    seen = torch.zeros(n, dtype=torch.uint8, device=dev)
    for lo in range(0, n, chunk):
        seen[dSA32[lo:lo + chunk].long() & M] = 1
    if not bool(seen.all()):
        return "not a permutation"
    del seen
    ISA = torch.empty(n, dtype=torch.int32, device=dev)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        ISA[dSA32[lo:hi].long() & M] = torch.arange(lo, hi, dtype=torch.int64, device=dev).to(torch.int32)
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
