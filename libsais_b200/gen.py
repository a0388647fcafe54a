"""Counter-based synthetic inputs, identical on every host: splitmix64 of
(seed << 40) + block index.  Used by bench.py and the tests; pure numpy, no GPU."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def sm64(x):
    """splitmix64 finaliser over a uint64 array (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _chunks(n, step):
    for lo in range(0, n, step):
        yield lo, min(n, lo + step)


def dna_codes(seed, n, start=0):
    """2-bit codes: dna(seed,i) = (sm64((seed<<40)+(i>>5)) >> ((i&31)*2)) & 3."""
    out = np.empty(n, dtype=np.uint8)
    for lo, hi in _chunks(n, 1 << 24):
        i = np.arange(start + lo, start + hi, dtype=np.uint64)
        w = sm64((np.uint64(seed) << np.uint64(40)) + (i >> np.uint64(5)))
        out[lo:hi] = ((w >> ((i & np.uint64(31)) * np.uint64(2))) & np.uint64(3)).astype(np.uint8)
    return out


_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def dna(seed, n):
    """iid uniform ACGT text (configs c1, c4, c5)."""
    return _ACGT[dna_codes(seed, n)]


def rand_bytes(seed, n):
    """iid uniform bytes: byte(seed,i) = (sm64((seed<<40)+(i>>3)) >> ((i&7)*8)) & 255 (config c2)."""
    out = np.empty(n, dtype=np.uint8)
    for lo, hi in _chunks(n, 1 << 24):
        i = np.arange(lo, hi, dtype=np.uint64)
        w = sm64((np.uint64(seed) << np.uint64(40)) + (i >> np.uint64(3)))
        out[lo:hi] = ((w >> ((i & np.uint64(7)) * np.uint64(8))) & np.uint64(255)).astype(np.uint8)
    return out


def repetitive_dna(base_len, copies, seed=3, rate_num=1049):
    """config c3: `copies` mutated copies of an iid ACGT base genome (star phylogeny).
    copy 0 is unmutated; copy c>0 substitutes position j when (u & 0xFFFFF) < rate_num
    (~1e-3) with u = sm64(((3000+c)<<40)+j), new code = (code+1+((u>>20)%3))&3."""
    base = dna_codes(seed, base_len)
    out = np.empty(base_len * copies, dtype=np.uint8)
    j = np.arange(base_len, dtype=np.uint64)
    for c in range(copies):
        code = base
        if c > 0:
            u = sm64((np.uint64(3000 + c) << np.uint64(40)) + j)
            hit = (u & np.uint64(0xFFFFF)) < np.uint64(rate_num)
            sub = (base + np.uint8(1) + ((u >> np.uint64(20)) % np.uint64(3)).astype(np.uint8)) & np.uint8(3)
            code = np.where(hit, sub, base)
        out[c * base_len:(c + 1) * base_len] = _ACGT[code]
    return out


def _sm64_torch(x):
    """splitmix64 finaliser over an int64 tensor (wrap-around arithmetic; logical shifts emulated with masks)."""
    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)

    def c(v):                                     # 64-bit constant as a signed python int
        return v - (1 << 64) if v >= (1 << 63) else v

    z = x + c(0x9E3779B97F4A7C15)
    z = (z ^ lsr(z, 30)) * c(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * c(0x94D049BB133111EB)
    return z ^ lsr(z, 31)


def dna_codes_torch(seed, n, device="cuda", chunk=1 << 26, start=0):
    """Same 2-bit codes as dna_codes(), on the device."""
    import torch
    out = torch.empty(n, dtype=torch.uint8, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        i = torch.arange(start + lo, start + hi, dtype=torch.int64, device=device)
        z = _sm64_torch((seed << 40) + (i >> 5))
        sh = (i & 31) * 2
        code = torch.where(sh == 0, z & 3, ((z >> 1) & ((1 << 63) - 1)) >> (sh - 1).clamp(min=0)) & 3
        out[lo:hi] = code.to(torch.uint8)
    return out


def repetitive_dna_torch(base_len, copies, seed=3, rate_num=1049, device="cuda"):
    """Same text as repetitive_dna(), generated on the device (config c3: 1.9 GB takes minutes in numpy)."""
    import torch
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    base = dna_codes_torch(seed, base_len, device=device)
    out = torch.empty(base_len * copies, dtype=torch.uint8, device=device)
    j = torch.arange(base_len, dtype=torch.int64, device=device)
    for c in range(copies):
        code = base
        if c > 0:
            u = _sm64_torch(((3000 + c) << 40) + j)
            hit = (u & 0xFFFFF) < rate_num
            um = (u >> 20) & ((1 << 44) - 1)                       # logical u >> 20
            sub = (base.long() + 1 + (um % 3)) & 3
            code = torch.where(hit, sub.to(torch.uint8), base)
        out[c * base_len:(c + 1) * base_len] = lut[code.long()]
    return out


def rand_bytes_torch(seed, n, device="cuda", chunk=1 << 26):
    """Same bytes as rand_bytes(), on the device."""
    import torch
    out = torch.empty(n, dtype=torch.uint8, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        i = torch.arange(lo, hi, dtype=torch.int64, device=device)
        z = _sm64_torch((seed << 40) + (i >> 3))
        sh = (i & 7) * 8
        v = torch.where(sh == 0, z & 255, ((z >> 1) & ((1 << 63) - 1)) >> (sh - 1).clamp(min=0)) & 255
        out[lo:hi] = v.to(torch.uint8)
    return out


def dna_torch(seed, n, device="cuda", chunk=1 << 27):
    """Same iid ACGT text as dna(), generated on the device with torch integer ops (int64 wrap-around
    arithmetic; logical shifts emulated with masks).  For texts too large to generate on the host in time."""
    import torch
    out = torch.empty(n, dtype=torch.uint8, device=device)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)

    def lsr(x, s):
        return (x >> s) & ((1 << (64 - s)) - 1)

    def c(v):                                     # 64-bit constant as a signed python int
        return v - (1 << 64) if v >= (1 << 63) else v

    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        i = torch.arange(lo, hi, dtype=torch.int64, device=device)
        x = (seed << 40) + (i >> 5)
        z = x + c(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * c(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * c(0x94D049BB133111EB)
        z = z ^ lsr(z, 31)
        code = lsr(z, 1) >> ((i & 31) * 2 - 1).clamp(min=0)
        code = torch.where((i & 31) == 0, z & 3, code & 3)
        out[lo:hi] = lut[code]
    return out


def dna_torch_range(seed, lo, hi, device="cuda"):
    """Symbols [lo, hi) of dna(seed, .) on the device."""
    import torch
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    return lut[dna_codes_torch(seed, hi - lo, device=device, start=lo).long()]
