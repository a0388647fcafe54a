// lazy_rank.cuh -- the lazy ISA: with few unresolved suffixes after round 0 the ranks of the round-0 singletons are never
// scattered into ISA; the rare look-ups that hit one recompute its rank by binary search in the round-0 order.
#pragma once
#include "common.cuh"

namespace lsc {

static const u32 kIsaInvalid = 0xFFFFFFFFu;

// k-mer of suffix p: the K most significant bits of the 64-bit window at bit p*b
__device__ __forceinline__ u64 kmer_at(const u64 *__restrict__ words, u64 p, int b, int K)
{
    u64 bit = p * (u64)b;
    u64 q = bit >> 6; int off = (int)(bit & 63);
    u64 hi = words[q], lo = words[q + 1];
    u64 x = off ? ((hi << off) | (lo >> (64 - off))) : hi;
    return x >> (64 - K);
}

// Rank of a round-0 singleton q, recomputed from the sorted round-0 keys (s0_keys == nullptr: from the k-mers of the
// suffixes in slot order -- the fused MSD path never writes the sorted keys; later rounds only permute positions inside
// groups of equal k-mers, so the suffix array in progress serves as s0_pos).  boff16 (MSD path): the slot range of every
// 16-bit k-mer prefix -- the search starts inside the bucket of q's prefix, ~12 steps instead of ~28.
struct LazyArgs { const u64 *words; int b; int K; const u64 *s0_keys; const u32 *s0_pos; int key_shift; u64 tail_start; u64 n; const u32 *boff16; };

__device__ __forceinline__ u32 lazy_rank(const LazyArgs &la, u64 q)
{
    const u64 kq = kmer_at(la.words, q, la.b, la.K);
    u64 lo = 0, hi = la.n;                               // lower bound of kq among the sorted k-mers
    if (la.boff16 != nullptr) { const u32 pre = (u32)(kq >> (la.K - 16)); lo = la.boff16[pre]; hi = la.boff16[pre + 1]; }
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        const u64 km = la.s0_keys ? la.s0_keys[mid] >> la.key_shift : kmer_at(la.words, (u64)la.s0_pos[mid], la.b, la.K);
        if (km < kq) lo = mid + 1; else hi = mid;
    }
    u64 s = lo;
    if (q >= la.tail_start) { while (s + 1 < la.n && (u64)la.s0_pos[s] != q) ++s; }       // its own slot
    else { while (s + 1 < la.n && (u64)la.s0_pos[s] >= la.tail_start) ++s; }              // first full-length suffix
    return (u32)s;
}

}  // namespace lsc
