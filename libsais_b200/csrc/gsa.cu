// gsa.cu -- generalized suffix arrays (reference libsais_gsa*, libsais_plcp_gsa*:
// src/libsais.c:7033-7048, :6886-6889, :3022-3053, :5472-5492, :8215-8238, :8381-8397).
//
// In a GSA every 0 byte of T (T[n-1] must be 0) is a distinct terminator, ordered by position,
// and comparisons never run past one.  That is exactly the plain suffix array of the integer
// text  T'[p] = (T[p] == 0) ? ordinal of that separator : m + T[p]   (m = number of separators),
// so the GSA reuses the integer-alphabet SA core, and PLCP-GSA (matches stop at separators) is the
// integer PLCP of T'.  This file builds T' on the device: count, scan, apply.
#include "core.h"

namespace lsc {

static const int kSepThreads = 256;
static const int kSepPerThread = 16;
static const int kSepTile = kSepThreads * kSepPerThread;

template <typename SymT>
__global__ void __launch_bounds__(kSepThreads)
sep_count_kernel(const SymT *__restrict__ T, u64 n, u32 *__restrict__ tile_counts, u64 *__restrict__ invalid)
{
    __shared__ u32 s_w[kSepThreads / 32];
    const u64 base = (u64)blockIdx.x * kSepTile + (u64)threadIdx.x * kSepPerThread;
    u32 c = 0;
    bool bad = false;
    bool prev0 = base == 0 ? true : (base < n && T[base - 1] == 0);     // T[0] == 0 is an empty first member
#pragma unroll
    for (int i = 0; i < kSepPerThread; ++i) {
        bool z = base + i < n && T[base + i] == 0;
        c += z ? 1u : 0u;
        bad = bad || (z && prev0);                                      // empty member: the reference returns -1
        prev0 = z;
    }
    if (bad) *invalid = 1;
    for (int off = 16; off; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int w = 0; w < kSepThreads / 32; ++w) t += s_w[w];
        tile_counts[blockIdx.x] = t;
    }
}

// one CTA: exclusive scan in place, total -> *total
__global__ void __launch_bounds__(1024)
sep_scan_kernel(u32 *__restrict__ v, u64 count, u64 *__restrict__ total)
{
    __shared__ u64 s_tot[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const u64 per = ((count + 31) / 32 + 31) / 32 * 32;
    const u64 lo = (u64)warp * per, hi = lo + per < count ? lo + per : count;
    u64 r = 0;
    for (u64 i = lo + lane; i < hi; i += 32) r += v[i];
    for (int off = 16; off; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    if (lane == 0) s_tot[warp] = r;
    __syncthreads();
    if (warp == 0) {
        u64 x = s_tot[lane], inc = x;
        for (int off = 1; off < 32; off <<= 1) { u64 o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
        s_tot[lane] = inc - x;
        if (lane == 31) *total = inc;
    }
    __syncthreads();
    u64 carry = s_tot[warp];
    for (u64 b = lo; b < hi; b += 32) {
        const u64 i = b + lane;
        u32 x = i < hi ? v[i] : 0;
        u32 inc = x;
        for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
        if (i < hi) v[i] = (u32)(carry + inc - x);
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

template <typename SymT>
__global__ void __launch_bounds__(kSepThreads)
sep_apply_kernel(const SymT *__restrict__ T, u64 n, const u32 *__restrict__ tile_excl, const u64 *__restrict__ total,
                 u32 *__restrict__ Tint)
{
    __shared__ u32 s_w[kSepThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 base = (u64)blockIdx.x * kSepTile + (u64)threadIdx.x * kSepPerThread;
    SymT b[kSepPerThread];
    u32 c = 0;
#pragma unroll
    for (int i = 0; i < kSepPerThread; ++i) { b[i] = base + i < n ? T[base + i] : (SymT)1; c += b[i] == 0 ? 1u : 0u; }
    u32 inc = c;
    for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    u32 before = tile_excl[blockIdx.x] + inc - c;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    const u32 m = (u32)*total;
#pragma unroll
    for (int i = 0; i < kSepPerThread; ++i)
        if (base + i < n) Tint[base + i] = b[i] == 0 ? before++ : m + (u32)b[i];
}

size_t gsa_workspace_bytes(u64 n) { return (size_t)n * 4 + ceil_div(n, kSepTile) * 4 + 1024; }

// T' (u32[n] + 64 bytes of zero padding for the PLCP compare) in the ctx arena; returns nullptr on failure.
template <typename SymT>
static u32 *build_gsa_text_t(Ctx &c, const SymT *d_T, u64 n)
{
    const u64 tiles = ceil_div(n, kSepTile);
    u32 *Tint = (u32 *)c.alloc((size_t)n * 4 + 64);
    u32 *counts = c.alloc_n<u32>(tiles);
    if (!Tint || !counts) return nullptr;
    u64 *total = c.d_scalars + S_GSA_TOTAL;
    c.check(cudaMemsetAsync((char *)Tint + (size_t)n * 4, 0, 64, c.stream));
    c.check(cudaMemsetAsync(c.d_scalars + S_GSA_INVALID, 0, sizeof(u64), c.stream));
    LSC_LAUNCH(c, KC_CONVERT, (double)n * sizeof(SymT), sep_count_kernel<SymT>, (u32)tiles, kSepThreads, 0, d_T, n, counts, c.d_scalars + S_GSA_INVALID);
    LSC_LAUNCH(c, KC_CONVERT, (double)tiles * 8, sep_scan_kernel, 1, 1024, 0, counts, tiles, total);
    LSC_LAUNCH(c, KC_CONVERT, (double)n * (4 + sizeof(SymT)), sep_apply_kernel<SymT>, (u32)tiles, kSepThreads, 0, d_T, n, counts, total, Tint);
    return c.failed() ? nullptr : Tint;
}

u32 *build_gsa_text(Ctx &c, const u8 *d_T, u64 n) { return build_gsa_text_t<u8>(c, d_T, n); }
u32 *build_gsa_text16(Ctx &c, const uint16_t *d_T, u64 n) { return build_gsa_text_t<uint16_t>(c, d_T, n); }

}  // namespace lsc
