// local_sort.cuh -- rounds >= 1 when every unresolved group is small: sort the groups where they lie.
//
// After a doubling round the active suffixes sit in slot order, group after group.  The next
// round only has to order each group by the rank of the suffix h positions further -- nothing
// moves between groups.  When no group is larger than 2048 elements the global onesweep
// (7 digit passes over (u64,u32) pairs in HBM at n ~ 2^31, each with its look-back) is replaced
// by ONE kernel: a CTA takes the groups that START in a window [t*C, (t+1)*C) of the active list
// (whole groups, at most the tile capacity: find_tile), gathers their keys -- the ISA look-up of
// round_keys_kernel is fused in -- orders them in shared memory and writes (key, position) back
// in place.  One read and one write of the active list per round.
//   local_count_kernel  no group > 128: every element counts the members of its group that
//                       precede it (s compares for a group of s, no barriers), small tiles,
//                       5-7 CTAs per SM
//   local_sort_kernel   groups up to 2048: stable 8-bit LSD passes inside shared memory with the
//                       ballot multisplit of the global pass, as many as the tile needs
//                       (bits(rank) + bits(groups in the tile)); tiles that happen to hold only
//                       small groups take the counting path
// big_group_kernel classifies the round (largest group vs 128 / 512 / 1024 / 2048); build_sa
// (sa_core.cu) picks the kernel and the window, or falls back to the global sort.
#pragma once
#include "radix_sort.cuh"
#include "lazy_rank.cuh"

namespace lsc {

static const int kLocalCap = 4096;               // elements a tile can hold
static const int kLocalThreads = 512;
static const int kLocalIPT = kLocalCap / kLocalThreads;
static const int kLocalWarps = kLocalThreads / 32;
// A tile is the groups that start in a window of C elements; it holds at most C + (largest group) - 1
// elements, so the window is chosen from the largest group of the round: C = kLocalCap - limit for the
// smallest limit in {512, 1024, 2048} that no group exceeds (fuller tiles for smaller groups).
static const u32 kLocalLimits[3] = {512, 1024, 2048};
static const u32 kLocalCountLimit = 128;         // tiles whose groups are all smaller are ranked by counting, not by radix passes
// Rounds in which NO group exceeds kLocalCountLimit use a lighter kernel (count path only: small tiles, 32 KB
// of shared memory, 7 CTAs per SM, so the latency of the ISA gather hides behind other tiles' counting).
static const int kCountCap = 2048;
static const int kCountThreads = 256;
static const int kCountIPT = kCountCap / kCountThreads;
static const u32 kCountWindow = kCountCap - kLocalCountLimit;

struct LocalSmem {
    u64 keys[2][kLocalCap];
    u32 vals[2][kLocalCap];
    u32 whist[kLocalWarps * kRadixSize];
    u32 scan_tmp[32];
    u32 bounds[2];
};

// flag bits: some group of the active list has more than 128 (bit 0) / kLocalLimits[i] (bit i + 1) elements
// (elements that far apart share the group id)
__global__ void __launch_bounds__(256)
big_group_kernel(const u32 *__restrict__ a_grp, const u64 *__restrict__ d_count, u64 *__restrict__ flag)
{
    const u64 N = *d_count;
    u32 bits = 0;
    for (u64 j = (u64)blockIdx.x * 256 + threadIdx.x; j + kLocalCountLimit < N; j += (u64)gridDim.x * 256) {
        const u32 g = a_grp[j];
        if (a_grp[j + kLocalCountLimit] == g) {
            bits |= 1u;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (j + (512u << i) < N && a_grp[j + (512u << i)] == g) bits |= 2u << i;      // kLocalLimits[i]
        }
    }
    if (bits) atomicOr((unsigned long long *)flag, (unsigned long long)bits);
}

// The tile of CTA t: the groups whose head lies in the window [t0, t0 + C) = from the first head at or
// after t0 to the first head at or after t0 + C.  bounds[] (shared) receives both, relative to t0;
// false (and *err set) when no head turns up within `cap` elements -- a group larger than promised.
template <int THREADS>
__device__ __forceinline__ bool find_tile(const u32 *__restrict__ a_grp, u64 N, u64 t0, u32 C, u32 cap, u32 *bounds, u32 *err)
{
    const int tid = threadIdx.x;
    if (tid < 2) bounds[tid] = 0xFFFFFFFFu;
    __syncthreads();
    bool ok = true;
    for (int which = 0; which < 2; ++which) {
        const u64 target = t0 + (u64)which * C;
        if (target >= N) { if (tid == 0) bounds[which] = (u32)(N - t0); continue; }
        bool found = false;
        for (u32 off = 0; !found; off += THREADS) {
            const u64 j = target + off + tid;
            const bool hd = j < N ? (j == 0 || a_grp[j] != a_grp[j - 1]) : j == N;
            if (hd) atomicMin(&bounds[which], (u32)(j - t0));
            found = __syncthreads_or(hd) != 0;
            if (!found && off > cap) { if (tid == 0) *err = 2; ok = false; found = true; }
        }
    }
    __syncthreads();
    return ok && bounds[1] >= bounds[0] && bounds[1] - bounds[0] <= cap;
}

// LAZY: the lazy ISA -- an invalid entry is a round-0 singleton whose rank is recomputed (lazy_rank.cuh)
template <bool LAZY>
__global__ void __launch_bounds__(kCountThreads, 5)
local_count_kernel(const u32 *__restrict__ a_pos, const u32 *__restrict__ a_grp, const u32 *__restrict__ ISA,
                   u64 N, u64 n, u64 h, int rank_bits, u64 *__restrict__ keys_out, u32 *__restrict__ pos_out, u32 *err, const LazyArgs la)
{
    __shared__ u64 comp[kCountCap];        // ((group in tile << rank_bits | rank) << 11) | index: unique, ties by index
    __shared__ u32 vals[kCountCap];
    __shared__ u32 gstart[kCountCap];
    __shared__ u32 bounds[2];
    const int tid = threadIdx.x;
    const u64 t0 = (u64)blockIdx.x * kCountWindow;
    if (!find_tile<kCountThreads>(a_grp, N, t0, kCountWindow, (u32)kCountCap, bounds, err)) {
        if (tid == 0) *err = 2;
        return;
    }
    const u32 cnt = bounds[1] - bounds[0];
    if (cnt == 0) return;
    const u64 s = t0 + bounds[0];
    const u32 g0 = a_grp[s], gmax = a_grp[s + cnt - 1] - g0;
    {
        u32 p[kCountIPT], g[kCountIPT], r[kCountIPT];
#pragma unroll
        for (int i = 0; i < kCountIPT; ++i) {
            const u32 idx = i * kCountThreads + tid;
            p[i] = idx < cnt ? a_pos[s + idx] : 0;
            g[i] = idx < cnt ? a_grp[s + idx] - g0 : 0;
        }
#pragma unroll
        for (int i = 0; i < kCountIPT; ++i) {
            const u32 idx = i * kCountThreads + tid;
            const u64 q = (u64)p[i] + h;
            r[i] = (idx < cnt && q < n) ? ISA[q] + 1 : 0;
            if (LAZY && idx < cnt && q < n && r[i] == 0) r[i] = lazy_rank(la, q) + 1;          // kIsaInvalid + 1 == 0
        }
#pragma unroll
        for (int i = 0; i < kCountIPT; ++i) {
            const u32 idx = i * kCountThreads + tid;
            if (idx < cnt) {
                comp[idx] = ((((u64)g[i] << rank_bits) | (u64)r[i]) << 11) | idx;
                vals[idx] = p[i];
                if (idx == 0 || g[i] != a_grp[s + idx - 1] - g0) gstart[g[i]] = idx;
            }
        }
    }
    __syncthreads();
    const u64 rmask = ((u64)1 << rank_bits) - 1;
    for (u32 idx = tid; idx < cnt; idx += kCountThreads) {
        const u64 mine = comp[idx], key = mine >> 11;
        const u32 g = (u32)(key >> rank_bits);
        const u32 a = gstart[g], b = g == gmax ? cnt : gstart[g + 1];
        u32 r = 0;
#pragma unroll 4
        for (u32 j = a; j < b; ++j) r += comp[j] < mine;
        const u64 o = s + a + r;
        keys_out[o] = (((u64)g + (u64)g0) << rank_bits) | (key & rmask);
        pos_out[o] = vals[idx];
    }
}

__global__ void __launch_bounds__(kLocalThreads, 2)
local_sort_kernel(const u32 *__restrict__ a_pos, const u32 *__restrict__ a_grp, const u32 *__restrict__ ISA,
                  u64 N, u64 n, u64 h, int rank_bits, u32 C, u64 *__restrict__ keys_out, u32 *__restrict__ pos_out, u32 *err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LocalSmem &sm = *reinterpret_cast<LocalSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 t0 = (u64)blockIdx.x * C;

    if (!find_tile<kLocalThreads>(a_grp, N, t0, C, (u32)kLocalCap, sm.bounds, err)) {
        if (tid == 0) *err = 2;                 // a group larger than promised slipped through
        return;
    }
    const u32 b0 = sm.bounds[0], cnt = sm.bounds[1] - b0;
    if (cnt == 0) return;
    const u64 s = t0 + b0;

    // ---- gather: key = (group - first group of the tile) << rank_bits | rank(p + h) + 1
    const u32 g0 = a_grp[s], gmax = a_grp[s + cnt - 1] - g0;
    {
        u32 p[kLocalIPT], g[kLocalIPT], r[kLocalIPT];
#pragma unroll
        for (int i = 0; i < kLocalIPT; ++i) {
            const u32 idx = i * kLocalThreads + tid;
            p[i] = idx < cnt ? a_pos[s + idx] : 0;
            g[i] = idx < cnt ? a_grp[s + idx] - g0 : 0;
        }
#pragma unroll
        for (int i = 0; i < kLocalIPT; ++i) {
            const u32 idx = i * kLocalThreads + tid;
            const u64 q = (u64)p[i] + h;
            r[i] = (idx < cnt && q < n) ? ISA[q] + 1 : 0;
        }
#pragma unroll
        for (int i = 0; i < kLocalIPT; ++i) {
            const u32 idx = i * kLocalThreads + tid;
            if (idx < cnt) {
                const u64 key = ((u64)g[i] << rank_bits) | (u64)r[i];
                sm.keys[0][idx] = key;
                sm.keys[1][idx] = (key << 12) | idx;          // unique: ties broken by the index (counting path)
                sm.vals[0][idx] = p[i];
            }
        }
    }
    const int key_bits = rank_bits + (gmax ? 32 - __clz(gmax) : 0);
    const int passes = (key_bits + kRadixBits - 1) / kRadixBits;
    const u64 rmask = ((u64)1 << rank_bits) - 1;
    __syncthreads();

    // ---- a tile of small groups: every element counts the members of its group that precede it
    // (a group of s elements costs s compares per element, broadcast reads of shared memory, no barriers;
    // cheaper than the digit passes up to s ~ 150) and writes itself straight to its final place
    {
        u32 *gstart = sm.whist;                       // first index of every group of the tile
        for (u32 idx = tid; idx < cnt; idx += kLocalThreads) {
            const u32 g = (u32)(sm.keys[0][idx] >> rank_bits);
            if (idx == 0 || (u32)(sm.keys[0][idx - 1] >> rank_bits) != g) gstart[g] = idx;
        }
        __syncthreads();
        bool large = false;
        for (u32 idx = tid; idx < cnt; idx += kLocalThreads)
            large |= idx - gstart[(u32)(sm.keys[0][idx] >> rank_bits)] >= kLocalCountLimit;
        if (!__syncthreads_or(large)) {
            for (u32 idx = tid; idx < cnt; idx += kLocalThreads) {
                const u64 key = sm.keys[0][idx], mine = sm.keys[1][idx];
                const u32 g = (u32)(key >> rank_bits);
                const u32 a = gstart[g], b = g == gmax ? cnt : gstart[g + 1];
                u32 r = 0;
#pragma unroll 4
                for (u32 j = a; j < b; ++j) r += sm.keys[1][j] < mine;
                const u64 o = s + a + r;
                keys_out[o] = (((u64)g + (u64)g0) << rank_bits) | (key & rmask);
                pos_out[o] = sm.vals[0][idx];
            }
            return;
        }
    }

    // ---- stable LSD passes inside shared memory.  The cnt elements are dealt to the warps in equal
    // contiguous shares (ipw items per lane); element order = index order (warp, item, lane).
    const u32 lt = lanemask_lt();
    u32 *wh = sm.whist + warp * kRadixSize;
    const u32 ipw = (cnt + kLocalThreads - 1) / kLocalThreads;
    const u32 wbase = warp * ipw * 32 + lane;
    int cur = 0;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = pass * kRadixBits;
        const bool runs = shift + kRadixBits > rank_bits;          // digits that hold group bits come in runs
        {
            uint4 *z = reinterpret_cast<uint4 *>(sm.whist);
            for (int i = tid; i < kLocalWarps * kRadixSize / 4; i += kLocalThreads) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        const u64 *kc = sm.keys[cur];
        if (runs) {
            for (u32 i = 0; i < ipw; ++i) {
                const u32 li = wbase + i * 32;
                const bool valid = li < cnt;
                warp_hist_add(wh, valid ? (u32)(kc[li] >> shift) & 255u : 0u, valid, lane);
            }
        } else {
#pragma unroll 4
            for (u32 i = 0; i < ipw; ++i) {
                const u32 li = wbase + i * 32;
                if (li < cnt) atomicAdd(&wh[(u32)(kc[li] >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        // digit d (thread d): offset of every (warp, digit) run inside the tile
        u32 dcnt = 0, tileoff = 0;
        if (tid < kRadixSize) {
#pragma unroll
            for (int w = 0; w < kLocalWarps; ++w) dcnt += sm.whist[w * kRadixSize + tid];
            u32 x = dcnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
            if (lane == 31) sm.scan_tmp[warp] = x;
            tileoff = x - dcnt;
        }
        __syncthreads();
        if (tid < kRadixSize) {
#pragma unroll
            for (int w = 0; w < kRadixSize / 32; ++w) if (w < warp) tileoff += sm.scan_tmp[w];
            u32 run = tileoff;
#pragma unroll
            for (int w = 0; w < kLocalWarps; ++w) { u32 c = sm.whist[w * kRadixSize + tid]; sm.whist[w * kRadixSize + tid] = run; run += c; }
        }
        __syncthreads();
#pragma unroll 2
        for (u32 i = 0; i < ipw; ++i) {
            const u32 li = wbase + i * 32;
            const bool valid = li < cnt;
            const u64 key = valid ? kc[li] : 0;
            const u32 d = valid ? (u32)(key >> shift) & 255u : (u32)kRadixSize;
            const u32 peers = digit_peers(d, valid, lane);
            const int leader = __ffs(peers) - 1;
            u32 basepos = 0;
            if (lane == leader && valid) basepos = atomicAdd(&wh[d], (u32)__popc(peers));
            basepos = __shfl_sync(0xffffffffu, basepos, leader);
            if (valid) {
                const u32 pos = basepos + __popc(peers & lt);
                sm.keys[cur ^ 1][pos] = key;
                sm.vals[cur ^ 1][pos] = sm.vals[cur][li];
            }
        }
        __syncthreads();
        cur ^= 1;
    }

    // ---- write back in place, group ids global again
    for (u32 idx = tid; idx < cnt; idx += kLocalThreads) {
        const u64 k = sm.keys[cur][idx];
        keys_out[s + idx] = (((k >> rank_bits) + (u64)g0) << rank_bits) | (k & rmask);
        pos_out[s + idx] = sm.vals[cur][idx];
    }
}

}  // namespace lsc
