// local_sort.cuh -- rounds >= 1 when every unresolved group is small: sort the groups where they lie.
//
// After a doubling round the active suffixes sit in slot order, group after group.  The next
// round only has to order each group by the rank of the suffix h positions further -- nothing
// moves between groups.  When no group is larger than 2048 elements the global onesweep
// (7 digit passes over (u64,u32) pairs in HBM at n ~ 2^31, each with its look-back) is replaced
// by one kernel: a CTA takes the groups that START in a window [t*C, (t+1)*C) of the active list --
// at most kLocalCap elements, see kLocalLimits -- gathers their keys (the ISA look-up of round_keys_kernel is fused in), sorts the tile in shared
// memory by (group id relative to the tile, rank) with the same stable ballot multisplit as the
// global pass, and writes (key, position) back in place.  The number of digit passes adapts to the
// tile: bits(rank) + bits(groups in the tile).  One read and one write of the active list per round.
#pragma once
#include "radix_sort.cuh"

namespace lsc {

static const int kLocalCap = 4096;               // elements a tile can hold
static const int kLocalThreads = 512;
static const int kLocalIPT = kLocalCap / kLocalThreads;
static const int kLocalWarps = kLocalThreads / 32;
// A tile is the groups that start in a window of C elements; it holds at most C + (largest group) - 1
// elements, so the window is chosen from the largest group of the round: C = kLocalCap - limit for the
// smallest limit in {512, 1024, 2048} that no group exceeds (fuller tiles for smaller groups).
static const u32 kLocalLimits[3] = {512, 1024, 2048};

struct LocalSmem {
    u64 keys[2][kLocalCap];
    u32 vals[2][kLocalCap];
    u32 whist[kLocalWarps * kRadixSize];
    u32 scan_tmp[32];
    u32 bounds[2];
};

// flag bit i set when some group of the active list has more than kLocalLimits[i] elements (elements that
// far apart share the group id)
__global__ void __launch_bounds__(256)
big_group_kernel(const u32 *__restrict__ a_grp, const u64 *__restrict__ d_count, u64 *__restrict__ flag)
{
    const u64 N = *d_count;
    u32 bits = 0;
    for (u64 j = (u64)blockIdx.x * 256 + threadIdx.x; j + 512 < N; j += (u64)gridDim.x * 256) {
        const u32 g = a_grp[j];
        if (a_grp[j + 512] == g) {
            bits |= 1u;
            if (j + 1024 < N && a_grp[j + 1024] == g) {
                bits |= 2u;
                if (j + 2048 < N && a_grp[j + 2048] == g) bits |= 4u;
            }
        }
    }
    if (bits) atomicOr((unsigned long long *)flag, (unsigned long long)bits);
}

__global__ void __launch_bounds__(kLocalThreads, 2)
local_sort_kernel(const u32 *__restrict__ a_pos, const u32 *__restrict__ a_grp, const u32 *__restrict__ ISA,
                  u64 N, u64 n, u64 h, int rank_bits, u32 C, u64 *__restrict__ keys_out, u32 *__restrict__ pos_out, u32 *err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LocalSmem &sm = *reinterpret_cast<LocalSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 t0 = (u64)blockIdx.x * C;

    // ---- tile = the groups whose head lies in [t0, t0 + C): first head at or after t0, first head at or after t0 + C
    if (tid < 2) sm.bounds[tid] = 0xFFFFFFFFu;
    __syncthreads();
    for (int which = 0; which < 2; ++which) {
        const u64 target = t0 + (u64)which * C;
        if (target >= N) { if (tid == 0) sm.bounds[which] = (u32)(N - t0); continue; }
        bool found = false;
        for (u32 off = 0; !found; off += kLocalThreads) {
            const u64 j = target + off + tid;
            const bool hd = j < N ? (j == 0 || a_grp[j] != a_grp[j - 1]) : j == N;
            if (hd) atomicMin(&sm.bounds[which], (u32)(j - t0));
            found = __syncthreads_or(hd) != 0;
            if (!found && off > (u32)kLocalCap) { if (tid == 0) { *err = 2; sm.bounds[which] = 0xFFFFFFFEu; } found = true; }
        }
    }
    __syncthreads();
    const u32 b0 = sm.bounds[0], b1 = sm.bounds[1];
    if (b0 >= 0xFFFFFFFEu || b1 >= 0xFFFFFFFEu || b1 < b0) return;
    const u32 cnt = b1 - b0;
    if (cnt == 0) return;
    if (cnt > (u32)kLocalCap) { if (tid == 0) *err = 2; return; }     // a group larger than C slipped through
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
            if (idx < cnt) { sm.keys[0][idx] = ((u64)g[i] << rank_bits) | (u64)r[i]; sm.vals[0][idx] = p[i]; }
        }
    }
    const int key_bits = rank_bits + (gmax ? 32 - __clz(gmax) : 0);
    const int passes = (key_bits + kRadixBits - 1) / kRadixBits;
    __syncthreads();

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
    const u64 rmask = ((u64)1 << rank_bits) - 1;
    for (u32 idx = tid; idx < cnt; idx += kLocalThreads) {
        const u64 k = sm.keys[cur][idx];
        keys_out[s + idx] = (((k >> rank_bits) + (u64)g0) << rank_bits) | (k & rmask);
        pos_out[s + idx] = sm.vals[cur][idx];
    }
}

}  // namespace lsc
