// dist64.cu -- suffix array of ONE text across several GPUs of a node: distributed prefix doubling with a
// sample sort, 64-bit positions and ranks (texts beyond 2^32 symbols / beyond one GPU's working set;
// BASELINE config 5).  This is what libsais64() runs when the text does not fit the single-GPU
// core (reference: the native 64-bit path of src/libsais64.c:7058-7086 computes the same SA on the CPU).
//
// One process, one host thread + one context (stream, workspace) per GPU, peer access enabled.  The packed text
// is replicated; rank r owns
//   * the positions [r*B, (r+1)*B): its slice of ISA, and
//   * after the round-0 sample sort a contiguous range of KEYS = a contiguous slice of the suffix array
//     (global slots [base_r, base_r + m_r)).  Equal keys never straddle ranks, so every group of tied suffixes
//     lives on one rank and ranking is local.
// Every exchange is FUSED into the kernel that produces the data: the routing pass is the unstable partition
// pass of partition.cuh whose per-destination output bases point into the PEERS' receive buffers, so elements
// leave the SM as coalesced runs of P2P stores over NVLink while the tile is still being processed -- no send
// buffer, no separate all-to-all.  Only the counts (a G x G matrix) cross through host memory, between two host
// barriers.  Exchanges per round: requests p+h -> owners (keys to the peers, ids stay local), answers back
// (contiguous peer copies), and the new (position, rank) pairs -> owners.
#include "core.h"
#include "radix_sort.cuh"
#include "partition.cuh"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/libsais_cuda.h"

namespace lsc {

// k-mer of suffix p from the packed text (same bit stream as sa_core.cu)
__device__ __forceinline__ u64 kmer64_at(const u64 *__restrict__ words, u64 p, int b, int K)
{
    const u64 bit = p * (u64)b;
    const u64 q = bit >> 6; const int off = (int)(bit & 63);
    const u64 hi = words[q], lo = words[q + 1];
    const u64 x = off ? ((hi << off) | (lo >> (64 - off))) : hi;
    return x >> (64 - K);
}

// symbols -> b-bit codes of the owned slice (slice start is a multiple of 64 symbols: whole words)
static __global__ void __launch_bounds__(256)
d64_pack_kernel(const u8 *__restrict__ T, u64 count, int b, u64 *__restrict__ words, u64 nwords, const u8 *__restrict__ lut)
{
    const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j >= nwords) return;
    const u64 bit0 = j * 64;
    u64 w = 0;
    for (u64 s = bit0 / (u64)b;; ++s) {
        const u64 sb = s * (u64)b;
        if (sb >= bit0 + 64) break;
        const u64 code = s < count ? (u64)lut[T[s]] : 0;
        const i64 sh = (i64)(bit0 + 64) - (i64)(sb + (u64)b);
        w |= sh >= 0 ? (code << sh) : (code >> (-sh));
    }
    words[j] = w;
}

// round-0 keys of the owned positions: (k-mer << len_bits) | length field (all ones = full length), so suffixes that run
// past the end are unique and sort before every longer suffix with the same zero-padded k-mer
static __global__ void __launch_bounds__(256)
d64_keys_kernel(const u64 *__restrict__ words, u64 n, int b, int k, int K, int len_bits, u64 lo, u64 count,
                u64 *__restrict__ keys, u64 *__restrict__ pos)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    const u64 p = lo + i;
    const u64 len = p + (u64)k <= n ? (((u64)1 << len_bits) - 1) : n - p;
    keys[i] = (kmer64_at(words, p, b, K) << len_bits) | len;
    pos[i] = p;
}

static const int kD64MaxRanks = 64;

// destination of a key = number of splitters <= key; it goes into the key's top byte; per-destination counts
static __global__ void __launch_bounds__(256)
d64_dest_splitters_kernel(u64 *__restrict__ keys, u64 count, const u64 *__restrict__ splitters, u32 nsplit, u64 *__restrict__ dest_counts)
{
    __shared__ u32 sh[kD64MaxRanks + 1];
    __shared__ u64 sp[kD64MaxRanks];
    if (threadIdx.x <= kD64MaxRanks) sh[threadIdx.x] = 0;
    if (threadIdx.x < nsplit) sp[threadIdx.x] = splitters[threadIdx.x];
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < count; i += (u64)gridDim.x * 256) {
        const u64 key = keys[i];
        u32 d = 0;
        for (u32 j = 0; j < nsplit; ++j) d += sp[j] <= key ? 1u : 0u;
        // keys travel RELATIVE to the lower splitter of their destination: the receiver sorts log2(G) fewer bits
        keys[i] = (key - (d ? sp[d - 1] : 0)) | ((u64)d << 56);
        atomicAdd(&sh[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x <= nsplit && sh[threadIdx.x]) atomicAdd((unsigned long long *)&dest_counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// routing key of a position: (owner << 56) | offset of pos + add inside the owner's block; positions at or beyond `limit` get
// owner = world (dropped)
static __global__ void __launch_bounds__(256)
d64_owner_keys_kernel(const u64 *__restrict__ pos, u64 count, u64 add, u64 limit, u64 block, u32 world,
                      u64 *__restrict__ keys, u64 *__restrict__ ident, u64 *__restrict__ dest_counts)
{
    __shared__ u32 sh[kD64MaxRanks + 1];
    if (threadIdx.x <= kD64MaxRanks) sh[threadIdx.x] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < count; i += (u64)gridDim.x * 256) {
        const u64 v = pos[i] + add;
        u64 owner = v < limit ? v / block : (u64)world;
        if (v < limit && owner >= world) owner = world - 1;
        keys[i] = (owner << 56) | (owner < world ? v - owner * block : v);
        if (ident != nullptr) ident[i] = i;
        atomicAdd(&sh[owner], 1u);
    }
    __syncthreads();
    if (threadIdx.x <= world && sh[threadIdx.x]) atomicAdd((unsigned long long *)&dest_counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

static const u64 kLow56 = (((u64)1) << 56) - 1;

// ISA[pos - lo] = rank for the received (routing key, rank) pairs
static __global__ void __launch_bounds__(256)
d64_scatter_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ ranks, u64 count, u64 len, u64 *__restrict__ ISA)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    const u64 j = keys[i] & kLow56;
    if (j < len) ISA[j] = ranks[i];
}
// answers to the received requests: ISA[q - lo] + 1
static __global__ void __launch_bounds__(256)
d64_gather_kernel(const u64 *__restrict__ req, u64 count, u64 len, const u64 *__restrict__ ISA, u64 *__restrict__ ans)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    const u64 j = req[i] & kLow56;
    ans[i] = j < len ? ISA[j] + 1 : 0;
}
// keys of a doubling round: (group << rank_bits) | k2, k2 = answer of the request that carried this element's id (0: none)
static __global__ void __launch_bounds__(256)
d64_fill_kernel(u64 *__restrict__ a, u64 count, u64 v)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < count) a[i] = v;
}
static __global__ void __launch_bounds__(256)
d64_place_answers_kernel(const u64 *__restrict__ ids, const u64 *__restrict__ ans, u64 count, u64 *__restrict__ k2)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < count) k2[ids[i]] = ans[i];
}
static __global__ void __launch_bounds__(256)
d64_round_keys_kernel(const u32 *__restrict__ grp, const u64 *__restrict__ k2, u64 count, int rank_bits, u64 *__restrict__ keys)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < count) keys[i] = ((u64)grp[i] << rank_bits) | k2[i];
}

// BWT rows of a slice of the suffix array: rows[j] = T[SA[j] - 1] (from the packed text through the inverse code map);
// the slot of suffix 0 is reported (primary index - 1)
static __global__ void __launch_bounds__(256)
d64_bwt_rows_kernel(const u64 *__restrict__ sa, u64 M, u64 base, const u64 *__restrict__ words, int b, const u8 *__restrict__ inv,
                    u8 *__restrict__ rows, u64 *__restrict__ primary)
{
    const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j >= M) return;
    const u64 p = sa[j];
    u8 v = 0;
    if (p == 0) *primary = base + j + 1;
    else v = inv[(u32)kmer64_at(words, p - 1, b, b)];
    rows[j] = v;
}
// aux samples of the owned positions: I[p / r] = ISA[p] + 1 for p % r == 0
static __global__ void __launch_bounds__(256)
d64_aux_kernel(const u64 *__restrict__ ISA, u64 lo, u64 count, u64 r, u64 first_idx, u64 nidx, i64 *__restrict__ out)
{
    const u64 t = (u64)blockIdx.x * 256 + threadIdx.x;
    if (t >= nidx) return;
    const u64 p = (first_idx + t) * r;
    if (p >= lo && p - lo < count) out[t] = (i64)ISA[p - lo] + 1;
}

// ---------------------------------------------------------------------------------------------
// 64-bit rank stage on a sorted slice of N < 2^32 elements (element j: key, position; global slot base + j in
// round 0, slot_in[j] later).  flags -> rank_scan (three exclusive scans of the tile aggregates) -> apply.
//   rank[j] = slot of the head of j's group; active = group of more than one element
// ---------------------------------------------------------------------------------------------
static const int kR64Threads = 256;
static const int kR64Tile = 1024;

static __global__ void __launch_bounds__(kR64Threads)
r64_flags_kernel(const u64 *__restrict__ keys, u64 N, u64 mask, u8 *__restrict__ flags, u32 *__restrict__ tagg, u64 ntiles)
{
    __shared__ u32 s_w[3][kR64Threads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tile = blockIdx.x;
    const u64 j0 = tile * kR64Tile + (u64)tid * 4;
    u32 last_head = 0, nact = 0, ngrp = 0;
    u64 kprev = j0 > 0 && j0 <= N ? keys[j0 - 1] & mask : 0;
    u64 kcur = j0 < N ? keys[j0] & mask : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u64 j = j0 + i;
        if (j < N) {
            const u64 knext = j + 1 < N ? keys[j + 1] & mask : 0;
            const bool head = j == 0 || kcur != kprev;
            const bool nhead = j + 1 >= N || knext != kcur;
            const bool active = !(head && nhead);
            flags[j] = (u8)((head ? 1 : 0) | (active ? 2 : 0));
            if (head) last_head = (u32)(j - tile * kR64Tile) + 1;
            nact += active ? 1 : 0;
            ngrp += (active && head) ? 1 : 0;
            kprev = kcur; kcur = knext;
        }
    }
    last_head = __reduce_max_sync(0xffffffffu, last_head);
    nact = __reduce_add_sync(0xffffffffu, nact);
    ngrp = __reduce_add_sync(0xffffffffu, ngrp);
    if (lane == 0) { s_w[0][warp] = last_head; s_w[1][warp] = nact; s_w[2][warp] = ngrp; }
    __syncthreads();
    if (tid < 3) {
        u32 r = 0;
        for (int w = 0; w < kR64Threads / 32; ++w) { const u32 v = s_w[tid][w]; r = tid == 0 ? (v > r ? v : r) : r + v; }
        // aggregate 0: (tile index * tile + local index of the last head) + 1 as a global local-index + 1, so a max-scan carries it
        if (tid == 0) r = r ? (u32)(tile * kR64Tile) + r : 0;
        tagg[(u64)tid * ntiles + tile] = r;
    }
}

template <bool ROUND0>
static __global__ void __launch_bounds__(kR64Threads)
r64_apply_kernel(const u64 *__restrict__ pos, const u64 *__restrict__ slot_in, const u8 *__restrict__ flags, const u32 *__restrict__ tagg,
                 u64 N, u64 ntiles, u64 base, u64 *__restrict__ rank_out, u64 *__restrict__ sa_local,
                 u64 *__restrict__ a_pos, u64 *__restrict__ a_slot, u32 *__restrict__ a_grp)
{
    __shared__ u32 s_h[kR64Threads / 32], s_a[kR64Threads / 32], s_g[kR64Threads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tile = blockIdx.x;
    const u64 j0 = tile * kR64Tile + (u64)tid * 4;
    u32 f[4]; u32 lh = 0, na = 0, ng = 0;                     // thread aggregates: last head (index + 1), actives, active groups
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u64 j = j0 + i;
        f[i] = j < N ? flags[j] : 0u;
        if (f[i] & 1u) lh = (u32)j + 1;
        na += (f[i] >> 1) & 1u;
        ng += ((f[i] & 3u) == 3u) ? 1u : 0u;
    }
    // exclusive scans over the threads of the block: max for the head, sums for the counts
    u32 ih = lh, ia = na, ig = ng;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u32 oh = __shfl_up_sync(0xffffffffu, ih, off), oa = __shfl_up_sync(0xffffffffu, ia, off), og = __shfl_up_sync(0xffffffffu, ig, off);
        if (lane >= off) { ih = oh > ih ? oh : ih; ia += oa; ig += og; }
    }
    if (lane == 31) { s_h[warp] = ih; s_a[warp] = ia; s_g[warp] = ig; }
    __syncthreads();
    u32 ch = tagg[tile], ca = tagg[ntiles + tile], cg = tagg[2 * ntiles + tile];       // exclusive prefixes of the tile
    for (int w = 0; w < warp; ++w) { ch = s_h[w] > ch ? s_h[w] : ch; ca += s_a[w]; cg += s_g[w]; }
    u32 eh = __shfl_up_sync(0xffffffffu, ih, 1);
    if (lane == 0) eh = 0;
    u32 curh = eh > ch ? eh : ch;                             // last head before my first element (index + 1)
    u32 cura = ca + ia - na, curg = cg + ig - ng;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u64 j = j0 + i;
        if (j < N) {
            if (f[i] & 1u) curh = (u32)j + 1;
            const u64 hidx = (u64)curh - 1;                   // every element has a head at or before it (element 0 is one)
            const u64 rk = ROUND0 ? base + hidx : slot_in[hidx];
            const u64 slot = ROUND0 ? base + j : slot_in[j];
            rank_out[j] = rk;
            const u64 p = pos[j];
            if (!ROUND0) sa_local[slot - base] = p;
            if (f[i] & 2u) {
                if ((f[i] & 3u) == 3u) ++curg;
                a_pos[cura] = p; a_slot[cura] = slot; a_grp[cura] = curg - 1;
                ++cura;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct HostBarrier {
    std::mutex m; std::condition_variable cv; int count = 0, gen = 0, n = 1;
    void wait() {
        std::unique_lock<std::mutex> l(m);
        const int g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(l, [&] { return g != gen; });
    }
};

// Device memory of one rank: ONE cudaMalloc per call (cudaMalloc / cudaFree take a process-wide lock and cost milliseconds for
// multi-GB blocks: with eight rank threads allocating dozens of buffers each they serialised the whole run), carved up by a
// first-fit free list on the host.  A request that does not fit falls back to cudaMalloc, so the size estimate is not critical.
struct Slab {
    char *base = nullptr; size_t cap = 0;
    std::vector<std::pair<size_t, size_t>> holes;          // (offset, size), sorted by offset
    bool init(size_t bytes) {
        if (cudaMalloc(&base, bytes) != cudaSuccess) { cudaGetLastError(); base = nullptr; return false; }
        cap = bytes; holes.assign(1, std::make_pair((size_t)0, bytes));
        return true;
    }
    void *alloc(size_t bytes) {
        bytes = (bytes + 511) & ~(size_t)511;
        for (size_t i = 0; i < holes.size(); ++i) if (holes[i].second >= bytes) {
            void *p = base + holes[i].first;
            holes[i].first += bytes; holes[i].second -= bytes;
            if (holes[i].second == 0) holes.erase(holes.begin() + i);
            return p;
        }
        return nullptr;
    }
    bool owns(const void *p) const { return base && (const char *)p >= base && (const char *)p < base + cap; }
    void release(void *p, size_t bytes) {
        bytes = (bytes + 511) & ~(size_t)511;
        const size_t off = (size_t)((char *)p - base);
        size_t i = 0;
        while (i < holes.size() && holes[i].first < off) ++i;
        holes.insert(holes.begin() + i, std::make_pair(off, bytes));
        if (i + 1 < holes.size() && holes[i].first + holes[i].second == holes[i + 1].first) { holes[i].second += holes[i + 1].second; holes.erase(holes.begin() + i + 1); }
        if (i > 0 && holes[i - 1].first + holes[i - 1].second == holes[i].first) { holes[i - 1].second += holes[i].second; holes.erase(holes.begin() + i); }
    }
    ~Slab() { if (base) cudaFree(base); }
};
static thread_local Slab *tl_slab = nullptr;               // the slab of the rank thread that is running

struct DevBuf {
    void *p = nullptr; size_t bytes = 0; Slab *from = nullptr;
    bool alloc(size_t b) {
        release();
        if (b == 0) b = 256;
        if (tl_slab != nullptr) { p = tl_slab->alloc(b); if (p) { from = tl_slab; bytes = b; return true; } }
        if (cudaMalloc(&p, b) != cudaSuccess) { cudaGetLastError(); p = nullptr; return false; }
        from = nullptr; bytes = b; return true;
    }
    void release() {
        if (p) { if (from) from->release(p, bytes); else cudaFree(p); }
        p = nullptr; bytes = 0; from = nullptr;
    }
    ~DevBuf() { release(); }
    template <typename T> T *as() const { return (T *)p; }
};

struct DistGroup {
    int G = 1;
    std::vector<int> devs;
    HostBarrier bar;
    std::atomic<int> failed{0};
    u64 n = 0, B = 0;
    const u8 *T = nullptr; i64 *SA = nullptr;
    u8 *rows = nullptr;              // BWT mode: host scratch of n bytes for the per-slot rows (the caller's A array)
    i64 *aux_I = nullptr; u64 aux_r = 0;
    std::atomic<unsigned long long> primary{0};
    // shared metadata, indexed by rank
    std::vector<u64> cnt;            // [G][G+1] send counts of the current exchange: cnt[src * (G+1) + dst]
    std::vector<void *> pk, pv;      // receive buffers of the current exchange (keys / values)
    std::vector<void *> pwords;      // replicated packed text
    std::vector<void *> pans;        // answer buffers
    std::vector<u64> scal;           // [G][8] small per-rank scalars (all-reduce / all-gather through the host)
    std::vector<u64> hist;           // [G][256]
    std::vector<u64> samples;        // [G][S]
    // decided by rank 0 after the histogram
    int b = 8, k = 1, K = 8, len_bits = 1;
    u8 lut[256];
    std::vector<u64> splitters;
    libsais_cuda_dist_stats stats;
    bool want_verify = false;
    std::mutex stats_mutex;
};

static const int kD64Samples = 4096;

struct DistRank {
    DistGroup &g; const int r; Ctx *c = nullptr;
    Slab slab;                                 // declared before the buffers: destroyed after them
    u64 lo = 0, hi = 0, m = 0;                 // owned positions
    DevBuf words, isa, sa, bufK[2], bufV[2], rankbuf, flags, tagg, aPos[2], aSlot[2], aGrp, k2, ids, ans, req, cntd, ptrs, misc;
    u64 exchanged_bytes = 0;
    DistRank(DistGroup &grp, int rank) : g(grp), r(rank) {}

    bool fail() { g.failed.store(1); return false; }
    bool sync_all_at(bool ok, int line) {      // barrier + common failure check: every rank takes the same branch afterwards
        if (!ok || (c && c->failed())) {
            static const bool dbg = [] { const char *e = getenv("LIBSAIS_CUDA_DEBUG"); return e && *e && atoi(e) != 0; }();
            if (dbg && g.failed.load() == 0)
                fprintf(stderr, "libsais_cuda dist64: rank %d failed before dist64.cu:%d (ok=%d, cuda error %d: %s)\n", r, line, (int)ok,
                        c ? (int)c->last_error : -1, c ? cudaGetErrorString(c->last_error) : "no context");
            g.failed.store(1);
        }
        g.bar.wait();
        const bool good = g.failed.load() == 0;
        g.bar.wait();
        return good;
    }
    u32 grid_for(u64 count) const { const u64 w = ceil_div(count, 256); return (u32)(w ? w : 1); }
    u32 grid_stride(u64 count) const { const u64 w = ceil_div(count, 256 * 8); const u64 cap = (u64)c->sm_count * 16; return (u32)(w < cap ? (w ? w : 1) : cap); }

    // counts of the current exchange -> host matrix
    bool publish_counts() {
        c->check(cudaMemcpyAsync(c->h_scalars + S_ROUTE, cntd.as<u64>(), (g.G + 1) * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        if (!c->sync()) return false;
        for (int d = 0; d <= g.G; ++d) g.cnt[(size_t)r * (g.G + 1) + d] = c->h_scalars[S_ROUTE + d];
        return true;
    }
    u64 recv_total() const { u64 t = 0; for (int s = 0; s < g.G; ++s) t += g.cnt[(size_t)s * (g.G + 1) + r]; return t; }
    u64 recv_offset(int src, int dst) const { u64 t = 0; for (int s = 0; s < src; ++s) t += g.cnt[(size_t)s * (g.G + 1) + dst]; return t; }
    u64 send_offset(int dst) const { u64 t = 0; for (int d = 0; d < dst; ++d) t += g.cnt[(size_t)r * (g.G + 1) + d]; return t; }

    // The fused route + exchange: partition (keys, vals) by the destination in the keys' top byte; destination d's run is
    // written straight into kdst[d] / vdst[d] (peer memory).  Returns false on failure.
    bool route(const u64 *keys, const u64 *vals, u64 count, void **kdst, void **vdst, bool vals_remote = true) {
        if (count == 0) return true;
        c->reset_arena();
        void **h = (void **)(c->h_scalars + S_SGRAM);                      // pinned staging of the two pointer tables (G + 1 entries each)
        const int ne = kD64MaxRanks + 1;
        for (int d = 0; d < ne; ++d) { h[d] = d <= g.G ? kdst[d] : nullptr; h[ne + d] = d <= g.G ? vdst[d] : nullptr; }
        c->check(cudaMemcpyAsync(ptrs.p, h, ne * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
        c->check(cudaMemcpyAsync((void **)ptrs.p + kRadixSize, h + ne, ne * sizeof(void *), cudaMemcpyHostToDevice, c->stream));
        const u64 nt = ceil_div(count, (u64)part_tile());
        const size_t stw = count < (1ull << 30) ? sizeof(u32) : sizeof(u64);
        void *status = c->alloc(nt * kRadixSize * stw);
        if (!status) return false;
        c->check(cudaMemsetAsync(status, 0, nt * kRadixSize * stw, c->stream));
        PartArgs pa; pa.n = count; pa.shift = 56; pa.dmask = 255u; pa.base = nullptr; pa.cp = nullptr; pa.nseg = 1; pa.tpc = (u32)nt;
        pa.boff = nullptr; pa.tstart = nullptr; pa.tinfo = nullptr; pa.ticket = nullptr; pa.err = (u32 *)(c->d_scalars + S_ERR); pa.use_bulk = 0;
        pa.kptr = (void *const *)ptrs.p; pa.vptr = (void *const *)((void **)ptrs.p + kRadixSize);
        launch_part_pass<u64, u64, ArraySrc, false>(*c, KC_SCATTER, (double)count * 32.0, ArraySrc(), keys, vals, (u64 *)nullptr, (u64 *)nullptr, pa, nt, status);
        exchanged_bytes += (count - g.cnt[(size_t)r * (g.G + 1) + r] - g.cnt[(size_t)r * (g.G + 1) + g.G]) * (vals_remote ? 16 : 8);   // bytes that leave this GPU
        if (!c->sync()) return false;                                      // the pointer staging is reused by the next route
        return !c->failed();
    }

    int run();
    bool rank_stage(const u64 *keys, const u64 *pos, const u64 *slot_in, u64 N, u64 mask, u64 base, u64 *rank_out, u64 *sa_local,
                    u64 *a_pos, u64 *a_slot, u32 *a_grp, u64 *n_act, u64 *n_grp);
    bool update_isa(const u64 *pos, const u64 *rank, u64 count);
    bool fetch_isa(const u64 *pos, u64 count, u64 add, u64 *out, bool ok_in);
    bool verify(const u64 *sa_slice, u64 M, u64 base);
};

bool DistRank::rank_stage(const u64 *keys, const u64 *pos, const u64 *slot_in, u64 N, u64 mask, u64 base, u64 *rank_out, u64 *sa_local,
                          u64 *a_pos, u64 *a_slot, u32 *a_grp, u64 *n_act, u64 *n_grp)
{
    *n_act = *n_grp = 0;
    if (N == 0) return true;
    const u64 ntiles = ceil_div(N, (u64)kR64Tile);
    if (!flags.p || flags.bytes < N) { if (!flags.alloc(N)) return false; }
    if (!tagg.p || tagg.bytes < ntiles * 3 * sizeof(u32)) { if (!tagg.alloc(ntiles * 3 * sizeof(u32))) return false; }
    LSC_LAUNCH(*c, KC_RANK_INIT, (double)N * 9, r64_flags_kernel, (u32)ntiles, kR64Threads, 0, keys, N, mask, flags.as<u8>(), tagg.as<u32>(), ntiles);
    run_rank_scan(*c, tagg.as<u32>(), ntiles, c->d_scalars + S_NACT);   // sa_core.cu: max / sum / sum scans of the tile aggregates
    if (slot_in == nullptr)
        LSC_LAUNCH(*c, KC_RANK_INIT, (double)N * 33, r64_apply_kernel<true>, (u32)ntiles, kR64Threads, 0, pos, slot_in, flags.as<u8>(), tagg.as<u32>(),
                   N, ntiles, base, rank_out, sa_local, a_pos, a_slot, a_grp);
    else
        LSC_LAUNCH(*c, KC_RANK_UPDATE, (double)N * 49, r64_apply_kernel<false>, (u32)ntiles, kR64Threads, 0, pos, slot_in, flags.as<u8>(), tagg.as<u32>(),
                   N, ntiles, base, rank_out, sa_local, a_pos, a_slot, a_grp);
    c->check(cudaMemcpyAsync(c->h_scalars + S_NACT, c->d_scalars + S_NACT, 2 * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    if (!c->sync() || c->failed()) return false;
    *n_act = c->h_scalars[S_NACT]; *n_grp = c->h_scalars[S_NGRP];
    return true;
}

// (position, rank) pairs -> the owners of the positions -> their ISA slices.  Collective: every rank calls it.
bool DistRank::update_isa(const u64 *pos, const u64 *rank, u64 count)
{
    const int G = g.G;
    bool ok = true;
    DevBuf rk;                                        // routing keys of my pairs
    ok = rk.alloc((count + 1) * sizeof(u64));
    c->check(cudaMemsetAsync(cntd.p, 0, (kD64MaxRanks + 1) * sizeof(u64), c->stream));
    if (ok && count) LSC_LAUNCH(*c, KC_SCATTER, (double)count * 16, d64_owner_keys_kernel, grid_stride(count), 256, 0,
                                pos, count, (u64)0, g.n, g.B, (u32)G, rk.as<u64>(), (u64 *)nullptr, cntd.as<u64>());
    ok = ok && publish_counts();
    if (!sync_all_at(ok, __LINE__)) return false;
    const u64 total = recv_total();
    DevBuf rkeys, rvals;
    ok = rkeys.alloc((total + 1) * sizeof(u64)) && rvals.alloc((total + 1) * sizeof(u64));
    g.pk[r] = rkeys.p; g.pv[r] = rvals.p;
    if (!sync_all_at(ok, __LINE__)) return false;
    void *kd[kD64MaxRanks + 1], *vd[kD64MaxRanks + 1];
    for (int d = 0; d < G; ++d) { kd[d] = (u64 *)g.pk[d] + recv_offset(r, d); vd[d] = (u64 *)g.pv[d] + recv_offset(r, d); }
    kd[G] = misc.p; vd[G] = misc.p;                   // nothing is dropped here (every position is < n)
    ok = route(rk.as<u64>(), rank, count, kd, vd);
    if (!sync_all_at(ok, __LINE__)) return false;                  // all peers' stores into my buffers are complete
    rk.release();
    const u64 *fk = rkeys.as<u64>(), *fv = rvals.as<u64>();
    DevBuf ak, av;
    if (total >= ((u64)1 << 22) && m > ((u64)1 << 23)) {
        // a random scatter of 8-byte words over a multi-GB slice runs at ~20 G/s: group the pairs by the top 8 bits of the
        // offset first (one stable digit pass), so the stores walk the slice window by window and merge in L2 (scatter.cuh)
        const int hb = bits_for(m - 1);
        bool fits = ak.alloc((total + 1) * 8) && av.alloc((total + 1) * 8);
        c->reset_arena();
        void *temp = fits ? c->alloc(RadixSort<u64, u64>::temp_bytes(total)) : nullptr;
        if (fits && temp) {
            const int w = RadixSort<u64, u64>::sort(*c, rkeys.as<u64>(), rvals.as<u64>(), ak.as<u64>(), av.as<u64>(), total, hb > 8 ? hb - 8 : 0, hb, temp,
                                                    (u32 *)(c->d_scalars + S_ERR));
            if (w == 1) { fk = ak.as<u64>(); fv = av.as<u64>(); }
            else if (w < 0) ok = false;
        } else c->last_error = cudaSuccess;                          // no room: plain scatter
    }
    if (total) LSC_LAUNCH(*c, KC_SCATTER, (double)total * 24, d64_scatter_kernel, grid_for(total), 256, 0, fk, fv, total, m, isa.as<u64>());
    ok = ok && c->sync() && !c->failed();
    return sync_all_at(ok, __LINE__);                              // buffers are freed on return: nobody may still be writing
}

// out[j] = ISA[pos[j] + add] + 1 (0 when pos[j] + add >= n) for my `count` positions.  Collective: every rank calls it.
// Requests go to the owners with the fused route (keys into the peers' request buffers, the element ids stay local in
// destination order); every owner answers with one contiguous peer copy per requester.
bool DistRank::fetch_isa(const u64 *pos, u64 count, u64 add, u64 *out, bool ok_in)
{
    const int G = g.G;
    bool ok = ok_in;
    DevBuf rk, idloc, idsrc, reqb, ansb, myans;
    ok = ok && rk.alloc((count + 1) * 8) && idsrc.alloc((count + 1) * 8) && idloc.alloc((count + 1) * 8);
    c->check(cudaMemsetAsync(cntd.p, 0, (kD64MaxRanks + 1) * 8, c->stream));
    if (ok && count) LSC_LAUNCH(*c, KC_ROUND_KEYS, (double)count * 24, d64_owner_keys_kernel, grid_stride(count), 256, 0,
                                pos, count, add, g.n, g.B, (u32)G, rk.as<u64>(), idsrc.as<u64>(), cntd.as<u64>());
    ok = ok && publish_counts();
    if (!sync_all_at(ok, __LINE__)) return false;
    const u64 nreq_in = recv_total();
    u64 nreq_out = 0;
    for (int d = 0; d < G; ++d) nreq_out += g.cnt[(size_t)r * (G + 1) + d];
    ok = reqb.alloc((nreq_in + 1) * 8) && ansb.alloc((nreq_out + 1) * 8);
    g.pk[r] = reqb.p; g.pans[r] = ansb.p;
    if (!sync_all_at(ok, __LINE__)) return false;
    {
        void *kd[kD64MaxRanks + 1], *vd[kD64MaxRanks + 1];
        for (int d = 0; d < G; ++d) { kd[d] = (u64 *)g.pk[d] + recv_offset(r, d); vd[d] = idloc.as<u64>() + send_offset(d); }
        DevBuf dump;
        ok = dump.alloc((g.cnt[(size_t)r * (G + 1) + G] + 1) * 8);
        kd[G] = dump.p; vd[G] = dump.p;
        ok = ok && route(rk.as<u64>(), idsrc.as<u64>(), count, kd, vd, false);
    }
    if (!sync_all_at(ok, __LINE__)) return false;
    ok = myans.alloc((nreq_in + 1) * 8);
    if (ok && nreq_in) LSC_LAUNCH(*c, KC_ROUND_KEYS, (double)nreq_in * 24, d64_gather_kernel, grid_for(nreq_in), 256, 0,
                                  reqb.as<u64>(), nreq_in, m, isa.as<u64>(), myans.as<u64>());
    if (ok) {
        for (int s = 0; s < G; ++s) {
            const u64 cnt_s = g.cnt[(size_t)s * (G + 1) + r];
            if (!cnt_s) continue;
            u64 off_in_s = 0;                                     // my block inside s's destination-ordered send list
            for (int d = 0; d < r; ++d) off_in_s += g.cnt[(size_t)s * (G + 1) + d];
            c->check(cudaMemcpyPeerAsync((u64 *)g.pans[s] + off_in_s, g.devs[s], myans.as<u64>() + recv_offset(s, r), g.devs[r], cnt_s * 8, c->stream));
        }
        exchanged_bytes += (nreq_in - g.cnt[(size_t)r * (G + 1) + r]) * 8;   // answers that leave this GPU
        ok = c->sync();
    }
    if (!sync_all_at(ok, __LINE__)) return false;
    if (count) {
        LSC_LAUNCH(*c, KC_ROUND_KEYS, (double)count * 8, d64_fill_kernel, grid_for(count), 256, 0, out, count, (u64)0);
        if (nreq_out) LSC_LAUNCH(*c, KC_ROUND_KEYS, (double)nreq_out * 24, d64_place_answers_kernel, grid_for(nreq_out), 256, 0, idloc.as<u64>(), ansb.as<u64>(), nreq_out, out);
    }
    ok = c->sync() && !c->failed();
    return sync_all_at(ok, __LINE__);                             // the request / answer buffers die here: nobody may still be using them
}

// Distributed result check (no CPU reference exists at 2^34 symbols): with the final ISA slices,
//   (1) ISA[SA[i]] == i for every slot i          => SA is a permutation of [0, n) and ISA its inverse
//   (2) (T[SA[i-1]], ISA[SA[i-1]+1]) < (T[SA[i]], ISA[SA[i]+1]) for every i > 0 (rank past the end = -1)
//       => every neighbouring pair is in suffix order (Burkhardt-Kaerkkaeinen)
// which together prove the suffix array.  Two fetch_isa rounds + one streaming kernel; slice boundaries through the host.
static __global__ void __launch_bounds__(256)
d64_verify_kernel(const u64 *__restrict__ sa, const u64 *__restrict__ r0, const u64 *__restrict__ r1, u64 M, u64 base,
                  const u64 *__restrict__ words, int b, u64 prev_sym, u64 prev_r1, int has_prev, u64 *__restrict__ bad, u64 *__restrict__ edge)
{
    const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j >= M) return;
    const u64 p = sa[j];
    const u64 sym = kmer64_at(words, p, b, b);
    u64 nbad = 0;
    if (r0[j] != base + j + 1) nbad = 1;
    u64 ps, pr; bool have = true;
    if (j > 0) { ps = kmer64_at(words, sa[j - 1], b, b); pr = r1[j - 1]; }
    else if (has_prev) { ps = prev_sym; pr = prev_r1; }
    else { have = false; ps = pr = 0; }
    if (have && !(ps < sym || (ps == sym && pr < r1[j]))) nbad = 1;
    if (nbad) atomicAdd((unsigned long long *)bad, 1ull);
    if (j == M - 1) { edge[0] = sym; edge[1] = r1[j]; }
}

bool DistRank::verify(const u64 *sa_slice, u64 M, u64 base)
{
    // in chunks of 2^27 slots, so the request / answer buffers stay small next to the resident SA and ISA slices;
    // every rank runs the same number of chunks (the collectives inside fetch_isa must match)
    const u64 CH = (u64)1 << 27;
    g.scal[(size_t)r * 8 + 2] = M;
    if (!sync_all_at(true, __LINE__)) return false;
    u64 maxM = 0, total_m = 0;
    for (int q = 0; q < g.G; ++q) { maxM = std::max(maxM, g.scal[(size_t)q * 8 + 2]); total_m += g.scal[(size_t)q * 8 + 2]; }
    const u64 nchunks = ceil_div(maxM, CH);
    DevBuf r0, r1, res;
    const u64 cap = std::min(CH, M) + 1;
    bool ok = r0.alloc(cap * 8) && r1.alloc(cap * 8) && res.alloc(64);
    if (ok) c->check(cudaMemsetAsync(res.p, 0, 64, c->stream));
    // the (symbol, next rank) of the element before the current chunk: from my previous chunk, or the last element of the
    // nearest non-empty slice below mine (known after that rank's last chunk: published through the host at the end)
    int has_prev = 0; u64 ps = 0, pr = 0;
    u64 first_sym = 0, first_r1 = 0;             // my very first element, compared with the previous slice at the end
    u64 last_sym = 0, last_r1 = 0;
    for (u64 ch = 0; ch < nchunks; ++ch) {
        const u64 c0 = std::min(M, ch * CH), c1 = std::min(M, c0 + CH), cn = c1 - c0;
        if (!fetch_isa(sa_slice + c0, cn, 0, r0.as<u64>(), ok)) return false;
        if (!fetch_isa(sa_slice + c0, cn, 1, r1.as<u64>(), true)) return false;
        if (cn) {
            LSC_LAUNCH(*c, KC_CONVERT, (double)cn * 24, d64_verify_kernel, grid_for(cn), 256, 0, sa_slice + c0, r0.as<u64>(), r1.as<u64>(), cn, base + c0,
                       words.as<u64>(), g.b, ps, pr, has_prev, res.as<u64>(), res.as<u64>() + 1);
            u64 e[2], f[2];
            c->check(cudaMemcpyAsync(e, res.as<u64>() + 1, 16, cudaMemcpyDeviceToHost, c->stream));
            c->check(cudaMemcpyAsync(f, sa_slice + c0, 8, cudaMemcpyDeviceToHost, c->stream));
            c->check(cudaMemcpyAsync(f + 1, r1.as<u64>(), 8, cudaMemcpyDeviceToHost, c->stream));
            ok = c->sync() && !c->failed();
            if (ok) {
                ps = e[0]; pr = e[1]; has_prev = 1; last_sym = e[0]; last_r1 = e[1];
                if (c0 == 0) { first_sym = (u64)g.lut[g.T[f[0]]]; first_r1 = f[1]; }
            }
        }
    }
    u64 nbad = 0;
    c->check(cudaMemcpyAsync(&nbad, res.p, 8, cudaMemcpyDeviceToHost, c->stream));
    ok = ok && c->sync() && !c->failed();
    g.scal[(size_t)r * 8 + 3] = last_sym; g.scal[(size_t)r * 8 + 4] = last_r1; g.scal[(size_t)r * 8 + 5] = nbad;
    if (!sync_all_at(ok, __LINE__)) return false;
    // slice boundary: my first element against the last element of the nearest non-empty slice below
    u64 edge_bad = 0;
    if (M) for (int q = r - 1; q >= 0; --q) if (g.scal[(size_t)q * 8 + 2]) {
        const u64 qs = g.scal[(size_t)q * 8 + 3], qr = g.scal[(size_t)q * 8 + 4];
        if (!(qs < first_sym || (qs == first_sym && qr < first_r1))) edge_bad = 1;
        break;
    }
    g.scal[(size_t)r * 8 + 6] = edge_bad;
    if (!sync_all_at(true, __LINE__)) return false;
    u64 total_bad = 0;
    for (int q = 0; q < g.G; ++q) total_bad += g.scal[(size_t)q * 8 + 5] + g.scal[(size_t)q * 8 + 6];
    if (r == 0) { std::lock_guard<std::mutex> lk(g.stats_mutex); g.stats.verify = (total_bad == 0 && total_m == g.n) ? 1 : -1; g.stats.verify_violations = total_bad; }
    return sync_all_at(true, __LINE__);
}

int DistRank::run()
{
    const int G = g.G;
    DeviceGuard guard(g.devs[r]);
    bool ok = true;
    // ---- context, peer access
    {
        std::unique_ptr<Ctx> holder(new (std::nothrow) Ctx());
        c = holder.get();
        ok = c && c->init(g.devs[r]);
        if (ok) {
            for (int p = 0; p < G; ++p) if (p != r && g.devs[p] != g.devs[r]) {
                cudaError_t e = cudaDeviceEnablePeerAccess(g.devs[p], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                cudaGetLastError();
            }
        }
        if (!sync_all_at(ok, __LINE__)) { if (c) c->destroy(); return -2; }
        holder.release();
    }
    struct CtxCleanup { Ctx *c; ~CtxCleanup() { if (c) { c->destroy(); delete c; } } } cleanup{c};
    cudaStream_t st = c->stream;
    const u64 n = g.n;
    lo = std::min(n, (u64)r * g.B); hi = std::min(n, lo + g.B); m = hi - lo;
    u32 *err = (u32 *)(c->d_scalars + S_ERR);
    c->check(cudaMemsetAsync(c->d_scalars + S_ERR, 0, (S_MISC - S_ERR) * sizeof(u64), st));
    ok = c->reserve((size_t)(RadixSort<u64, u64>::temp_bytes(m + m / 2 + 4096) + (ceil_div(m + m / 2, 3072) + 4096) * kRadixSize * 8 + (64 << 20)));
    if (ok) {   // the rank's slab: what the run is expected to need, at most this rank's share of the free memory
        size_t fr = 0, tot = 0;
        int share = 0;
        for (int q = 0; q < G; ++q) share += g.devs[q] == g.devs[r] ? 1 : 0;
        if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); fr = 0; }
        const size_t want = (size_t)(m + m / 6) * 88 + (size_t)n + ((size_t)512 << 20);
        size_t give = (size_t)((double)fr * 0.94 / share);
        if (give > want) give = want;
        if (give >= ((size_t)64 << 20) && slab.init(give)) tl_slab = &slab;      // else: plain cudaMalloc per buffer
        ok = cntd.alloc((kD64MaxRanks + 1) * sizeof(u64)) && ptrs.alloc(2 * kRadixSize * sizeof(void *)) && misc.alloc(1 << 20);
    }
    struct SlabGuard { ~SlabGuard() { tl_slab = nullptr; } } slab_guard;
    if (!sync_all_at(ok, __LINE__)) return -2;

    // ---- own slice of the text -> histogram -> (all ranks) alphabet -> packed slice -> replicated packed text
    DevBuf dT;
    ok = dT.alloc(m + 64);
    if (ok && m) {
        c->check(cudaMemsetAsync((char *)dT.p + m, 0, 64, st));
        c->check(cudaMemcpyAsync(dT.p, g.T + lo, m, cudaMemcpyHostToDevice, st));
        run_byte_histogram(*c, dT.as<u8>(), m);
        c->check(cudaMemcpyAsync(c->h_scalars + S_FREQ, c->d_scalars + S_FREQ, 256 * sizeof(u64), cudaMemcpyDeviceToHost, st));
        ok = c->sync();
    }
    for (int s = 0; s < 256; ++s) g.hist[(size_t)r * 256 + s] = (ok && m) ? c->h_scalars[S_FREQ + s] : 0;
    if (!sync_all_at(ok, __LINE__)) return -2;
    if (r == 0) {
        int sigma = 0; double entropy = 0;
        for (int s = 0; s < 256; ++s) {
            u64 f = 0;
            for (int q = 0; q < G; ++q) f += g.hist[(size_t)q * 256 + s];
            g.hist[s] = f;                                            // rank 0's row becomes the total
            g.lut[s] = (u8)sigma;
            if (f) { ++sigma; const double pr = (double)f / (double)n; entropy -= pr * std::log2(pr); }
        }
        g.b = bits_for((u64)(sigma > 1 ? sigma - 1 : 1));
        int kmax = 48 / g.b; if (kmax < 1) kmax = 1;
        const double need = std::log2((double)(n < 2 ? 2 : n)) + 10.0;
        const double kk = std::ceil(need / (entropy < 0.05 ? 0.05 : entropy));
        g.k = kk > (double)kmax ? kmax : (int)kk;
        if (g.k < 1) g.k = 1;
        g.K = g.k * g.b; g.len_bits = bits_for((u64)g.k);
    }
    if (!sync_all_at(true, __LINE__)) return -2;
    const int b = g.b, k = g.k, K = g.K, len_bits = g.len_bits;
    const int key_bits = K + len_bits;                               // <= 48 + 6: the top byte stays free for the destination
    const u64 nwords = ceil_div(n * (u64)b, 64) + 2;
    ok = words.alloc(nwords * 8);
    if (ok) c->check(cudaMemsetAsync(words.p, 0, nwords * 8, st));
    g.pwords[r] = words.p;
    DevBuf myw;
    const u64 my_nw = ceil_div(m * (u64)b, 64);
    ok = ok && myw.alloc((my_nw + 1) * 8);
    if (ok && m) {
        u8 *h_lut = (u8 *)(c->h_scalars + S_MISC);
        for (int s = 0; s < 256; ++s) h_lut[s] = g.lut[s];
        c->check(cudaMemcpyAsync(misc.p, h_lut, 256, cudaMemcpyHostToDevice, st));
        LSC_LAUNCH(*c, KC_PACK, (double)m + (double)my_nw * 8, d64_pack_kernel, grid_for(my_nw), 256, 0, dT.as<u8>(), m, b, myw.as<u64>(), my_nw, misc.as<u8>());
    }
    ok = ok && c->sync();
    if (!sync_all_at(ok, __LINE__)) return -2;                                    // everybody's words buffer exists and is zeroed
    if (m) {
        const u64 woff = (lo * (u64)b) >> 6;                          // B is a multiple of 64 symbols: whole words
        for (int p = 0; p < G; ++p)
            c->check(cudaMemcpyPeerAsync((u64 *)g.pwords[p] + woff, g.devs[p], myw.p, g.devs[r], my_nw * 8, st));
        exchanged_bytes += my_nw * 8 * (u64)(G - 1);
    }
    ok = c->sync();
    dT.release();
    if (!sync_all_at(ok, __LINE__)) return -2;
    myw.release();
    const auto t_start = std::chrono::steady_clock::now();           // the packed text is resident on every GPU
    auto t_last = t_start;
    auto mark = [&](int slot) {                                       // rank 0, right after a barrier: phase wall times
        if (r != 0) return;
        const auto now = std::chrono::steady_clock::now();
        std::lock_guard<std::mutex> lk(g.stats_mutex);
        g.stats.phase_seconds[slot] += std::chrono::duration<double>(now - t_last).count();
        t_last = now;
    };

    // ---- round 0: keys of the owned positions, splitters from a sample, fused route to the key owners, local sort
    const u64 cap0 = m + 1;
    ok = bufK[0].alloc(cap0 * 8) && bufV[0].alloc(cap0 * 8);
    if (ok && m) LSC_LAUNCH(*c, KC_MAKE_KEYS, (double)m * 18, d64_keys_kernel, grid_for(m), 256, 0, words.as<u64>(), n, b, k, K, len_bits, lo, m, bufK[0].as<u64>(), bufV[0].as<u64>());
    {   // sample: every (m / S)-th key (keys of an iid text are exchangeable; a structured text only skews the balance, not the result)
        std::vector<u64> smp(kD64Samples, ~0ull);
        if (ok && m) {
            DevBuf ds;
            ok = ds.alloc(kD64Samples * 8);
            const u64 step = m / kD64Samples ? m / kD64Samples : 1;
            if (ok) {
                c->check(cudaMemcpy2DAsync(ds.p, 8, bufK[0].p, step * 8, 8, std::min<u64>(kD64Samples, m), cudaMemcpyDeviceToDevice, st));
                c->check(cudaMemcpyAsync(smp.data(), ds.p, std::min<u64>(kD64Samples, m) * 8, cudaMemcpyDeviceToHost, st));
                ok = c->sync();
            }
        }
        for (int i = 0; i < kD64Samples; ++i) g.samples[(size_t)r * kD64Samples + i] = smp[i];
    }
    if (!sync_all_at(ok, __LINE__)) return -2;
    if (r == 0) {
        std::vector<u64> all(g.samples);
        all.erase(std::remove(all.begin(), all.end(), ~0ull), all.end());
        std::sort(all.begin(), all.end());
        g.splitters.assign(G - 1, ~0ull >> 8);
        for (int i = 1; i < G && !all.empty(); ++i) g.splitters[i - 1] = all[(size_t)i * all.size() / G];
    }
    if (!sync_all_at(true, __LINE__)) return -2;
    DevBuf dsp;
    ok = dsp.alloc(kD64MaxRanks * 8);
    if (ok) {
        if (G > 1) c->check(cudaMemcpyAsync(dsp.p, g.splitters.data(), (G - 1) * 8, cudaMemcpyHostToDevice, st));
        c->check(cudaMemsetAsync(cntd.p, 0, (kD64MaxRanks + 1) * 8, st));
        if (m) LSC_LAUNCH(*c, KC_SCATTER, (double)m * 16, d64_dest_splitters_kernel, grid_stride(m), 256, 0, bufK[0].as<u64>(), m, dsp.as<u64>(), (u32)(G - 1), cntd.as<u64>());
        ok = publish_counts();
    }
    if (!sync_all_at(ok, __LINE__)) return -2;
    const u64 M = recv_total();                                       // my slice of the suffix array
    u64 base = 0, maxM = 0;
    for (int d = 0; d < G; ++d) { u64 t = 0; for (int s = 0; s < G; ++s) t += g.cnt[(size_t)s * (G + 1) + d]; if (d < r) base += t; maxM = std::max(maxM, t); }
    if (M >= (1ull << 32) - 4096) ok = false;                        // local indexes are 32-bit
    ok = ok && bufK[1].alloc((M + 1) * 8) && bufV[1].alloc((M + 1) * 8);
    g.pk[r] = bufK[1].p; g.pv[r] = bufV[1].p;
    if (!sync_all_at(ok, __LINE__)) return -2;
    {
        void *kd[kD64MaxRanks + 1], *vd[kD64MaxRanks + 1];
        for (int d = 0; d < G; ++d) { kd[d] = (u64 *)g.pk[d] + recv_offset(r, d); vd[d] = (u64 *)g.pv[d] + recv_offset(r, d); }
        kd[G] = misc.p; vd[G] = misc.p;
        ok = route(bufK[0].as<u64>(), bufV[0].as<u64>(), m, kd, vd);
    }
    if (!sync_all_at(ok, __LINE__)) return -2;
    mark(0);
    // local sort of the received pairs on the key bits (the destination byte is above them)
    ok = bufK[0].alloc((M + 1) * 8) && bufV[0].alloc((M + 1) * 8);
    int where = 1;
    if (ok && M) {
        c->reset_arena();
        void *temp = c->alloc(RadixSort<u64, u64>::temp_bytes(M));
        ok = temp != nullptr;
        if (ok) {
            // my keys are relative to my lower splitter: their width is that of my key range
            const u64 range_lo = r ? g.splitters[r - 1] : 0, range_hi = r + 1 < G ? g.splitters[r] : ((u64)1 << key_bits);
            const int my_bits = range_hi > range_lo ? bits_for(range_hi - range_lo) : 1;
            const int w = RadixSort<u64, u64>::sort(*c, bufK[1].as<u64>(), bufV[1].as<u64>(), bufK[0].as<u64>(), bufV[0].as<u64>(), M, 0,
                                                    my_bits < key_bits ? my_bits : key_bits, temp, err);
            ok = w >= 0;
            where = w == 0 ? 1 : 0;                                   // sort() returns 0 when the result is in its first buffer pair (= bufK[1])
        }
    }
    ok = ok && c->sync();
    if (!sync_all_at(ok, __LINE__)) return -2;
    mark(1);
    u64 *sk = where ? bufK[1].as<u64>() : bufK[0].as<u64>();          // sorted keys
    std::swap(sa.p, bufV[where].p); std::swap(sa.bytes, bufV[where].bytes); std::swap(sa.from, bufV[where].from);   // the SA slice owns the sorted positions from here on
    u64 *sp = sa.as<u64>();                                           // sorted positions = my slice of the SA (singletons are final)
    const int other = where ? 0 : 1;
    // rank stage: heads, ranks (global slot of the group head), unresolved suffixes
    u64 nact = 0, ngrp = 0;
    ok = ok && rankbuf.alloc((M + 1) * 8);
    // actives: worst case M; the buffer pair that does not hold the sorted data provides two of the arrays
    ok = ok && aGrp.alloc((M + 1) * 4) && aSlot[0].alloc((M + 1) * 8);
    u64 *a_pos0 = bufK[other].as<u64>();
    if (ok) ok = rank_stage(sk, sp, nullptr, M, ((u64)1 << 56) - 1, base, rankbuf.as<u64>(), sp, a_pos0, aSlot[0].as<u64>(), aGrp.as<u32>(), &nact, &ngrp);
    if (!sync_all_at(ok, __LINE__)) return -2;
    // keep the actives in right-sized buffers, free the big ones
    ok = aPos[0].alloc((nact + 1) * 8) && aPos[1].alloc((nact + 1) * 8) && aSlot[1].alloc((nact + 1) * 8);
    if (ok && nact) c->check(cudaMemcpyAsync(aPos[0].p, a_pos0, nact * 8, cudaMemcpyDeviceToDevice, st));
    {   // the slot / group arrays were sized for the worst case (every suffix unresolved): move them to right-sized buffers
        DevBuf s2, g2;
        ok = ok && s2.alloc((nact + 1) * 8) && g2.alloc((nact + 1) * 4);
        if (ok && nact) {
            c->check(cudaMemcpyAsync(s2.p, aSlot[0].p, nact * 8, cudaMemcpyDeviceToDevice, st));
            c->check(cudaMemcpyAsync(g2.p, aGrp.p, nact * 4, cudaMemcpyDeviceToDevice, st));
        }
        ok = ok && c->sync();
        std::swap(aSlot[0].p, s2.p); std::swap(aSlot[0].bytes, s2.bytes); std::swap(aSlot[0].from, s2.from);
        std::swap(aGrp.p, g2.p); std::swap(aGrp.bytes, g2.bytes); std::swap(aGrp.from, g2.from);
    }
    bufK[other].release();
    bufV[other].release();
    if (where) bufK[1].release(); else bufK[0].release();            // the sorted keys are dead; sp (positions) stays: it is the SA slice
    // ISA: ranks of ALL suffixes to the position owners
    ok = ok && isa.alloc((m + 1) * 8);
    if (!sync_all_at(ok, __LINE__)) return -2;
    mark(2);
    if (!update_isa(sp, rankbuf.as<u64>(), M)) return -2;
    rankbuf.release();
    mark(3);
    {
        std::lock_guard<std::mutex> lk(g.stats_mutex);
        g.stats.rounds = 1; g.stats.slice_max = std::max<u64>(g.stats.slice_max, M); g.stats.active_after_round0 += nact;
    }

    // ---- doubling rounds on the unresolved suffixes
    const int rank_bits = bits_for(n);
    u64 h = (u64)k;
    int cur = 0;
    for (int round = 1;; ++round) {
        g.scal[(size_t)r * 8] = nact; g.scal[(size_t)r * 8 + 1] = ngrp;
        if (!sync_all_at(true, __LINE__)) return -2;
        u64 tot = 0, maxg = 0;
        for (int q = 0; q < G; ++q) { tot += g.scal[(size_t)q * 8]; maxg = std::max(maxg, g.scal[(size_t)q * 8 + 1]); }
        if (!sync_all_at(true, __LINE__)) return -2;                               // scal is rewritten below
        if (tot == 0) break;
        if (round > 80 || rank_bits + bits_for(maxg ? maxg : 1) > 63) { sync_all_at(false, __LINE__); return -2; }
        const u64 N = nact;
        // k2[j] = rank of suffix p_j + h, plus 1 (0 past the end): request / answer exchange with the position owners
        ok = k2.alloc((N + 1) * 8);
        if (!fetch_isa(ok ? aPos[cur].as<u64>() : nullptr, N, h, k2.as<u64>(), ok)) return -2;
        // keys (group, rank of the suffix h further + 1), local sort, rank stage
        ok = bufK[0].alloc((N + 1) * 8) && bufK[1].alloc((N + 1) * 8) && bufV[0].alloc((N + 1) * 8) && bufV[1].alloc((N + 1) * 8) && rankbuf.alloc((N + 1) * 8);
        u64 nact2 = 0, ngrp2 = 0;
        if (ok && N) {
            LSC_LAUNCH(*c, KC_ROUND_KEYS, (double)N * 20, d64_round_keys_kernel, grid_for(N), 256, 0, aGrp.as<u32>(), k2.as<u64>(), N, rank_bits, bufK[0].as<u64>());
            c->check(cudaMemcpyAsync(bufV[0].p, aPos[cur].p, N * 8, cudaMemcpyDeviceToDevice, st));
            c->reset_arena();
            void *temp = c->alloc(RadixSort<u64, u64>::temp_bytes(N));
            ok = temp != nullptr;
            int w = -1;
            if (ok) w = RadixSort<u64, u64>::sort(*c, bufK[0].as<u64>(), bufV[0].as<u64>(), bufK[1].as<u64>(), bufV[1].as<u64>(), N, 0, rank_bits + bits_for(maxg ? maxg : 1), temp, err);
            ok = ok && w >= 0;
            if (ok) {
                const u64 *kk = w ? bufK[1].as<u64>() : bufK[0].as<u64>();
                const u64 *pp = w ? bufV[1].as<u64>() : bufV[0].as<u64>();
                DevBuf g2;
                ok = g2.alloc((N + 1) * 4);
                ok = ok && rank_stage(kk, pp, aSlot[cur].as<u64>(), N, ~0ull, base, rankbuf.as<u64>(), sp, aPos[cur ^ 1].as<u64>(), aSlot[cur ^ 1].as<u64>(), g2.as<u32>(), &nact2, &ngrp2);
                if (ok && nact2) c->check(cudaMemcpyAsync(aGrp.p, g2.p, nact2 * 4, cudaMemcpyDeviceToDevice, st));
                ok = ok && c->sync();
                if (!sync_all_at(ok, __LINE__)) return -2;
                if (!update_isa(pp, rankbuf.as<u64>(), N)) return -2;
            } else if (!sync_all_at(false, __LINE__)) return -2;
        } else {
            if (!sync_all_at(ok, __LINE__)) return -2;
            if (!update_isa(nullptr, nullptr, 0)) return -2;
        }
        nact = nact2; ngrp = ngrp2; cur ^= 1; h *= 2;
        if (r == 0) { std::lock_guard<std::mutex> lk(g.stats_mutex); g.stats.rounds = round + 1; }
    }

    mark(4);
    if (r == 0) {
        std::lock_guard<std::mutex> lk(g.stats_mutex);
        g.stats.seconds_device = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    }
    if (g.want_verify && !verify(sp, M, base)) return -2;
    // ---- BWT rows / aux samples (libsais64_bwt[_aux] beyond the single-GPU limit): local, the packed text is replicated
    if (g.rows != nullptr) {
        DevBuf drows, dinv, dprim;
        ok = drows.alloc(M + 1) && dinv.alloc(256) && dprim.alloc(8);
        if (ok) {
            u8 *h_inv = (u8 *)(c->h_scalars + S_MISC);
            for (int s2 = 0; s2 < 256; ++s2) h_inv[s2] = 0;
            for (int s2 = 0; s2 < 256; ++s2) if (g.hist[s2]) h_inv[g.lut[s2]] = (u8)s2;
            c->check(cudaMemcpyAsync(dinv.p, h_inv, 256, cudaMemcpyHostToDevice, st));
            c->check(cudaMemsetAsync(dprim.p, 0, 8, st));
            if (M) {
                LSC_LAUNCH(*c, KC_BWT, (double)M * 10, d64_bwt_rows_kernel, grid_for(M), 256, 0, sp, M, base, words.as<u64>(), b, dinv.as<u8>(), drows.as<u8>(), dprim.as<u64>());
                c->check(cudaMemcpyAsync(g.rows + base, drows.p, M, cudaMemcpyDeviceToHost, st));
            }
            u64 pr = 0;
            c->check(cudaMemcpyAsync(&pr, dprim.p, 8, cudaMemcpyDeviceToHost, st));
            ok = c->sync() && !c->failed();
            if (ok && pr) g.primary.store(pr);
        }
        if (ok && g.aux_I != nullptr && g.aux_r && m) {
            const u64 first_idx = ceil_div(lo, g.aux_r), last_idx = (hi - 1) / g.aux_r;
            if (last_idx >= first_idx) {
                const u64 nidx = last_idx - first_idx + 1;
                DevBuf dI;
                ok = dI.alloc(nidx * 8);
                if (ok) {
                    LSC_LAUNCH(*c, KC_BWT, (double)nidx * 16, d64_aux_kernel, grid_for(nidx), 256, 0, isa.as<u64>(), lo, m, g.aux_r, first_idx, nidx, dI.as<i64>());
                    c->check(cudaMemcpyAsync(g.aux_I + first_idx, dI.p, nidx * 8, cudaMemcpyDeviceToHost, st));
                    ok = c->sync() && !c->failed();
                }
            }
        }
        if (!sync_all_at(ok, __LINE__)) return -2;
    }
    // ---- my slice of the suffix array -> the caller's array
    if (g.SA != nullptr && M) {
        c->check(cudaMemcpyAsync(g.SA + base, sp, M * 8, cudaMemcpyDeviceToHost, st));
        ok = c->sync();
    }
    {
        std::lock_guard<std::mutex> lk(g.stats_mutex);
        g.stats.exchanged_bytes += exchanged_bytes;
    }
    if (!sync_all_at(ok, __LINE__)) return -2;
    return 0;
}

// Suffix array of a HOST text over the given GPUs (one host thread per GPU).  SA may be NULL (timing / memory tests).
static int sa64_multi_impl(const u8 *T, i64 *SA, u64 n, i64 *freq, const int *devices, int ndev, void *stats_out,
                           u8 *rows, i64 *aux_I, u64 aux_r, u64 *primary_out)
{
    libsais_cuda_dist_stats *stats = (libsais_cuda_dist_stats *)stats_out;
    if (ndev < 1 || ndev > kD64MaxRanks - 1 || n < 2) return -1;
    DistGroup g;
    g.G = ndev; g.devs.assign(devices, devices + ndev);
    g.bar.n = ndev; g.n = n; g.T = T; g.SA = SA;
    g.rows = rows; g.aux_I = aux_I; g.aux_r = aux_r;
    g.B = ceil_div(ceil_div(n, (u64)ndev), 512) * 512;
    g.cnt.assign((size_t)ndev * (ndev + 1), 0); g.pk.assign(ndev, nullptr); g.pv.assign(ndev, nullptr); g.pwords.assign(ndev, nullptr); g.pans.assign(ndev, nullptr);
    g.scal.assign((size_t)ndev * 8, 0); g.hist.assign((size_t)ndev * 256, 0); g.samples.assign((size_t)ndev * kD64Samples, ~0ull);
    g.want_verify = stats != nullptr && stats->verify != 0;
    { const char *e = getenv("LIBSAIS_CUDA_DIST_VERIFY"); if (e && *e && atoi(e) != 0) g.want_verify = true; }
    std::memset(&g.stats, 0, sizeof(g.stats));
    std::vector<int> rc(ndev, 0);
    std::vector<std::thread> th;
    std::vector<std::unique_ptr<DistRank>> ranks;
    for (int r = 0; r < ndev; ++r) ranks.emplace_back(new DistRank(g, r));
    for (int r = 0; r < ndev; ++r) th.emplace_back([&, r] { rc[r] = ranks[r]->run(); });
    for (auto &t : th) t.join();
    int out = 0;
    for (int r = 0; r < ndev; ++r) if (rc[r] != 0) out = -2;
    if (out == 0 && freq != nullptr) for (int s = 0; s < 256; ++s) freq[s] = (i64)g.hist[s];
    if (stats) { *stats = g.stats; stats->n_gpus = ndev; stats->key_symbols = g.k; stats->key_bits = g.K; }
    if (primary_out) *primary_out = g.primary.load();
    return out;
}

int sa64_multi(const u8 *T, i64 *SA, u64 n, i64 *freq, const int *devices, int ndev, void *stats_out)
{
    return sa64_multi_impl(T, SA, n, freq, devices, ndev, stats_out, nullptr, nullptr, 0, nullptr);
}

// BWT (+ aux samples) of a host text over several GPUs: libsais64_bwt[_aux] beyond the single-GPU limit
// (reference src/libsais64.c:7133, :7172).  A (the caller's temporary int64[n]) holds the n per-slot rows while the ranks work;
// U may alias T (every rank has consumed its slice of T before anything is written).  Returns the primary index or < 0.
i64 bwt64_multi(const u8 *T, u8 *U, i64 *A, u64 n, i64 *freq, u64 aux_r, i64 *aux_I, const int *devices, int ndev)
{
    u8 *rows = (u8 *)A;
    const u8 last = T[n - 1];
    u64 primary = 0;
    const int rc = sa64_multi_impl(T, nullptr, n, freq, devices, ndev, nullptr, rows, aux_I, aux_r, &primary);
    if (rc != 0) return rc;
    if (primary < 1 || primary > n) return -2;
    // U[0] = T[n-1]; slot i < p0 -> U[i+1]; slot i > p0 -> U[i]  (the row of the virtual sentinel is dropped; p0 = primary - 1)
    const u64 p0 = primary - 1;
    U[0] = last;
    std::memmove(U + 1, rows, p0);
    std::memmove(U + p0 + 1, rows + p0 + 1, n - 1 - p0);
    return (i64)primary;
}

}  // namespace lsc
