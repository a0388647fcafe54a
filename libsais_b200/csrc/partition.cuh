// partition.cuh -- UNSTABLE one-pass digit partition of (key, value) pairs, and the MSD round-0 pipeline
// built on it (hist16 -> partition by the key's top 8 bits -> segmented partition by the next 8 bits ->
// bucket_sort: every 16-bit bucket finished inside shared memory).
//
// The stable onesweep pass of radix_sort.cuh spends most of its instructions on the warp-level multisplit
// that keeps equal digits in input order (8 ballots + leader atomics per item: 136 instructions per pair,
// issue-bound at 0.42 of the HBM roofline).  A partition needs no order inside a digit: one shared-memory
// atomicAdd per element hands out the rank inside the tile, which leaves the load / look-back / coalesced
// write-out skeleton -- about a third of the instructions.  Users: the round-0 MSD path below (ties are
// broken explicitly by position in bucket_sort_kernel, so the result is the same array the stable LSD sort
// produces) and the locality partition of scatter.cuh (ISA / phi scatters), where order never mattered.
//
// Replaces, with radix_sort.cuh, the reference's induced-sorting scans (src/libsais.c:2157-4101).
#pragma once
#include "radix_sort.cuh"

namespace lsc {

// ---- element sources of a partition pass
// SRC_ARRAYS : (kin, vin) arrays; full, 16-byte aligned tiles come in by TMA bulk copies
// SRC_FUNC   : a functor produces element i (key(i), val(i)): the pair array is never materialised
// SRC_KMER   : round 0 -- element i <-> text position p = n-1-i, key = K-bit k-mer of suffix p (<< key_shift,
//              | preceding text byte in BWT mode), value = p.  The tile's slice of the packed text (and of the
//              raw text, for the preceding bytes) is staged in shared memory with coalesced loads first.
enum { SRC_ARRAYS = 0, SRC_FUNC = 1, SRC_KMER = 2 };

struct ArraySrc {
    static const int kMode = SRC_ARRAYS;
    __device__ __forceinline__ u64 key(u64) const { return 0; }
    __device__ __forceinline__ u32 val(u64) const { return 0; }
};
template <typename F> struct FuncSrc {
    static const int kMode = SRC_FUNC;
    F f;
    __device__ __forceinline__ u64 key(u64 i) const { return f.key(i); }
    __device__ __forceinline__ u32 val(u64 i) const { return f.val(i); }
};
struct KmerSrc {
    static const int kMode = SRC_KMER;
    const u64 *words; u64 nwords; const u8 *text; u64 n; int b, K, key_shift;
    __device__ __forceinline__ u64 key(u64) const { return 0; }
    __device__ __forceinline__ u32 val(u64) const { return 0; }
};

template <typename KeyT, typename ValT, int THREADS, int IPT>
struct PartSmem {
    static const int TILE = THREADS * IPT;
    alignas(16) KeyT keys[TILE];           // TMA / generator staging, then the tile in digit order
    u64  goff[kRadixSize];                 // directly behind keys[]: the k-mer staging may spill a few words into it
    alignas(16) ValT vals[TILE];
    alignas(16) ValT vals_in[TILE];        // TMA staging of the values; SRC_KMER: the tile's raw text bytes
    u32  cnt[kRadixSize];
    u32  tileoff[kRadixSize];
    u32  scan_tmp[32];
    alignas(8) u64 mbar[2];
    u32  tile;
};

// Tickets are INTERLEAVED over independent segments: ticket t works on segment t % nseg, tile t / nseg of that
// segment, and the chained scan of a tile only spans its own segment.  The predecessor of a tile was therefore
// issued nseg tickets earlier -- with nseg >= the number of resident CTAs it has long finished, its INCLUSIVE
// prefix is published, and the look-back ends after one (prefetched) status word instead of ~22 tiles and
// several L2 round trips with spinning (profiles/part_pass_r2.md).
//   segmented mode (second MSD level): segment = top-level bucket c = [boff[c << 8], boff[(c + 1) << 8]); a tile is
//     the intersection of a TILE-aligned cell of the array with one bucket, so interior tiles stay full and
//     16-byte aligned (TMA).  Elements go to boff[c * 256 + d] + (digit-d elements of earlier tiles of c) + rank.
//   chunked mode (first level): segment = chunk of tpc consecutive tiles; cp[s][d] = digit-d elements in earlier
//     chunks (from the per-chunk histograms hist16 takes anyway).  nseg == 1: plain chained scan over all tiles.
struct PartArgs {
    u64 n; int shift; u32 dmask;
    const u64 *base;        // [256] global digit bases (not segmented)
    const u32 *cp;          // [nseg][256] chunk prefixes, or null
    u32 nseg, tpc;
    const u32 *boff;        // segmented: [65537] offsets of the 16-bit buckets
    const u32 *tstart;      // segmented: [257] exclusive scan of the tiles per top-level bucket
    const uint2 *tinfo;     // segmented: per ticket (first element, count) from seg_tiles_kernel; count 0 = unused ticket
    u32 *ticket;            // null: the ticket is blockIdx.x (CTAs are dispatched in index order, so a tile's predecessors are
                            // always resident or done -- the same assumption CUB's decoupled look-back scan makes); else atomic
    u32 *err; int use_bulk;
    void *const *kptr;      // non-null: per-digit output bases (device array of 256 pointers each), possibly in PEER memory:
    void *const *vptr;      //   element i of digit d goes to kptr[d][i] / vptr[d][i] -- the fused route + exchange of dist64.cu
    int out32 = 0;          // part_pipe_kernel only: write the low 32 bits of the 64-bit keys (second MSD level when the key bits
                            // below the 16-bit bucket prefix fit a word: the bucket is implied by where the element lands)
};

template <typename KeyT, typename ValT, int THREADS, int IPT, int MINB, typename ST, typename Src, bool SEG>
__global__ void __launch_bounds__(THREADS, MINB)
part_pass_kernel(const KeyT *__restrict__ kin, const ValT *__restrict__ vin,
                 KeyT *__restrict__ kout, ValT *__restrict__ vout, const PartArgs a, ST *status, const Src src)
{
    typedef PartSmem<KeyT, ValT, THREADS, IPT> Smem;
    constexpr int TILE = THREADS * IPT;
    static_assert(THREADS >= kRadixSize && THREADS % 32 == 0, "one thread per digit is assumed");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = a.shift; const u32 dmask = a.dmask;

    uint2 ti = make_uint2(0, 0);
    if (SEG && a.ticket == nullptr) ti = a.tinfo[blockIdx.x];        // issued before the barrier: one round trip saved
    if (tid == 0) {
        sm.tile = a.ticket != nullptr ? atomicAdd(a.ticket, 1u) : blockIdx.x;
        mbar_init(&sm.mbar[0], 1); mbar_init(&sm.mbar[1], 1);
        mbar_fence_init();
    }
    if (tid < kRadixSize) sm.cnt[tid] = 0;
    __syncthreads();
    const u32 ticket = a.ticket != nullptr ? sm.tile : blockIdx.x;
    const u32 segi = ticket % a.nseg, j = ticket / a.nseg;           // segment, tile inside the segment

    // ---- which elements: [lo, lo + count) of the input
    u64 lo; u32 count;
    if (SEG) {
        if (a.ticket != nullptr) ti = a.tinfo[ticket];
        if (ti.y == 0) return;                                        // grid = nseg * (most tiles of any bucket)
        lo = ti.x; count = ti.y;
    } else {
        const u64 tile = (u64)segi * a.tpc + j;
        lo = tile * TILE;
        if (j >= a.tpc || lo >= a.n) return;
        count = (u32)((a.n - lo) < (u64)TILE ? (a.n - lo) : (u64)TILE);
    }
    const bool full = count == (u32)TILE;

    // ---- load: element li of the tile is held by thread (li & 31) + 32 * warp-slot, warp-striped
    KeyT key[IPT];
    const u32 wbase = warp * (IPT * 32) + lane;
    const bool bulk = Src::kMode == SRC_ARRAYS && full && a.use_bulk && ((lo * sizeof(ValT)) & 15) == 0 && ((lo * sizeof(KeyT)) & 15) == 0;
    u64 gbase = 0;
    if (tid < kRadixSize) {
        if (SEG) gbase = (u64)a.boff[(segi << 8) + tid];
        else if (a.kptr == nullptr) { gbase = a.base[tid]; if (a.cp != nullptr) gbase += (u64)a.cp[(u64)segi * kRadixSize + tid]; }
    }
    if constexpr (Src::kMode == SRC_KMER) {
        // positions of the tile: p_hi down to p_lo; words [w_lo, w_lo + nw) cover bits [p_lo*b, p_hi*b + 64 + 63]
        const u64 p_hi = src.n - 1 - lo, p_lo = p_hi - (count - 1);
        const u64 w_lo = (p_lo * (u64)src.b) >> 6;
        u64 w_hi = ((p_hi * (u64)src.b) >> 6) + 1;
        if (w_hi > src.nwords - 1) w_hi = src.nwords - 1;
        const u32 nw = (u32)(w_hi - w_lo + 1);
        u64 *sw = reinterpret_cast<u64 *>(sm.keys);
        for (u32 i = tid; i < nw; i += THREADS) sw[i] = src.words[w_lo + i];
        // raw text bytes [p_lo - 1, p_hi - 1] as aligned 32-bit words
        u8 *sb = reinterpret_cast<u8 *>(sm.vals_in);
        u64 t0 = 0;
        if (src.text != nullptr) {
            const u64 first_b = p_lo ? p_lo - 1 : 0;
            const uintptr_t a0 = ((uintptr_t)(src.text + first_b)) & ~(uintptr_t)3;
            t0 = (u64)(a0 - (uintptr_t)src.text);                   // may wrap "below" the text by < 4 bytes: a0 is still inside the allocation's alignment
            const u32 nq = (u32)((p_hi + 3 - t0) >> 2);             // aligned 32-bit words that hold bytes [t0, p_hi - 1]
            const u32 *src32 = reinterpret_cast<const u32 *>(a0);
            u32 *sb32 = reinterpret_cast<u32 *>(sb);
            for (u32 i = tid; i < nq && i < (u32)TILE; i += THREADS) sb32[i] = src32[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const u32 li = wbase + i * 32;
            u64 k = 0;
            if (full || li < count) {
                const u64 p = p_hi - li;
                const u64 bit = p * (u64)src.b;
                const u32 q = (u32)((bit >> 6) - w_lo); const int off = (int)(bit & 63);
                const u64 hi = sw[q], lw = sw[q + 1 < nw ? q + 1 : q];
                const u64 x = off ? ((hi << off) | (lw >> (64 - off))) : hi;
                k = (x >> (64 - src.K)) << src.key_shift;
                if (src.text != nullptr && p > 0) k |= (u64)sb[(p - 1) - t0];
            }
            key[i] = (KeyT)k;
        }
    } else if (bulk) {
        if (tid == 0) {
            mbar_expect_tx(&sm.mbar[0], (u32)(TILE * sizeof(KeyT)));
            bulk_load(sm.keys, kin + lo, (u32)(TILE * sizeof(KeyT)), &sm.mbar[0]);
            mbar_expect_tx(&sm.mbar[1], (u32)(TILE * sizeof(ValT)));
            bulk_load(sm.vals_in, vin + lo, (u32)(TILE * sizeof(ValT)), &sm.mbar[1]);
        }
        mbar_wait(&sm.mbar[0], 0);
#pragma unroll
        for (int i = 0; i < IPT; ++i) key[i] = sm.keys[wbase + i * 32];
    } else {
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const u32 li = wbase + i * 32;
            KeyT k = (KeyT)0;
            if (full || li < count) {
                if constexpr (Src::kMode == SRC_FUNC) k = (KeyT)src.key(lo + li); else k = kin[lo + li];
            }
            key[i] = k;
        }
        if constexpr (Src::kMode == SRC_ARRAYS) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const u32 li = wbase + i * 32;
                if (full || li < count) cp_async_val(&sm.vals_in[li], vin + lo + li);
            }
        }
    }

    // ---- rank inside the tile: ONE shared atomic per element (no order inside a digit)
    u32 rk[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const u32 li = wbase + i * 32;
        rk[i] = 0;
        if (full || li < count) rk[i] = atomicAdd(&sm.cnt[digit_of(key[i], shift, dmask)], 1u);
    }
    __syncthreads();

    // ---- per digit: publish the tile count, prefetch the look-back, exclusive scan of the counts
    u32 cnt = 0, tileoff = 0;
    LookState<ST> ls;
    const u32 nseg = a.nseg;
    if (tid < kRadixSize) {
        cnt = sm.cnt[tid];
        st_relaxed(status + (u64)ticket * kRadixSize + tid, j == 0 ? StWord<ST>::inc(cnt) : StWord<ST>::agg(cnt));
#pragma unroll
        for (int q = 0; q < kLookBatch; ++q)
            ls.w[q] = (u32)(q + 1) <= j ? ld_relaxed(status + (u64)(ticket - (u32)(q + 1) * nseg) * kRadixSize + tid) : StWord<ST>::inc(0);
        u32 x = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
        if (lane == 31) sm.scan_tmp[warp] = x;
        tileoff = x - cnt;
    }
    __syncthreads();
    if (tid < kRadixSize) {
#pragma unroll
        for (int w = 0; w < kRadixSize / 32; ++w) if (w < warp) tileoff += sm.scan_tmp[w];
        sm.tileoff[tid] = tileoff;
    }
    __syncthreads();

    // ---- scatter into shared memory in digit order
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const u32 li = wbase + i * 32;
        if (full || li < count) {
            const u32 pos = sm.tileoff[digit_of(key[i], shift, dmask)] + rk[i];
            sm.keys[pos] = key[i];
            if constexpr (Src::kMode == SRC_KMER) sm.vals[pos] = (ValT)(src.n - 1 - lo - li);
            else if constexpr (Src::kMode == SRC_FUNC) sm.vals[pos] = (ValT)src.val(lo + li);
            else rk[i] = pos;
        }
    }
    if constexpr (Src::kMode == SRC_ARRAYS) {
        if (bulk) mbar_wait(&sm.mbar[1], 0); else cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const u32 li = wbase + i * 32;
            if (full || li < count) sm.vals[rk[i]] = sm.vals_in[li];
        }
    }

    // ---- chained scan over the earlier tiles of this segment (tickets ticket - nseg, ticket - 2 nseg, ...)
    if (tid < kRadixSize) {
        u64 excl = 0;
        if (j != 0) {
            bool done = false;
#pragma unroll
            for (int q = 0; q < kLookBatch; ++q)
                if (!done) done = lookback_consume<ST>(ls.w[q], status, (i64)ticket - (i64)(q + 1) * nseg, (u32)tid, excl, a.err);
            u32 back = kLookBatch;
            while (!done) {
                ST w[kLookRefill];
#pragma unroll
                for (int q = 0; q < kLookRefill; ++q) {
                    const u32 k = back + 1 + q;
                    w[q] = k <= j ? ld_relaxed(status + (u64)(ticket - k * nseg) * kRadixSize + tid) : StWord<ST>::inc(0);
                }
#pragma unroll
                for (int q = 0; q < kLookRefill; ++q)
                    if (!done) done = lookback_consume<ST>(w[q], status, (i64)ticket - (i64)(back + 1 + q) * nseg, (u32)tid, excl, a.err);
                back += kLookRefill;
            }
            st_relaxed(status + (u64)ticket * kRadixSize + tid, StWord<ST>::inc(excl + (u64)cnt));
        }
        sm.goff[tid] = gbase + excl - (u64)tileoff;
    }
    __syncthreads();

    // ---- write out: consecutive threads own consecutive slots of a digit's run
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const u32 idx = i * THREADS + tid;
        if (full || idx < count) {
            const KeyT k = sm.keys[idx];
            const u32 d = digit_of(k, shift, dmask);
            const u64 g = sm.goff[d] + idx;
            if (a.kptr != nullptr) {
                reinterpret_cast<KeyT *>(a.kptr[d])[g] = k;
                reinterpret_cast<ValT *>(a.vptr[d])[g] = sm.vals[idx];
            } else {
                kout[g] = k;
                vout[g] = sm.vals[idx];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent, double-buffered variant of the partition pass (round-0 MSD levels; u64 keys, u32 values).
// profiles/part_pass_r2.md: in the one-tile-per-CTA kernel ~40 % of all stall samples sit in the start-up chain of a tile
// (ticket -> tile table -> TMA issue -> DRAM latency), with only 2-3 CTAs per SM to hide it.  Here a CTA walks tickets
// blockIdx.x, + gridDim.x, ... and, while it ranks / scatters / writes out tile k, the loads of tile k+1 are already in
// flight into the other stage: TMA bulk copies completing on per-stage mbarriers (array source), cp.async of the tile's
// slice of the packed text and of the raw text (k-mer source).  The tile table entry is fetched two tiles ahead.
// All CTAs are resident (grid <= occupancy), so the chained scan cannot deadlock.
// ---------------------------------------------------------------------------------------------
template <int THREADS, int IPT, int MODE>
struct PipeSmem {
    static const int TILE = THREADS * IPT;
    static const int WMAX = TILE / 8 + 8;                       // k-mer source: words of one tile at <= 8 bits per symbol
    static const int TMAX = TILE / 4 + 8;                       // raw text of one tile as 32-bit words
    alignas(16) u64 kst[MODE == SRC_ARRAYS ? 2 : 1][TILE];      // TMA staging per stage; the current stage then holds the tile in digit order
    alignas(16) u32 vst[MODE == SRC_ARRAYS ? 2 : 1][MODE == SRC_ARRAYS ? TILE : 4];
    alignas(16) u32 vals[TILE];
    alignas(16) u64 wst[MODE == SRC_KMER ? 2 : 1][MODE == SRC_KMER ? WMAX : 2];
    alignas(16) u32 tst[MODE == SRC_KMER ? 2 : 1][MODE == SRC_KMER ? TMAX : 2];
    u64  goff[kRadixSize];
    u32  cnt[kRadixSize];
    u32  tileoff[kRadixSize];
    u32  scan_tmp[32];
    alignas(8) u64 mbar[2][2];                                  // [stage][keys / values]
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int THREADS, int IPT, int MINB, typename ST, typename Src, bool SEG>
__global__ void __launch_bounds__(THREADS, MINB)
part_pipe_kernel(const u64 *__restrict__ kin, const u32 *__restrict__ vin, u64 *__restrict__ kout, u32 *__restrict__ vout,
                 const PartArgs a, ST *status, const Src src, const u32 ntickets)
{
    typedef PipeSmem<THREADS, IPT, Src::kMode> Smem;
    constexpr int TILE = THREADS * IPT;
    static_assert(THREADS >= kRadixSize && THREADS % 32 == 0, "one thread per digit is assumed");
    static_assert(Src::kMode == SRC_ARRAYS || Src::kMode == SRC_KMER, "array or k-mer source");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = a.shift; const u32 dmask = a.dmask;
    const u32 stride = gridDim.x;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) { mbar_init(&sm.mbar[s][0], 1); mbar_init(&sm.mbar[s][1], 1); }
        mbar_fence_init();
    }
    __syncthreads();

    // range of a ticket: (first element, count); count 0 = nothing to do
    auto range_of = [&](u32 t, uint2 ti, u64 &lo, u32 &count) {
        if (SEG) { lo = ti.x; count = ti.y; return; }
        const u32 segi = t % a.nseg, j = t / a.nseg;
        const u64 tile = (u64)segi * a.tpc + j;
        lo = tile * TILE; count = 0;
        if (j < a.tpc && lo < a.n) count = (u32)((a.n - lo) < (u64)TILE ? (a.n - lo) : (u64)TILE);
    };
    // issue the loads of a tile into stage s.  Array source: two TMA bulk copies (full, aligned tiles only; other tiles are
    // loaded synchronously when they are processed).  K-mer source: cp.async of the words / text slices, one commit group.
    auto issue = [&](int s, u64 lo, u32 count) {
        if constexpr (Src::kMode == SRC_ARRAYS) {
            const bool bulk = count == (u32)TILE && a.use_bulk && ((lo * 4) & 15) == 0;
            if (bulk && tid == 0) {
                fence_proxy_async();                                   // the stage was last touched through the generic proxy
                mbar_expect_tx(&sm.mbar[s][0], (u32)(TILE * 8));
                bulk_load(sm.kst[s], kin + lo, (u32)(TILE * 8), &sm.mbar[s][0]);
                mbar_expect_tx(&sm.mbar[s][1], (u32)(TILE * 4));
                bulk_load(sm.vst[s], vin + lo, (u32)(TILE * 4), &sm.mbar[s][1]);
            }
        } else {
            if (count) {
                const u64 p_hi = src.n - 1 - lo, p_lo = p_hi - (count - 1);
                const u64 w_lo = (p_lo * (u64)src.b) >> 6;
                u64 w_hi = ((p_hi * (u64)src.b) >> 6) + 1;
                if (w_hi > src.nwords - 1) w_hi = src.nwords - 1;
                const u32 nw = (u32)(w_hi - w_lo + 1);
                for (u32 i = tid; i < nw && i < (u32)Smem::WMAX; i += THREADS) cp_async8(&sm.wst[s][i], src.words + w_lo + i);
                if (src.text != nullptr) {
                    const u64 first_b = p_lo ? p_lo - 1 : 0;
                    const uintptr_t a0 = ((uintptr_t)(src.text + first_b)) & ~(uintptr_t)3;
                    const u64 t0 = (u64)(a0 - (uintptr_t)src.text);
                    const u32 nq = (u32)((p_hi + 3 - t0) >> 2);
                    const u32 *src32 = reinterpret_cast<const u32 *>(a0);
                    for (u32 i = tid; i < nq && i < (u32)Smem::TMAX; i += THREADS) cp_async4(&sm.tst[s][i], src32 + i);
                }
            }
            cp_async_commit();                                         // one group per tile, also when empty: wait_group counts groups
        }
    };

    // ---- prologue: tile 0 of this CTA in flight, table entry of tile 1 fetched
    u32 t = blockIdx.x;
    uint2 ti_cur = make_uint2(0, 0), ti_next = make_uint2(0, 0);
    if (SEG) {
        if (t < ntickets) ti_cur = a.tinfo[t];
        if (t + stride < ntickets && t + stride >= t) ti_next = a.tinfo[t + stride];
    }
    u64 lo = 0; u32 count = 0;
    if (t < ntickets) { range_of(t, ti_cur, lo, count); issue(0, lo, count); }
    u32 waited0 = 0, waited1 = 0;                                  // completed TMA phases per stage (only full, aligned tiles use TMA)

    for (u32 it = 0; t < ntickets; ++it) {
        const int s = (int)(it & 1);
        const u32 par = (s ? waited1 : waited0) & 1;
        // ---- next tile: its loads go into the other stage now; the table entry after it is fetched for the next iteration
        const u32 tn = t + stride;
        const bool have_next = tn < ntickets && tn > t;
        u64 lo_n = 0; u32 count_n = 0;
        uint2 ti_next2 = make_uint2(0, 0);
        if (have_next) {
            range_of(tn, ti_next, lo_n, count_n);
            if (SEG && tn + stride < ntickets && tn + stride > tn) ti_next2 = a.tinfo[tn + stride];
        }
        if (have_next) issue(s ^ 1, lo_n, count_n);
        else if (Src::kMode == SRC_KMER) cp_async_commit();            // keep one group per iteration

        if (count != 0) {
            const u32 segi = t % a.nseg, j = t / a.nseg;
            const bool full = count == (u32)TILE;
            if (tid < kRadixSize) sm.cnt[tid] = 0;
            u64 gbase = 0;
            if (tid < kRadixSize) {
                if (SEG) gbase = (u64)a.boff[(segi << 8) + tid];
                else { gbase = a.base[tid]; if (a.cp != nullptr) gbase += (u64)a.cp[(u64)segi * kRadixSize + tid]; }
            }
            u64 key[IPT];
            const u32 wbase = warp * (IPT * 32) + lane;
            bool bulk = false;
            if constexpr (Src::kMode == SRC_KMER) {
                cp_async_wait_group<1>();                              // everything but the newest group (the next tile) has landed
                __syncthreads();
                const u64 p_hi = src.n - 1 - lo, p_lo = p_hi - (count - 1);
                const u64 w_lo = (p_lo * (u64)src.b) >> 6;
                u64 w_hi = ((p_hi * (u64)src.b) >> 6) + 1;
                if (w_hi > src.nwords - 1) w_hi = src.nwords - 1;
                const u32 nw = (u32)(w_hi - w_lo + 1);
                const u64 *sw = sm.wst[s];
                const u8 *sb = reinterpret_cast<const u8 *>(sm.tst[s]);
                u64 t0 = 0;
                if (src.text != nullptr) {
                    const u64 first_b = p_lo ? p_lo - 1 : 0;
                    const uintptr_t a0 = ((uintptr_t)(src.text + first_b)) & ~(uintptr_t)3;
                    t0 = (u64)(a0 - (uintptr_t)src.text);
                }
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    const u32 li = wbase + i * 32;
                    u64 k = 0;
                    if (full || li < count) {
                        const u64 p = p_hi - li;
                        const u64 bit = p * (u64)src.b;
                        const u32 q = (u32)((bit >> 6) - w_lo); const int off = (int)(bit & 63);
                        const u64 hi = sw[q], lw = sw[q + 1 < nw ? q + 1 : q];
                        const u64 x = off ? ((hi << off) | (lw >> (64 - off))) : hi;
                        k = (x >> (64 - src.K)) << src.key_shift;
                        if (src.text != nullptr && p > 0) k |= (u64)sb[(p - 1) - t0];
                    }
                    key[i] = k;
                }
            } else {
                bulk = full && a.use_bulk && ((lo * 4) & 15) == 0;
                __syncthreads();                                       // cnt zeroed; previous tile's readers of this stage are long gone
                if (bulk) {
                    mbar_wait(&sm.mbar[s][0], par);
#pragma unroll
                    for (int i = 0; i < IPT; ++i) key[i] = sm.kst[s][wbase + i * 32];
                } else {
#pragma unroll
                    for (int i = 0; i < IPT; ++i) {
                        const u32 li = wbase + i * 32;
                        key[i] = (full || li < count) ? kin[lo + li] : 0;
                    }
#pragma unroll
                    for (int i = 0; i < IPT; ++i) {
                        const u32 li = wbase + i * 32;
                        if (full || li < count) sm.vst[s][li] = vin[lo + li];
                    }
                }
            }

            // ---- rank inside the tile: one shared atomic per element
            u32 rk[IPT];
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const u32 li = wbase + i * 32;
                rk[i] = 0;
                if (full || li < count) rk[i] = atomicAdd(&sm.cnt[digit_of(key[i], shift, dmask)], 1u);
            }
            __syncthreads();

            // ---- per digit: publish the tile count, prefetch the look-back, exclusive scan of the counts
            u32 cnt = 0, tileoff = 0;
            LookState<ST> ls;
            const u32 nseg = a.nseg;
            if (tid < kRadixSize) {
                cnt = sm.cnt[tid];
                st_relaxed(status + (u64)t * kRadixSize + tid, j == 0 ? StWord<ST>::inc(cnt) : StWord<ST>::agg(cnt));
#pragma unroll
                for (int q = 0; q < kLookBatch; ++q)
                    ls.w[q] = (u32)(q + 1) <= j ? ld_relaxed(status + (u64)(t - (u32)(q + 1) * nseg) * kRadixSize + tid) : StWord<ST>::inc(0);
                u32 x = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
                if (lane == 31) sm.scan_tmp[warp] = x;
                tileoff = x - cnt;
            }
            __syncthreads();
            if (tid < kRadixSize) {
#pragma unroll
                for (int w = 0; w < kRadixSize / 32; ++w) if (w < warp) tileoff += sm.scan_tmp[w];
                sm.tileoff[tid] = tileoff;
            }
            __syncthreads();

            // ---- scatter into shared memory in digit order (the current stage's key area becomes the ordered tile)
            u64 *skeys = sm.kst[Src::kMode == SRC_ARRAYS ? s : 0];
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const u32 li = wbase + i * 32;
                if (full || li < count) {
                    const u32 pos = sm.tileoff[digit_of(key[i], shift, dmask)] + rk[i];
                    skeys[pos] = key[i];
                    if constexpr (Src::kMode == SRC_KMER) sm.vals[pos] = (u32)(src.n - 1 - lo - li);
                    else rk[i] = pos;
                }
            }
            if constexpr (Src::kMode == SRC_ARRAYS) {
                if (bulk) { mbar_wait(&sm.mbar[s][1], par); if (s) ++waited1; else ++waited0; }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < IPT; ++i) {
                    const u32 li = wbase + i * 32;
                    if (full || li < count) sm.vals[rk[i]] = sm.vst[s][li];
                }
            }

            // ---- chained scan over the earlier tiles of this segment
            if (tid < kRadixSize) {
                u64 excl = 0;
                if (j != 0) {
                    bool done = false;
#pragma unroll
                    for (int q = 0; q < kLookBatch; ++q)
                        if (!done) done = lookback_consume<ST>(ls.w[q], status, (i64)t - (i64)(q + 1) * nseg, (u32)tid, excl, a.err);
                    u32 back = kLookBatch;
                    while (!done) {
                        ST w[kLookRefill];
#pragma unroll
                        for (int q = 0; q < kLookRefill; ++q) {
                            const u32 k = back + 1 + q;
                            w[q] = k <= j ? ld_relaxed(status + (u64)(t - k * nseg) * kRadixSize + tid) : StWord<ST>::inc(0);
                        }
#pragma unroll
                        for (int q = 0; q < kLookRefill; ++q)
                            if (!done) done = lookback_consume<ST>(w[q], status, (i64)t - (i64)(back + 1 + q) * nseg, (u32)tid, excl, a.err);
                        back += kLookRefill;
                    }
                    st_relaxed(status + (u64)t * kRadixSize + tid, StWord<ST>::inc(excl + (u64)cnt));
                }
                sm.goff[tid] = gbase + excl - (u64)tileoff;
            }
            __syncthreads();

            // ---- write out
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                const u32 idx = i * THREADS + tid;
                if (full || idx < count) {
                    const u64 k = skeys[idx];
                    const u64 g = sm.goff[digit_of(k, shift, dmask)] + idx;
                    if (a.out32) reinterpret_cast<u32 *>(kout)[g] = (u32)k; else kout[g] = k;
                    vout[g] = sm.vals[idx];
                }
            }
        } else if (Src::kMode == SRC_KMER) {
            cp_async_wait_group<1>();
        }
        __syncthreads();                                               // the stage, cnt, goff, vals are free for the next tile
        t = tn; lo = lo_n; count = count_n; ti_next = ti_next2;
        if (!have_next) break;
    }
    if (Src::kMode == SRC_KMER) cp_async_wait_group<0>();
}

// Ticket table of the segmented pass: ticket t = tile t / 256 of top-level bucket t % 256 -> (first element, count).
static __global__ void __launch_bounds__(256)
seg_tiles_kernel(const u32 *__restrict__ boff, const u32 *__restrict__ tstart, u32 tile, u32 grid, uint2 *__restrict__ tinfo)
{
    const u32 t = blockIdx.x * 256 + threadIdx.x;
    if (t >= grid) return;
    const u32 segi = t % kRadixSize, j = t / kRadixSize;
    uint2 r = make_uint2(0, 0);
    if (j < tstart[segi + 1] - tstart[segi]) {
        const u64 blo = boff[segi << 8], bhi = boff[(segi + 1) << 8];
        const u64 cell = blo / tile + j;
        const u64 clo = cell * tile, chi = clo + tile;
        const u64 lo = blo > clo ? blo : clo;
        r = make_uint2((u32)lo, (u32)((bhi < chi ? bhi : chi) - lo));
    }
    tinfo[t] = r;
}

// Tile shapes of the partition pass; LIBSAIS_CUDA_PART_VARIANT selects one (profiles/part_pass_r2.md).
struct PartVariant { int threads, ipt, minb; };
static const PartVariant kPartVariants[] = { {384, 12, 2}, {384, 10, 3}, {512, 8, 2}, {256, 12, 4} };
static const int kNumPartVariants = sizeof(kPartVariants) / sizeof(kPartVariants[0]);
static inline int part_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LIBSAIS_CUDA_PART_VARIANT");
        v = (e && *e) ? atoi(e) : 1;                                    // 384 x 10, three CTAs per SM: fastest on B200 (profiles/part_pass_r2.md)
        if (v < 0 || v >= kNumPartVariants) v = 1;
    }
    return v;
}
static inline u32 part_tile() { const PartVariant &pv = kPartVariants[part_variant()]; return (u32)(pv.threads * pv.ipt); }

template <typename KeyT, typename ValT, int THREADS, int IPT, int MINB, typename Src, bool SEG>
static void launch_part_variant(Ctx &c, int kc, double algo_bytes, const Src &src, const KeyT *kin, const ValT *vin, KeyT *kout, ValT *vout,
                                PartArgs a, u64 grid, void *status)
{
    typedef PartSmem<KeyT, ValT, THREADS, IPT> Smem;
    static const bool bulk_env = [] { const char *e = getenv("LIBSAIS_CUDA_TMA"); return !(e && *e && atoi(e) == 0); }();
    a.use_bulk = bulk_env && Src::kMode == SRC_ARRAYS && (((uintptr_t)kin | (uintptr_t)vin) & 15) == 0
                 && ((size_t)THREADS * IPT * sizeof(KeyT)) % 16 == 0 && ((size_t)THREADS * IPT * sizeof(ValT)) % 16 == 0;
    if (a.n < (1ull << 30)) {
        auto kern = part_pass_kernel<KeyT, ValT, THREADS, IPT, MINB, u32, Src, SEG>;
        c.check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        LSC_LAUNCH(c, kc, algo_bytes, kern, (u32)grid, THREADS, sizeof(Smem), kin, vin, kout, vout, a, (u32 *)status, src);
    } else {
        auto kern = part_pass_kernel<KeyT, ValT, THREADS, IPT, MINB, u64, Src, SEG>;
        c.check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        LSC_LAUNCH(c, kc, algo_bytes, kern, (u32)grid, THREADS, sizeof(Smem), kin, vin, kout, vout, a, (u64 *)status, src);
    }
}

// Persistent launch: grid = resident CTAs (occupancy * SMs), tickets 0..ntickets-1 are walked with a grid stride.
template <typename ST, typename Src, bool SEG>
static void launch_part_pipe_st(Ctx &c, int kc, double algo_bytes, const Src &src, const u64 *kin, const u32 *vin, u64 *kout, u32 *vout,
                                const PartArgs &a, u64 ntickets, void *status)
{
    constexpr int THREADS = 384, IPT = 10;
    constexpr int MINB = Src::kMode == SRC_KMER ? 3 : 2;
    typedef PipeSmem<THREADS, IPT, Src::kMode> Smem;
    auto kern = part_pipe_kernel<THREADS, IPT, MINB, ST, Src, SEG>;
    c.check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, sizeof(Smem)) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
    u64 grid = (u64)per_sm * c.sm_count;
    if (grid > ntickets) grid = ntickets;
    LSC_LAUNCH(c, kc, algo_bytes, kern, (u32)grid, THREADS, sizeof(Smem), kin, vin, kout, vout, a, (ST *)status, src, (u32)ntickets);
}
template <typename Src, bool SEG>
static void launch_part_pipe(Ctx &c, int kc, double algo_bytes, const Src &src, const u64 *kin, const u32 *vin, u64 *kout, u32 *vout,
                             PartArgs a, u64 ntickets, void *status)
{
    static const bool bulk_env = [] { const char *e = getenv("LIBSAIS_CUDA_TMA"); return !(e && *e && atoi(e) == 0); }();
    a.use_bulk = bulk_env && Src::kMode == SRC_ARRAYS && (((uintptr_t)kin | (uintptr_t)vin) & 15) == 0;
    a.ticket = nullptr;
    if (a.n < (1ull << 30)) launch_part_pipe_st<u32, Src, SEG>(c, kc, algo_bytes, src, kin, vin, kout, vout, a, ntickets, status);
    else                    launch_part_pipe_st<u64, Src, SEG>(c, kc, algo_bytes, src, kin, vin, kout, vout, a, ntickets, status);
}
static const u32 kPipeTile = 384 * 10;

// Launch one partition pass with `grid` tickets.  `status` must hold grid * 256 status words (u32 when
// a.n < 2^30, else u64), zeroed; a.ticket one zeroed u32.
template <typename KeyT, typename ValT, typename Src, bool SEG>
static void launch_part_pass(Ctx &c, int kc, double algo_bytes, const Src &src, const KeyT *kin, const ValT *vin, KeyT *kout, ValT *vout,
                             const PartArgs &a, u64 grid, void *status)
{
    switch (part_variant()) {
    case 1:  launch_part_variant<KeyT, ValT, 384, 10, 3, Src, SEG>(c, kc, algo_bytes, src, kin, vin, kout, vout, a, grid, status); break;
    case 2:  launch_part_variant<KeyT, ValT, 512, 8, 2, Src, SEG>(c, kc, algo_bytes, src, kin, vin, kout, vout, a, grid, status); break;
    case 3:  launch_part_variant<KeyT, ValT, 256, 12, 4, Src, SEG>(c, kc, algo_bytes, src, kin, vin, kout, vout, a, grid, status); break;
    default: launch_part_variant<KeyT, ValT, 384, 12, 2, Src, SEG>(c, kc, algo_bytes, src, kin, vin, kout, vout, a, grid, status); break;
    }
}

// ---------------------------------------------------------------------------------------------
// hist16: histogram of the 16-bit prefixes of all n suffix keys, straight from the packed text (code width B
// divides 8, so a prefix is a whole number of symbols and every suffix's window starts on a symbol boundary).
// A CTA takes whole CHUNKS of the first partition pass (chunk s = elements [s * chunk_elems, ...) = a
// contiguous range of text positions), counts them in shared memory with packed 16-bit counters (128 KB for all
// 65536 bins), adds the chunk to the global histogram and stores the chunk's top-digit histogram H[s][256]
// (-> chunk prefixes cp of the first pass).  A 16-bit counter that wraps changes the chunk's total: the flag
// is set and the caller falls back to the stable LSD path (such a text is far too skewed for the MSD path anyway).
// ---------------------------------------------------------------------------------------------
static const int kHist16Threads = 1024;
static const int kHist16Words = 32768;

template <int B>
__device__ __forceinline__ void hist16_add(u32 *sh, u64 hi, u64 lw, int i)
{
    const int off = i * B;
    u32 win;
    if (off <= 48) win = (u32)(hi >> (48 - off)) & 0xFFFFu;
    else win = (u32)((hi << (off - 48)) | (lw >> (112 - off))) & 0xFFFFu;
    atomicAdd(&sh[win >> 1], 1u << ((win & 1u) << 4));
}

template <int B>
static __global__ void __launch_bounds__(kHist16Threads, 1)
hist16_kernel(const u64 *__restrict__ words, u64 n, u32 *__restrict__ hist, u32 *__restrict__ H, u32 nunits, u32 upc, u32 ut, u32 tpc, u32 tile,
              u64 *__restrict__ flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32 *sh = reinterpret_cast<u32 *>(smem_raw);
    __shared__ u32 s_red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int PER = 64 / B;                                    // suffixes whose window starts in one word
    // unit u = tiles [k * ut, (k + 1) * ut) of chunk s = u / upc (k = u % upc); the chunk's histogram H[s] is zeroed by the caller
    for (u32 unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const u32 chunk = unit / upc, k = unit % upc;
        const u64 tl0 = (u64)chunk * tpc + (u64)k * ut;
        u64 tl1 = tl0 + ut;
        if (tl1 > ((u64)chunk + 1) * tpc) tl1 = ((u64)chunk + 1) * tpc;
        const u64 e0 = tl0 * tile;
        const u64 e1 = tl1 * tile < n ? tl1 * tile : n;
        if (tl0 >= tl1 || e0 >= n) continue;                            // uniform over the CTA
        {
            uint4 *z = reinterpret_cast<uint4 *>(sh);
            for (int i = tid; i < kHist16Words / 4; i += kHist16Threads) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        const u64 p_lo = n - e1, p_hi = n - 1 - e0;                // positions of the chunk, inclusive
        const u64 w_first = (p_lo * B) >> 6, w_last = (p_hi * B) >> 6;
        // (the loads of the next step are issued before this step's shared atomics: one CTA per SM and a dependent
        // load per step left the kernel waiting on memory -- long scoreboard 10 per issue, 1.1 TB/s)
        u64 nhi = 0, nlw = 0;
        if (w_first + tid <= w_last) { nhi = words[w_first + tid]; nlw = words[w_first + tid + 1]; }
        for (u64 w = w_first + tid; w <= w_last; w += kHist16Threads) {
            const u64 hi = nhi, lw = nlw;
            if (w + kHist16Threads <= w_last) { nhi = words[w + kHist16Threads]; nlw = words[w + kHist16Threads + 1]; }
            const u64 q0 = w * PER;
            if (q0 >= p_lo && q0 + (PER - 1) <= p_hi) {
#pragma unroll
                for (int i = 0; i < PER; ++i) hist16_add<B>(sh, hi, lw, i);
            } else {
#pragma unroll 1
                for (int i = 0; i < PER; ++i) if (q0 + i >= p_lo && q0 + i <= p_hi) hist16_add<B>(sh, hi, lw, i);
            }
        }
        __syncthreads();
        // flush + check the total
        u32 total = 0;
        for (int i = tid; i < kHist16Words; i += kHist16Threads) {
            const u32 v = sh[i];
            if (v) {
                const u32 lo16 = v & 0xFFFFu, hi16 = v >> 16;
                if (lo16) atomicAdd(&hist[2 * i], lo16);
                if (hi16) atomicAdd(&hist[2 * i + 1], hi16);
                total += lo16 + hi16;
            }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
        if (lane == 0) s_red[warp] = total;
        // top-digit histogram of the chunk: digit d = bins [256 d, 256 d + 256) = words [128 d, 128 d + 128)
        if (tid < kRadixSize) {
            u32 sum = 0;
            for (int k = 0; k < 128; ++k) { const u32 v = sh[tid * 128 + ((k + tid) & 127)]; sum += (v & 0xFFFFu) + (v >> 16); }
            if (sum) atomicAdd(&H[(u64)chunk * kRadixSize + tid], sum);
        }
        __syncthreads();
        if (tid == 0) {
            u32 t = 0;
            for (int w = 0; w < kHist16Threads / 32; ++w) t += s_red[w];
            if ((u64)t != e1 - e0) *flag = 1;
        }
        __syncthreads();
    }
}

// One CTA: boff[0..65536] = exclusive scan of hist16 (boff[65536] = n), base256[c] = boff[c << 8] (u64, the digit
// bases of the first partition pass), tstart[0..256] = exclusive scan of the tiles per top-level bucket of the
// segmented pass, cp[s][d] = exclusive scan over the chunks of H[s][d] (in place), out[0] = largest bucket,
// out[1] = most tiles of any top-level bucket.
static __global__ void __launch_bounds__(1024)
scan16_kernel(const u32 *__restrict__ hist, u32 *__restrict__ boff, u64 *__restrict__ base256, u32 *__restrict__ tstart,
              u32 *__restrict__ H, u32 nchunks, u64 *__restrict__ out, u32 tile)
{
    __shared__ u32 s_tot[32];
    __shared__ u32 s_max[32];
    __shared__ u32 s_b[kRadixSize + 1];
    __shared__ u32 s_w[8];
    __shared__ u32 s_m[8];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint4 *h4 = reinterpret_cast<const uint4 *>(hist + t * 64);        // thread t owns bins [64t, 64t + 64)
    u32 sum = 0, mx = 0;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const uint4 v = h4[i];
        sum += v.x + v.y + v.z + v.w;
        const u32 m = max(max(v.x, v.y), max(v.z, v.w));
        mx = m > mx ? m : mx;
    }
    u32 inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
#pragma unroll
    for (int off = 16; off; off >>= 1) { u32 o = __shfl_xor_sync(0xffffffffu, mx, off); mx = o > mx ? o : mx; }
    if (lane == 31) s_tot[warp] = inc;
    if (lane == 0) s_max[warp] = mx;
    __syncthreads();
    if (warp == 0) {
        u32 x = s_tot[lane], y = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, y, off); if (lane >= off) y += o; }
        s_tot[lane] = y - x;
        u32 m = s_max[lane];
#pragma unroll
        for (int off = 16; off; off >>= 1) { u32 o = __shfl_xor_sync(0xffffffffu, m, off); m = o > m ? o : m; }
        if (lane == 0) out[0] = m;
    }
    __syncthreads();
    const u32 excl = s_tot[warp] + inc - sum;
    {
        u32 run = excl;
        uint4 *b4 = reinterpret_cast<uint4 *>(boff + t * 64);
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const uint4 v = h4[i];
            uint4 o;
            o.x = run; run += v.x; o.y = run; run += v.y; o.z = run; run += v.z; o.w = run; run += v.w;
            b4[i] = o;
        }
    }
    if ((t & 3) == 0) { s_b[t >> 2] = excl; base256[t >> 2] = excl; }       // bucket c << 8 is bin 64 t with t = 4c
    if (t == 1023) { boff[65536] = excl + sum; s_b[kRadixSize] = excl + sum; }
    __syncthreads();
    // tiles of the segmented pass per top-level bucket: cells of `tile` elements touched by [lo, hi)
    u32 nt = 0, ex = 0;
    if (t < kRadixSize) {
        const u32 lo = s_b[t], hi = s_b[t + 1];
        nt = hi > lo ? (hi - 1) / tile - lo / tile + 1 : 0;
        u32 x = nt, m = nt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += o; }
#pragma unroll
        for (int off = 16; off; off >>= 1) { u32 o = __shfl_xor_sync(0xffffffffu, m, off); m = o > m ? o : m; }
        if (lane == 31) s_w[warp] = x;
        if (lane == 0) s_m[warp] = m;
        ex = x - nt;
        // chunk prefixes of digit t
        u32 run = 0;
        for (u32 s = 0; s < nchunks; ++s) { const u32 v = H[(u64)s * kRadixSize + t]; H[(u64)s * kRadixSize + t] = run; run += v; }
    }
    __syncthreads();
    if (t < kRadixSize) {
        for (int w = 0; w < warp; ++w) ex += s_w[w];
        tstart[t] = ex;
        if (t == kRadixSize - 1) {
            tstart[kRadixSize] = ex + nt;
            u32 m = 0;
            for (int w = 0; w < 8; ++w) m = s_m[w] > m ? s_m[w] : m;
            out[1] = m;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bucket_sort: the array is partitioned by the key's top 16 bits (bucket b = [boff[b], boff[b+1])); finish
// every bucket inside shared memory and write the fully sorted (key, position) arrays.  CTA j takes the
// buckets that START in the window [j*C, (j+1)*C) of the array -- whole buckets, at most C + (largest
// bucket) - 1 <= CAP elements.  Order inside a bucket: by k-mer (key >> key_shift), ties by DESCENDING
// position -- exactly what the stable LSD sort of elements laid out in descending position yields, so the
// end-of-text rule of sa_core.cu holds unchanged.
//   1. histogram of the tile's elements over <= 8192 equal slices ("bins") of the tile's key range
//      (shared atomics on packed 16-bit counters), 2. exclusive scan, 3. elements dropped into their bin
//      (unordered), 4. every element counts the members of its bin that precede it (bins hold ~1 element
//      for uniform keys) and writes itself to its final slot -- a warp's 32 elements land in (nearly) the
//      same 32 consecutive slots, so the global stores coalesce without another staging step.
// When the tile's key range fits 32 bits (always, for the texts this path accepts in practice) an element is
// ONE 64-bit word in shared memory, (k-mer - first k-mer of the tile) << 32 | ~position: its bin is a shift,
// the order is an integer compare, and the first 8 elements of a thread stay in registers between steps 1 and 3
// (the others are re-read from global memory: L2 hits).  The preceding text byte of BWT calls rides in a byte
// array.  Wider ranges take the generic path (u64 key + u32 position per element).
// ---------------------------------------------------------------------------------------------
static const int kBucketCap = 7680;
static const int kBucketThreads = 512;
static const int kBucketBins = 8192;             // packed two per 32-bit word
static const int kBucketBinBits = 13;
static const int kBucketKeep = 8;                // elements of a thread kept in registers between steps 1 and 3
static const u32 kBucketMaxBucket = kBucketCap - 1536;     // largest 16-bit bucket the path accepts (window >= 1536)

struct BucketSmem {
    u64 keys[kBucketCap];                        // compact path: composite words
    u32 pos[kBucketCap];                         // compact path: preceding bytes (u8 view)
    u32 bins[kBucketBins / 2];
    u32 scan_tmp[32];
};

__device__ __forceinline__ u32 lower_bound_u32(const u32 *__restrict__ a, u32 lo, u32 hi, u64 x)
{
    // first index i in [lo, hi) with a[i] >= x, hi when none
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if ((u64)a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

// tile j of bucket_sort = the buckets that start in [j*C, (j+1)*C): tb[j] = first bucket with boff >= j*C
// (one thread per tile does the binary search once, so the sort CTAs start with a single load)
static __global__ void __launch_bounds__(256)
bucket_tiles_kernel(const u32 *__restrict__ boff, u64 ntiles, u32 C, uint4 *__restrict__ tb)
{
    const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j >= ntiles) return;
    const u32 B0 = lower_bound_u32(boff, 0, 65536, j * (u64)C);
    const u32 B1 = lower_bound_u32(boff, B0, 65536, (j + 1) * (u64)C);
    const u32 s = boff[B0];
    tb[2 * j] = make_uint4(s, boff[B1] - s, B0, B1 - B0);             // first element, count, first bucket, buckets
    // first elements of the next three buckets (32-bit keys: the bucket of an element is three compares, bucket_load_key)
    const u32 nb = B1 - B0;
    tb[2 * j + 1] = make_uint4(nb > 1 ? boff[B0 + 1] : 0xFFFFFFFFu, nb > 2 ? boff[B0 + 2] : 0xFFFFFFFFu, nb > 3 ? boff[B0 + 3] : 0xFFFFFFFFu, 0u);
}

// exclusive scan of the packed 16-bit bin counters (8192 bins, thread t owns words [8t, 8t + 8)); afterwards the
// low / high half of a word is the first slot of its bin.  All threads of the CTA call it.
__device__ __forceinline__ void bucket_scan_bins(BucketSmem &sm, int tid, int lane, int warp)
{
    u32 w[8]; u32 sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = sm.bins[tid * 8 + i]; sum += (w[i] & 0xFFFFu) + (w[i] >> 16); }
    u32 inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
    if (lane == 31) sm.scan_tmp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 x = lane < kBucketThreads / 32 ? sm.scan_tmp[lane] : 0, y = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, y, off); if (lane >= off) y += o; }
        sm.scan_tmp[lane] = y - x;
    }
    __syncthreads();
    u32 run = sm.scan_tmp[warp] + inc - sum;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const u32 a = w[i] & 0xFFFFu, bq = w[i] >> 16;
        sm.bins[tid * 8 + i] = run | ((run + a) << 16);
        run += a + bq;
    }
    __syncthreads();
}
// after step 3 a bin's counter is its END; its begin is the end of the bin before
__device__ __forceinline__ void bucket_bin_range(const BucketSmem &sm, u32 bin, u32 &beg, u32 &end)
{
    const u32 wv = sm.bins[bin >> 1];
    end = (bin & 1) ? (wv >> 16) : (wv & 0xFFFFu);
    beg = bin == 0 ? 0u : ((bin & 1) ? (wv & 0xFFFFu) : (sm.bins[(bin >> 1) - 1] >> 16));
}

// Fused rank stage (round 0): equal k-mers always share a bin, so the counting loop that ranks an element inside its
// bin also tells whether it heads its group (no equal k-mer sorts before it) and whether the group has other members.
// With `on` the kernel writes, instead of the sorted key array that rank_flags_kernel<true> would read back:
// positions in slot order (= the suffix array of round 0), the BWT row bytes, the head / active bits in the rank
// stage's mask layout (one byte per 4 slots: heads | actives << 4; rank_agg_kernel derives the per-warp and per-tile
// aggregates from them), and the primary index / aux samples of the suffixes that are already final.
// Suffixes that run past the end of the text (position >= tail_start) are groups of their own (sa_core.cu, end-of-text rule).
struct BucketFuse {
    int on;
    u8 *flags; u8 *rows; u64 tail_start;        // flags[slot]: bit 0 head, bit 1 active (rank_agg_kernel packs them into the mask bytes)
    u64 aux_mask; int aux_shift; u32 *aux_I; u64 *primary;
    u64 *big_flag;          // bit 0 is set when some group has more than 128 members (sa_core.cu: the lazy rounds then sort globally)
};

// key of element idx of the partitioned array.  in32: the array holds only the key bits below the 16-bit bucket
// prefix (PartArgs::out32); the prefix is the bucket the element lies in, B0 <= B < B0 + nb with boff[B] <= idx < boff[B+1].
// (b1..b3: boff[B0 + 1..3], read once per thread -- tiles of large texts hold a handful of buckets and the bucket is three
// compares; tiles with more buckets search boff)
template <bool IN32>
__device__ __forceinline__ u64 bucket_load_key(const u64 *__restrict__ kin, u64 idx, const u32 *__restrict__ boff, u32 B0, u32 nb, int lowbits,
                                               u32 b1, u32 b2, u32 b3)
{
    if (!IN32) return kin[idx];
    const u32 k32 = reinterpret_cast<const u32 *>(kin)[idx];
    u32 lo;
    if (nb <= 4) lo = B0 + (u32)(idx >= b1) + (u32)(idx >= b2) + (u32)(idx >= b3);
    else {
        lo = B0; u32 hi = B0 + nb;
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if ((u64)boff[mid] <= idx) lo = mid; else hi = mid; }
    }
    return ((u64)lo << lowbits) | (u64)k32;
}

// IN32 / FUSED are compile-time: with run-time flags the compiler no longer batched the tile's loads (3.36 vs 2.74 ms)
template <bool IN32, bool FUSED>
static __global__ void __launch_bounds__(kBucketThreads, 2)
bucket_sort_kernel(const u64 *__restrict__ kin, const u32 *__restrict__ vin, const uint4 *__restrict__ tb,
                   u64 n, u32 C, int key_shift, int R,             // R = K - 16: k-mer bits below the bucket prefix
                   u64 *__restrict__ kout, u32 *__restrict__ vout, u32 *err, const BucketFuse fz,
                   const u32 *__restrict__ boff, const u32 tile0)       // tile0: first tile of this launch (the tiles may be launched in chunks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BucketSmem &sm = *reinterpret_cast<BucketSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4 ti = tb[2 * (u64)(blockIdx.x + tile0)], tn = tb[2 * (u64)(blockIdx.x + tile0) + 1];
    const u64 s = ti.x;
    const u32 cnt = ti.y, B0 = ti.z, nb = ti.w;
    if (cnt == 0) return;
    if (cnt > (u32)kBucketCap) { if (tid == 0) *err = 3; return; }  // a bucket larger than promised
    const int span_bits = R + (nb > 1 ? 32 - __clz(nb - 1) : 0);     // (k-mer - kbase) < 2^span_bits
    const int sh = span_bits > kBucketBinBits ? span_bits - kBucketBinBits : 0;
    const u64 kbase = (u64)B0 << R;
    const u64 lowmask = ((u64)1 << key_shift) - 1;
    // first elements of the next three buckets, from the tile table (0xFFFFFFFF past the tile's buckets: never reached)
    const u32 bb1 = tn.x, bb2 = tn.y, bb3 = tn.z;

    for (int i = tid; i < kBucketBins / 2; i += kBucketThreads) sm.bins[i] = 0;
    __syncthreads();

    if (span_bits <= 32) {
        // ---- compact path
        u8 *prevb = reinterpret_cast<u8 *>(sm.pos);
        const int csh = 32 + sh;
        u64 keep[kBucketKeep]; u32 keepb = 0, keepb2 = 0;
        // 1. histogram (first kBucketKeep elements of the thread stay in registers)
#pragma unroll
        for (int q = 0; q < kBucketKeep; ++q) {
            const u32 i = q * kBucketThreads + tid;
            keep[q] = 0;
            if (i < cnt) {
                const u64 key = bucket_load_key<IN32>(kin, s + i, boff, B0, nb, R + key_shift, bb1, bb2, bb3);
                const u32 p = vin[s + i];
                const u64 comp = (((key >> key_shift) - kbase) << 32) | (u64)(~p);
                keep[q] = comp;
                if (q < 4) keepb |= (u32)(key & lowmask) << (8 * q); else keepb2 |= (u32)(key & lowmask) << (8 * (q - 4));
                const u32 bin = min((u32)(comp >> csh), (u32)kBucketBins - 1);
                atomicAdd(&sm.bins[bin >> 1], 1u << ((bin & 1) * 16));
            }
        }
        for (u32 i = kBucketKeep * kBucketThreads + tid; i < cnt; i += kBucketThreads) {
            const u64 rel = (bucket_load_key<IN32>(kin, s + i, boff, B0, nb, R + key_shift, bb1, bb2, bb3) >> key_shift) - kbase;
            const u32 bin = min((u32)(rel >> sh), (u32)kBucketBins - 1);
            atomicAdd(&sm.bins[bin >> 1], 1u << ((bin & 1) * 16));
        }
        __syncthreads();
        // 2. exclusive scan
        bucket_scan_bins(sm, tid, lane, warp);
        // 3. drop every element into its bin (the bin's counter becomes its end)
#pragma unroll
        for (int q = 0; q < kBucketKeep; ++q) {
            const u32 i = q * kBucketThreads + tid;
            if (i < cnt) {
                const u64 comp = keep[q];
                const u32 bin = min((u32)(comp >> csh), (u32)kBucketBins - 1);
                const u32 old = atomicAdd(&sm.bins[bin >> 1], 1u << ((bin & 1) * 16));
                const u32 slot = (old >> ((bin & 1) * 16)) & 0xFFFFu;
                sm.keys[slot] = comp;
                prevb[slot] = (u8)((q < 4 ? keepb >> (8 * q) : keepb2 >> (8 * (q - 4))) & 255u);
            }
        }
        for (u32 i = kBucketKeep * kBucketThreads + tid; i < cnt; i += kBucketThreads) {
            const u64 key = bucket_load_key<IN32>(kin, s + i, boff, B0, nb, R + key_shift, bb1, bb2, bb3);
            const u32 p = vin[s + i];
            const u64 comp = (((key >> key_shift) - kbase) << 32) | (u64)(~p);
            const u32 bin = min((u32)(comp >> csh), (u32)kBucketBins - 1);
            const u32 old = atomicAdd(&sm.bins[bin >> 1], 1u << ((bin & 1) * 16));
            const u32 slot = (old >> ((bin & 1) * 16)) & 0xFFFFu;
            sm.keys[slot] = comp;
            prevb[slot] = (u8)(key & lowmask);
        }
        __syncthreads();
        // 4. rank inside the bin (composite words are distinct: the positions are), write to the final slot
        if (!FUSED) {
            for (u32 i = tid; i < cnt; i += kBucketThreads) {
                const u64 comp = sm.keys[i];
                const u32 bin = min((u32)(comp >> csh), (u32)kBucketBins - 1);
                u32 beg, end;
                bucket_bin_range(sm, bin, beg, end);
                u32 r = 0;
                for (u32 q = beg; q < end; ++q) r += sm.keys[q] < comp ? 1u : 0u;
                const u64 o = s + beg + r;
                kout[o] = (((comp >> 32) + kbase) << key_shift) | (u64)prevb[i];
                vout[o] = ~(u32)comp;
            }
            return;
        }
        // fused rank stage: flag and row bytes go straight to their slots (a warp's elements land in ~32 consecutive slots)
        const u32 ts = fz.tail_start > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)fz.tail_start;
        for (u32 i = tid; i < cnt; i += kBucketThreads) {
            const u64 comp = sm.keys[i];
            const u32 bin = min((u32)(comp >> csh), (u32)kBucketBins - 1);
            u32 beg, end;
            bucket_bin_range(sm, bin, beg, end);
            const u32 p = ~(u32)comp, km = (u32)(comp >> 32);
            const bool tail = p >= ts;
            u32 r = 0, same = 0, before = 0;                             // members of my group (equal k-mer, not past the end), those before me
            for (u32 q = beg; q < end; ++q) {
                const u64 x = sm.keys[q];
                const bool lt = x < comp;
                const bool eq = (u32)(x >> 32) == km && ~(u32)x < ts;
                r += lt ? 1u : 0u; same += eq ? 1u : 0u; before += (eq && lt) ? 1u : 0u;
            }
            const bool head = tail || before == 0, active = !tail && same > 1;
            const u64 o = s + beg + r;
            vout[o] = p;
            fz.flags[o] = (u8)((head ? 1u : 0u) | (active ? 2u : 0u));
            if (same > 128u && head) atomicOr((unsigned long long *)fz.big_flag, 1ull);
            if (fz.rows) fz.rows[o] = prevb[i];
            if (!active) {
                if (p == 0) *fz.primary = o + 1;
                if (fz.aux_I && ((u64)p & fz.aux_mask) == 0) fz.aux_I[p >> fz.aux_shift] = (u32)o + 1;
            }
        }
        return;
    }

    // ---- generic path: u64 key + u32 position per element
    for (u32 i = tid; i < cnt; i += kBucketThreads) {
        const u64 km = bucket_load_key<IN32>(kin, s + i, boff, B0, nb, R + key_shift, bb1, bb2, bb3) >> key_shift;
        const u32 bin = min((u32)((km - kbase) >> sh), (u32)kBucketBins - 1);
        atomicAdd(&sm.bins[bin >> 1], 1u << ((bin & 1) * 16));
    }
    __syncthreads();
    bucket_scan_bins(sm, tid, lane, warp);
    for (u32 i = tid; i < cnt; i += kBucketThreads) {
        const u64 key = bucket_load_key<IN32>(kin, s + i, boff, B0, nb, R + key_shift, bb1, bb2, bb3);
        const u32 p = vin[s + i];
        const u32 bin = min((u32)(((key >> key_shift) - kbase) >> sh), (u32)kBucketBins - 1);
        const u32 old = atomicAdd(&sm.bins[bin >> 1], 1u << ((bin & 1) * 16));
        const u32 slot = (old >> ((bin & 1) * 16)) & 0xFFFFu;
        sm.keys[slot] = key;
        sm.pos[slot] = p;
    }
    __syncthreads();
    for (u32 i = tid; i < cnt; i += kBucketThreads) {
        const u64 key = sm.keys[i], km = key >> key_shift;
        const u32 p = sm.pos[i];
        const u32 bin = min((u32)((km - kbase) >> sh), (u32)kBucketBins - 1);
        u32 beg, end;
        bucket_bin_range(sm, bin, beg, end);
        u32 r = 0, same = 0, before = 0;
        for (u32 q = beg; q < end; ++q) {
            const u64 kj = sm.keys[q] >> key_shift;
            const bool lt = kj < km || (kj == km && sm.pos[q] > p);
            const bool eq = kj == km && (u64)sm.pos[q] < fz.tail_start;
            r += lt ? 1u : 0u; same += eq ? 1u : 0u; before += (eq && lt) ? 1u : 0u;
        }
        const u64 o = s + beg + r;
        vout[o] = p;
        if (!FUSED) { kout[o] = key; continue; }
        const bool tail = (u64)p >= fz.tail_start;
        const bool head = tail || before == 0, active = !tail && same > 1;
        fz.flags[o] = (u8)((head ? 1u : 0u) | (active ? 2u : 0u));
        if (same > 128u && head) atomicOr((unsigned long long *)fz.big_flag, 1ull);
        if (fz.rows) fz.rows[o] = (u8)(key & lowmask);
        if (!active) {
            if (p == 0) *fz.primary = o + 1;
            if (fz.aux_I && ((u64)p & fz.aux_mask) == 0) fz.aux_I[p >> fz.aux_shift] = (u32)o + 1;
        }
    }
}

}  // namespace lsc
