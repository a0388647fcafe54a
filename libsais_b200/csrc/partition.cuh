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
    u32  tstart[kRadixSize + 1];           // segmented mode: first tile of every top-level bucket
    alignas(8) u64 mbar[2];
    u32  tile;
};

// Segmented mode (second MSD level): the input is the concatenation of 256 top-level buckets
// [boff[c << 8], boff[(c + 1) << 8]); a tile is the intersection of a TILE-aligned cell of the array with one
// bucket, so interior tiles stay full and 16-byte aligned (TMA) and every tile has ONE top-level digit c.
// Its elements go to boff[c * 256 + d] + (elements with digit d in earlier tiles of the bucket) + rank.
struct SegArgs { const u32 *boff; const u32 *tstart; };

template <typename KeyT, typename ValT, int THREADS, int IPT, typename ST, typename Src, bool SEG>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? 3 : (THREADS <= 512 ? 2 : 1)))
part_pass_kernel(const KeyT *__restrict__ kin, const ValT *__restrict__ vin,
                 KeyT *__restrict__ kout, ValT *__restrict__ vout, u64 n,
                 int shift, u32 dmask, const u64 *__restrict__ base, const SegArgs seg,
                 ST *status, u32 *ticket, u32 *err, const Src src, const int use_bulk)
{
    typedef PartSmem<KeyT, ValT, THREADS, IPT> Smem;
    constexpr int TILE = THREADS * IPT;
    static_assert(THREADS >= kRadixSize + 1 && THREADS % 32 == 0, "one thread per digit (+1) is assumed");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        sm.tile = atomicAdd(ticket, 1u);
        mbar_init(&sm.mbar[0], 1); mbar_init(&sm.mbar[1], 1);
        mbar_fence_init();
    }
    if (tid < kRadixSize) sm.cnt[tid] = 0;
    if (SEG && tid <= kRadixSize) sm.tstart[tid] = seg.tstart[tid];
    __syncthreads();
    const u32 tile = sm.tile;

    // ---- which elements: [lo, lo + count) of the input; look-back stops at tile `first`
    u64 lo; u32 count, first = 0, bucket = 0;
    if (SEG) {
        if (tile >= sm.tstart[kRadixSize]) return;                 // grid is an upper bound
        int a = 0, bnd = kRadixSize;                               // largest c with tstart[c] <= tile
        while (bnd - a > 1) { const int mid = (a + bnd) >> 1; if (sm.tstart[mid] <= tile) a = mid; else bnd = mid; }
        bucket = (u32)a; first = sm.tstart[a];
        const u64 blo = seg.boff[(u32)a << 8], bhi = seg.boff[((u32)a + 1) << 8];
        const u64 cell = blo / TILE + (tile - first);
        const u64 clo = cell * TILE, chi = clo + TILE;
        lo = blo > clo ? blo : clo;
        count = (u32)((bhi < chi ? bhi : chi) - lo);
    } else {
        lo = (u64)tile * TILE;
        count = (u32)((n - lo) < (u64)TILE ? (n - lo) : (u64)TILE);
    }
    const bool full = count == (u32)TILE;

    // ---- load: element li of the tile is held by thread (li & 31) + 32 * warp-slot, warp-striped
    KeyT key[IPT];
    const u32 wbase = warp * (IPT * 32) + lane;
    const bool bulk = Src::kMode == SRC_ARRAYS && full && use_bulk && ((lo * sizeof(ValT)) & 15) == 0 && ((lo * sizeof(KeyT)) & 15) == 0;
    u64 gbase = 0;
    if (tid < kRadixSize) gbase = SEG ? (u64)seg.boff[(bucket << 8) + tid] : base[tid];
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
            t0 = (u64)(a0 - (uintptr_t)src.text);                   // may wrap "below" the text by < 4 bytes: a0 is still inside the allocation's 256-B alignment
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
                if (full || li < count) cp_async4(&sm.vals_in[li], vin + lo + li);
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
    if (tid < kRadixSize) {
        cnt = sm.cnt[tid];
        st_relaxed(status + (u64)tile * kRadixSize + tid, tile == first ? StWord<ST>::inc(cnt) : StWord<ST>::agg(cnt));
#pragma unroll
        for (int j = 0; j < kLookBatch; ++j) {
            const i64 idx = (i64)tile - 1 - j;
            ls.w[j] = idx >= (i64)first ? ld_relaxed(status + (u64)idx * kRadixSize + tid) : StWord<ST>::inc(0);
        }
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

    // ---- chained scan over the tiles of this segment
    if (tid < kRadixSize) {
        u64 excl = 0;
        if (tile != first) {
            i64 look = (i64)tile - 1;
            bool done = false;
#pragma unroll
            for (int j = 0; j < kLookBatch; ++j)
                if (!done) done = lookback_consume<ST>(ls.w[j], status, look - j, (u32)tid, excl, err);
            look -= kLookBatch;
            while (!done) {
                ST w[kLookRefill];
#pragma unroll
                for (int j = 0; j < kLookRefill; ++j) {
                    const i64 idx = look - j;
                    w[j] = idx >= (i64)first ? ld_relaxed(status + (u64)idx * kRadixSize + tid) : StWord<ST>::inc(0);
                }
#pragma unroll
                for (int j = 0; j < kLookRefill; ++j)
                    if (!done) done = lookback_consume<ST>(w[j], status, look - j, (u32)tid, excl, err);
                look -= kLookRefill;
            }
            st_relaxed(status + (u64)tile * kRadixSize + tid, StWord<ST>::inc(excl + (u64)cnt));
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
            const u64 g = sm.goff[digit_of(k, shift, dmask)] + idx;
            kout[g] = k;
            vout[g] = sm.vals[idx];
        }
    }
}

// Launch one partition pass.  `status` must hold (tiles + 1) * 256 words of ST, zeroed; `ticket` one zeroed u32.
template <typename KeyT, typename ValT, typename Src, bool SEG>
static void launch_part_pass(Ctx &c, int kc, double algo_bytes, const Src &src, const KeyT *kin, const ValT *vin, KeyT *kout, ValT *vout,
                             u64 n, u64 max_tiles, int shift, u32 dmask, const u64 *base, const SegArgs &seg,
                             void *status, u32 *ticket, u32 *err)
{
    constexpr int THREADS = 384, IPT = 12;
    typedef PartSmem<KeyT, ValT, THREADS, IPT> Smem;
    static const bool bulk_env = [] { const char *e = getenv("LIBSAIS_CUDA_TMA"); return !(e && *e && atoi(e) == 0); }();
    const int use_bulk = bulk_env && Src::kMode == SRC_ARRAYS && (((uintptr_t)kin | (uintptr_t)vin) & 15) == 0;
    if (n < (1ull << 30)) {
        auto kern = part_pass_kernel<KeyT, ValT, THREADS, IPT, u32, Src, SEG>;
        c.check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        LSC_LAUNCH(c, kc, algo_bytes, kern, (u32)max_tiles, THREADS, sizeof(Smem),
                   kin, vin, kout, vout, n, shift, dmask, base, seg, (u32 *)status, ticket, err, src, use_bulk);
    } else {
        auto kern = part_pass_kernel<KeyT, ValT, THREADS, IPT, u64, Src, SEG>;
        c.check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        LSC_LAUNCH(c, kc, algo_bytes, kern, (u32)max_tiles, THREADS, sizeof(Smem),
                   kin, vin, kout, vout, n, shift, dmask, base, seg, (u64 *)status, ticket, err, src, use_bulk);
    }
}
static const int kPartTile = 384 * 12;

// ---------------------------------------------------------------------------------------------
// hist16: histogram of the 16-bit prefixes of all n suffix keys, straight from the packed text (code width b
// divides 8, so a prefix is a whole number of symbols and every suffix's window starts on a symbol boundary).
// 65536 u32 bins do not fit one CTA's shared memory: a CTA counts only the windows of ITS half of the bin
// range (128 KB) over its share of the words; the two halves read the same words (L2 hits).
// ---------------------------------------------------------------------------------------------
static const int kHist16Threads = 1024;
static const int kHist16Ranges = 2;
static const int kHist16Bins = 65536 / kHist16Ranges;

static __global__ void __launch_bounds__(kHist16Threads, 1)
hist16_kernel(const u64 *__restrict__ words, u64 n, int b, u32 *__restrict__ hist)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32 *sh = reinterpret_cast<u32 *>(smem_raw);
    const int tid = threadIdx.x;
    const u32 range = blockIdx.x % kHist16Ranges;
    const u32 part = blockIdx.x / kHist16Ranges, nparts = gridDim.x / kHist16Ranges;
    for (int i = tid; i < kHist16Bins; i += kHist16Threads) sh[i] = 0;
    __syncthreads();
    const int per = 64 / b;                                      // suffixes whose window starts in one word
    const u64 nw = (n * (u64)b + 63) >> 6;
    for (u64 w = (u64)part * kHist16Threads + tid; w < nw; w += (u64)nparts * kHist16Threads) {
        const u64 hi = words[w], lw = words[w + 1];
        const u64 q0 = w * (u64)per;
        for (int i = 0; i < per; ++i) {
            if (q0 + i >= n) break;
            const int off = i * b;
            const u64 x = off ? ((hi << off) | (lw >> (64 - off))) : hi;
            const u32 win = (u32)(x >> 48);
            if ((win >> 15) == range) atomicAdd(&sh[win & (kHist16Bins - 1)], 1u);
        }
    }
    __syncthreads();
    for (int i = tid; i < kHist16Bins; i += kHist16Threads) {
        const u32 v = sh[i];
        if (v) atomicAdd(&hist[range * kHist16Bins + i], v);
    }
}

// One CTA: boff[0..65536] = exclusive scan of hist16 (boff[65536] = n), base256[c] = boff[c << 8] (u64, the digit
// bases of the first partition pass), tstart[0..256] = first tile of every top-level bucket in the segmented
// pass, out[0] = largest bucket, out[1] = number of tiles of the segmented pass.
static __global__ void __launch_bounds__(1024)
scan16_kernel(const u32 *__restrict__ hist, u32 *__restrict__ boff, u64 *__restrict__ base256, u32 *__restrict__ tstart,
              u64 *__restrict__ out, u32 tile)
{
    __shared__ u32 s_tot[32];
    __shared__ u32 s_max[32];
    __shared__ u32 s_b[kRadixSize + 1];
    __shared__ u32 s_w[8];
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
        u32 x = nt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += o; }
        if (lane == 31) s_w[warp] = x;
        ex = x - nt;
    }
    __syncthreads();
    if (t < kRadixSize) {
        for (int w = 0; w < warp; ++w) ex += s_w[w];
        tstart[t] = ex;
        if (t == kRadixSize - 1) { tstart[kRadixSize] = ex + nt; out[1] = ex + nt; }
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
// ---------------------------------------------------------------------------------------------
static const int kBucketCap = 7680;
static const int kBucketThreads = 512;
static const int kBucketBins = 8192;             // packed two per 32-bit word
static const int kBucketBinBits = 13;
static const u32 kBucketMaxBucket = kBucketCap - 1536;     // largest 16-bit bucket the path accepts (window >= 1536)

struct BucketSmem {
    u64 keys[kBucketCap];
    u32 pos[kBucketCap];
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
bucket_tiles_kernel(const u32 *__restrict__ boff, u64 ntiles, u32 C, u32 *__restrict__ tb)
{
    const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j <= ntiles) tb[j] = lower_bound_u32(boff, 0, 65536, j * (u64)C);
}

static __global__ void __launch_bounds__(kBucketThreads, 2)
bucket_sort_kernel(const u64 *__restrict__ kin, const u32 *__restrict__ vin, const u32 *__restrict__ boff, const u32 *__restrict__ tb,
                   u64 n, u32 C, int key_shift, int R,             // R = K - 16: k-mer bits below the bucket prefix
                   u64 *__restrict__ kout, u32 *__restrict__ vout, u32 *err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BucketSmem &sm = *reinterpret_cast<BucketSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 B0 = tb[blockIdx.x], B1 = tb[blockIdx.x + 1];
    const u64 s = boff[B0];
    const u32 cnt = (u32)((u64)boff[B1] - s);
    if (cnt == 0) return;
    if (cnt > (u32)kBucketCap) { if (tid == 0) *err = 3; return; }  // a bucket larger than promised
    const u32 nb = B1 - B0;
    const int sh_raw = R + (nb > 1 ? 32 - __clz(nb - 1) : 0) - kBucketBinBits;
    const int sh = sh_raw > 0 ? sh_raw : 0;
    const u64 kbase = (u64)B0 << R;

    for (int i = tid; i < kBucketBins / 2; i += kBucketThreads) sm.bins[i] = 0;
    __syncthreads();
    // 1. histogram
    for (u32 i = tid; i < cnt; i += kBucketThreads) {
        const u64 km = kin[s + i] >> key_shift;
        const u32 bin = min((u32)((km - kbase) >> sh), (u32)kBucketBins - 1);
        atomicAdd(&sm.bins[bin >> 1], 1u << ((bin & 1) * 16));
    }
    __syncthreads();
    // 2. exclusive scan over the 8192 bins: thread t owns words [8t, 8t + 8)
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
    }
    __syncthreads();
    // 3. drop every element into its bin (the bin's counter becomes its end)
    for (u32 i = tid; i < cnt; i += kBucketThreads) {
        const u64 key = kin[s + i];
        const u32 p = vin[s + i];
        const u32 bin = min((u32)(((key >> key_shift) - kbase) >> sh), (u32)kBucketBins - 1);
        const u32 old = atomicAdd(&sm.bins[bin >> 1], 1u << ((bin & 1) * 16));
        const u32 slot = (old >> ((bin & 1) * 16)) & 0xFFFFu;
        sm.keys[slot] = key;
        sm.pos[slot] = p;
    }
    __syncthreads();
    // 4. rank inside the bin, write to the final slot
    for (u32 i = tid; i < cnt; i += kBucketThreads) {
        const u64 key = sm.keys[i], km = key >> key_shift;
        const u32 p = sm.pos[i];
        const u32 bin = min((u32)((km - kbase) >> sh), (u32)kBucketBins - 1);
        const u32 wv = sm.bins[bin >> 1];
        const u32 end = (bin & 1) ? (wv >> 16) : (wv & 0xFFFFu);
        const u32 beg = bin == 0 ? 0u : ((bin & 1) ? (wv & 0xFFFFu) : (sm.bins[(bin >> 1) - 1] >> 16));
        u32 r = 0;
        for (u32 j = beg; j < end; ++j) {
            const u64 kj = sm.keys[j] >> key_shift;
            r += (kj < km || (kj == km && sm.pos[j] > p)) ? 1u : 0u;
        }
        const u64 o = s + beg + r;
        kout[o] = key;
        vout[o] = p;
    }
}

}  // namespace lsc
