// radix_sort.cuh -- hand-written onesweep LSD radix sort for (key, value) pairs on sm_100a.
//
// One upfront kernel builds the histograms of every digit (or the caller supplies them); each
// digit pass is then ONE kernel that reads the tile once and writes it once ("onesweep"): TMA
// bulk loads of the tile, warp-level multisplit with ballots, a chained scan over tiles with
// decoupled look-back per digit, staging of the tile in shared memory so the scatter leaves
// the SM as runs of consecutive addresses.  Stable; tiles are claimed through an atomic ticket
// so a tile's predecessors are always resident (forward progress of the look-back); a spin
// watchdog turns a would-be hang into an error code.
//
// Replaces, as the ordering engine of the SA core, the reference's induced-sorting scans
// (reference src/libsais.c:2157-4101 and :4777-6265); see DESIGN.md §3.
#pragma once
#include "common.cuh"
#include "ctx.h"
#include <cstdlib>

namespace lsc {

static const int kRadixBits = 8;
static const int kRadixSize = 256;
static const int kMaxPasses = 16;

// Tile status word of the chained scan: 2 flag bits (0 = not published, 1 = tile aggregate,
// 2 = inclusive prefix) above the count.  u32 words (30-bit counts) when n < 2^30 -- half the
// look-back traffic -- else u64.
template <typename ST> struct StWord {
    static const int kShift = (int)sizeof(ST) * 8 - 2;
    __host__ __device__ static ST agg(u64 v) { return (ST)(((ST)1 << kShift) | (ST)v); }
    __host__ __device__ static ST inc(u64 v) { return (ST)(((ST)2 << kShift) | (ST)v); }
    __host__ __device__ static u32 flag(ST w) { return (u32)(w >> kShift); }
    __host__ __device__ static u64 val(ST w) { return (u64)(w & ((((ST)1) << kShift) - 1)); }
};
static const u32 kSpinLimit  = 1u << 27;     // look-back watchdog: flag an error instead of hanging the GPU

struct SortPlan {
    int passes;
    int shift[kMaxPasses];
    int nbits[kMaxPasses];
};

static inline SortPlan make_sort_plan(int lo_bit, int hi_bit)
{
    SortPlan p; p.passes = 0;
    int bits = hi_bit - lo_bit;
    if (bits <= 0) return p;
    int P = (bits + kRadixBits - 1) / kRadixBits;
    int base = bits / P, rem = bits % P, s = lo_bit;
    for (int i = 0; i < P; ++i) {
        int nb = base + (i < rem ? 1 : 0);
        p.shift[i] = s; p.nbits[i] = nb; s += nb;
    }
    p.passes = P;
    return p;
}

template <typename KeyT>
__device__ __forceinline__ u32 digit_of(KeyT k, int shift, u32 mask) { return (u32)(k >> shift) & mask; }

// Key/value source of the FIRST pass.  NoGen: read the (key, value) arrays.  A generator type
// (kActive = true; key(i), val(i) device functions) produces element i on the fly instead, so
// the initial key array is never materialised: the histogram kernel and the first digit pass
// both evaluate the generator (sa_core.cu: k-mer keys straight from the packed text).
struct NoGen {
    static const bool kActive = false;
    __device__ __forceinline__ u64 key(u64) const { return 0; }
    __device__ __forceinline__ u32 val(u64) const { return 0; }
};

// One warp adds its 32 digits to a shared-memory histogram.  Sorted or skewed inputs put long runs of
// one digit into a warp (32 serialised adds to one address); each run is added once, by its first
// lane, with the run length.  Warp-collective: every lane calls it, `valid` masks lanes out.
__device__ __forceinline__ void warp_hist_add(u32 *bins, u32 d, bool valid, int lane)
{
    const u32 dp = __shfl_up_sync(0xffffffffu, d, 1);
    const u32 vm = __ballot_sync(0xffffffffu, valid);
    const bool head = valid && (lane == 0 || !((vm >> (lane - 1)) & 1u) || dp != d);
    const u32 hm = __ballot_sync(0xffffffffu, head);
    if (head) {
        const u32 above = ~((2u << lane) - 1u);            // lanes above mine (0 for lane 31)
        const u32 stop = (hm | ~vm) & above;               // next run head, or first masked-out lane
        const int end = stop ? __ffs(stop) - 1 : 32;
        atomicAdd(&bins[d], (u32)(end - lane));
    }
}

// ---------------------------------------------------------------------------------------------
// Upfront histograms of all digits: hist[pass][256] (u64).  One read of the keys.
// ---------------------------------------------------------------------------------------------
template <typename KeyT, int THREADS, typename Gen>
__global__ void __launch_bounds__(THREADS)
sort_hist_kernel(const KeyT *__restrict__ keys, u64 n, SortPlan plan, u64 *__restrict__ hist, const Gen gen)
{
    __shared__ u32 sh[kMaxPasses * kRadixSize];
    const int P = plan.passes, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < P * kRadixSize; i += THREADS) sh[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * THREADS;
    const u64 rounds = (n + stride - 1) / stride;
    for (u64 r = 0; r < rounds; ++r) {
        u64 idx = r * stride + (u64)blockIdx.x * THREADS + threadIdx.x;
        bool valid = idx < n;
        KeyT k = valid ? (Gen::kActive ? (KeyT)gen.key(idx) : keys[idx]) : (KeyT)0;
        for (int p = 0; p < P; ++p)
            warp_hist_add(sh + p * kRadixSize, digit_of(k, plan.shift[p], (1u << plan.nbits[p]) - 1), valid, lane);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * kRadixSize; i += THREADS) {
        u32 c = sh[i];
        if (c) atomicAdd((unsigned long long *)&hist[i], (unsigned long long)c);
    }
}

// Exclusive scan of each pass's histogram -> global base offset of every digit.  grid = passes.
static __global__ void sort_scan_kernel(const u64 *__restrict__ hist, u64 *__restrict__ base)
{
    __shared__ u64 s[kRadixSize];
    const int t = threadIdx.x;
    u64 v = hist[blockIdx.x * kRadixSize + t];
    s[t] = v;
    __syncthreads();
    for (int off = 1; off < kRadixSize; off <<= 1) {
        u64 add = t >= off ? s[t - off] : 0;
        __syncthreads();
        s[t] += add;
        __syncthreads();
    }
    base[blockIdx.x * kRadixSize + t] = s[t] - v;
}

// ---------------------------------------------------------------------------------------------
// One digit pass.  Template knobs:
//   VALS   : 1 = values loaded with the keys into registers, 2 = prefetched into shared memory
//            with cp.async (no registers; latency overlapped with the ranking)
//   NMATCH : how many of a thread's IPT items are ranked with match.any (one instruction, but
//            the shared ADU pipe runs it at ~2 cycles per lane); the rest use 8 ballots (ALU).
//            Splitting the items balances the two pipes.
// Phases of a tile: load -> digit counts (shared atomics) -> publish counts EARLY, before the
// long ranking phase, so successors' look-back rarely meets an unpublished tile -> warp-level
// multisplit ranking -> per-digit scan over warps + decoupled look-back -> scatter into shared
// memory in sorted order -> coalesced write-out of each digit's run.
// ---------------------------------------------------------------------------------------------
template <typename KeyT, typename ValT, int THREADS, int IPT, int VALS>
struct PassSmem {
    static const int TILE = THREADS * IPT;
    KeyT keys[TILE];
    u64  goff[kRadixSize];                 // global offset of a digit's run minus its offset in the tile
    ValT vals[TILE];
    ValT vals_in[VALS == 2 ? TILE : 1];
    alignas(16) u32 whist[(THREADS / 32) * kRadixSize];   // zeroed with 16-byte stores
    u32  cnt[kRadixSize];
    u32  scan_tmp[32];
    alignas(8) u64 mbar[2];                // TMA completion barriers: [0] keys, [1] values
    u32  tile;
};

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src)
{
    u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gmem_src) : "memory");
}
template <typename T> __device__ __forceinline__ void cp_async_val(T *smem_dst, const T *gmem_src)
{
    if (sizeof(T) == 8) cp_async8(smem_dst, gmem_src); else cp_async4(smem_dst, gmem_src);
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: one elected thread moves a
// whole tile of keys / values from global into shared memory; no per-thread load instructions, no registers.
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    u32 done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

static const int kLookBatch = 8;           // predecessor status words prefetched before the ranking
static const int kLookRefill = 16;         // words per further round trip

// Decoupled look-back for one digit, split in two so the first round trip overlaps the ranking:
// lookback_prefetch() issues the loads of the kLookBatch nearest predecessors' status words right
// after this tile published its counts; lookback_finish() consumes them (by then they have
// arrived), polls words that were not published yet, and continues further back if needed
// (measured on B200, 2 x 148 tiles in flight: the nearest inclusive prefix is ~22 tiles back).
template <typename ST> struct LookState { ST w[kLookBatch]; };

template <typename ST>
__device__ __forceinline__ void lookback_prefetch(LookState<ST> &ls, const ST *status, u32 tile, u32 digit)
{
#pragma unroll
    for (int j = 0; j < kLookBatch; ++j) {
        i64 idx = (i64)tile - 1 - j;
        ls.w[j] = idx >= 0 ? ld_relaxed(status + (u64)idx * kRadixSize + digit) : StWord<ST>::inc(0);
    }
}

// One status word of a predecessor tile: poll until published, accumulate; true when it carried
// an inclusive prefix (the look-back is complete).
template <typename ST>
__device__ __forceinline__ bool lookback_consume(ST x, const ST *status, i64 idx, u32 digit, u64 &excl, u32 *err)
{
    u32 spins = 0;
    while (StWord<ST>::flag(x) == 0) {
        if (++spins > kSpinLimit) { *err = 1; x = StWord<ST>::inc(0); break; }
        __nanosleep(20);
        x = ld_relaxed(status + (u64)idx * kRadixSize + digit);
    }
    excl += StWord<ST>::val(x);
    return StWord<ST>::flag(x) == 2;
}

template <typename ST>
__device__ __forceinline__ u64 lookback_finish(LookState<ST> &ls, ST *status, u32 tile, u32 digit, u32 cnt, u32 *err)
{
    if (tile == 0) return 0;
    u64 excl = 0;
    i64 look = (i64)tile - 1;
    bool done = false;
#pragma unroll
    for (int j = 0; j < kLookBatch; ++j)
        if (!done) done = lookback_consume<ST>(ls.w[j], status, look - j, digit, excl, err);
    look -= kLookBatch;
    while (!done) {
        ST w[kLookRefill];
#pragma unroll
        for (int j = 0; j < kLookRefill; ++j) {
            i64 idx = look - j;
            w[j] = idx >= 0 ? ld_relaxed(status + (u64)idx * kRadixSize + digit) : StWord<ST>::inc(0);
        }
#pragma unroll
        for (int j = 0; j < kLookRefill; ++j)
            if (!done) done = lookback_consume<ST>(w[j], status, look - j, digit, excl, err);
        look -= kLookRefill;
    }
    st_relaxed(status + (u64)tile * kRadixSize + digit, StWord<ST>::inc(excl + (u64)cnt));
    return excl;
}

// peers of this lane: lanes whose (valid) item has the same digit -- AND over one ballot per digit bit
__device__ __forceinline__ u32 digit_peers(u32 d, bool valid, int lane)
{
    u32 peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bit = 0; bit < kRadixBits; ++bit) {
        bool on = (d >> bit) & 1;
        u32 m = __ballot_sync(0xffffffffu, on);
        peers &= on ? m : ~m;
    }
    return valid ? peers : (1u << lane);
}

template <typename KeyT, typename ValT, int THREADS, int IPT, int VALS, typename ST, bool FULL, typename Gen>
__device__ __forceinline__ void sort_pass_tile(PassSmem<KeyT, ValT, THREADS, IPT, VALS> &sm, const Gen &gen,
                                               const KeyT *__restrict__ kin, const ValT *__restrict__ vin,
                                               KeyT *__restrict__ kout, ValT *__restrict__ vout,
                                               const u32 tile, const u32 count, int shift, u32 dmask,
                                               const u64 *__restrict__ base, ST *status, u32 *err, const bool use_bulk)
{
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * IPT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tile_base = (u64)tile * TILE;
    kin += tile_base; vin += tile_base;

    // ---- load keys (+ values), warp-striped: element order inside the tile = (warp, item, lane).
    // Full tiles of 16-byte aligned arrays come in by TMA: one thread issues two bulk copies (keys into the
    // staging area the sorted tile will later overwrite, values into vals_in); everybody else just waits
    // on the mbarrier and reads shared memory.
    KeyT key[IPT];
    ValT val[VALS == 1 ? IPT : 1];
    const u32 wbase = warp * (IPT * 32) + lane;
    const bool bulk = FULL && VALS == 2 && !Gen::kActive && use_bulk;
    if (bulk) {
        if (tid == 0) {
            mbar_expect_tx(&sm.mbar[0], (u32)(TILE * sizeof(KeyT)));
            bulk_load(sm.keys, kin, (u32)(TILE * sizeof(KeyT)), &sm.mbar[0]);
            mbar_expect_tx(&sm.mbar[1], (u32)(TILE * sizeof(ValT)));
            bulk_load(sm.vals_in, vin, (u32)(TILE * sizeof(ValT)), &sm.mbar[1]);
        }
        mbar_wait(&sm.mbar[0], 0);
#pragma unroll
        for (int i = 0; i < IPT; ++i) key[i] = sm.keys[wbase + i * 32];
    } else {
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            u32 li = wbase + i * 32;
            key[i] = (FULL || li < count) ? (Gen::kActive ? (KeyT)gen.key(tile_base + li) : kin[li]) : (KeyT)0;
        }
        if (Gen::kActive) {
            // values are recomputed at scatter time
        } else if (VALS == 1) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                u32 li = wbase + i * 32;
                val[VALS == 1 ? i : 0] = (FULL || li < count) ? vin[li] : (ValT)0;
            }
        } else {
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                u32 li = wbase + i * 32;
                if (FULL || li < count) cp_async_val(&sm.vals_in[VALS == 2 ? li : 0], vin + li);
            }
        }
    }

    // ---- counting pre-pass: digit counts of every warp (warp-private shared counters)
    u32 *wh = sm.whist + warp * kRadixSize;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        u32 li = wbase + i * 32;
        if (FULL || li < count) atomicAdd(&wh[digit_of(key[i], shift, dmask)], 1u);
    }
    __syncthreads();

    // ---- per digit: tile count -> publish EARLY (before the long ranking phase), prefetch the
    // look-back, offsets of every (warp, digit) run inside the tile
    u32 cnt = 0, tileoff = 0;
    u64 gbase = 0;
    LookState<ST> ls;
    if (tid < kRadixSize) {
        gbase = base[tid];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) cnt += sm.whist[w * kRadixSize + tid];
        st_relaxed(status + (u64)tile * kRadixSize + tid, tile == 0 ? StWord<ST>::inc(cnt) : StWord<ST>::agg(cnt));
        lookback_prefetch<ST>(ls, status, tile, (u32)tid);
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
        u32 run = tileoff;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { u32 t = sm.whist[w * kRadixSize + tid]; sm.whist[w * kRadixSize + tid] = run; run += t; }
    }
    __syncthreads();

    // ---- warp-level multisplit + scatter into shared memory in sorted order: the group leader
    // claims the run's next slots with one shared atomic (returns the group's base), every lane
    // adds its rank inside the group.  Items of a warp are claimed in program order -> stable.
    const u32 lt = lanemask_lt();
    constexpr bool STASH_IN_KEY = sizeof(KeyT) >= 4;       // the slot of a cp.async value rides in the dead key register
    u32 spos[(VALS == 2 && !STASH_IN_KEY) ? IPT : 1];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        u32 li = wbase + i * 32;
        bool valid = FULL || li < count;
        u32 d = valid ? digit_of(key[i], shift, dmask) : (u32)kRadixSize;
        u32 peers = digit_peers(d, valid, lane);
        int leader = __ffs(peers) - 1;
        u32 basepos = 0;
        if (lane == leader && valid) basepos = atomicAdd(&wh[d], (u32)__popc(peers));
        basepos = __shfl_sync(0xffffffffu, basepos, leader);
        if (valid) {
            u32 pos = basepos + __popc(peers & lt);
            sm.keys[pos] = key[i];
            if (Gen::kActive) sm.vals[pos] = (ValT)gen.val(tile_base + li);
            else if (VALS == 1) sm.vals[pos] = val[VALS == 1 ? i : 0];
            else if (STASH_IN_KEY) key[i] = (KeyT)pos;    // remember the slot for the value (cp.async still in flight)
            else spos[(VALS == 2 && !STASH_IN_KEY) ? i : 0] = pos;
        }
    }
    if (VALS == 2 && !Gen::kActive) {
        if (bulk) mbar_wait(&sm.mbar[1], 0); else cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            u32 li = wbase + i * 32;
            if (FULL || li < count) sm.vals[STASH_IN_KEY ? (u32)key[i] : spos[(VALS == 2 && !STASH_IN_KEY) ? i : 0]] = sm.vals_in[VALS == 2 ? li : 0];
        }
    }

    // ---- chained scan over tiles (first round trip was prefetched before the ranking)
    if (tid < kRadixSize) {
        u64 excl = lookback_finish<ST>(ls, status, tile, (u32)tid, cnt, err);
        sm.goff[tid] = gbase + excl - (u64)tileoff;
    }
    __syncthreads();

    // ---- write out: consecutive threads own consecutive slots of a digit's run
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        u32 idx = i * THREADS + tid;
        if (FULL || idx < count) {
            KeyT k = sm.keys[idx];
            u64 g = sm.goff[digit_of(k, shift, dmask)] + idx;
            kout[g] = k;
            vout[g] = sm.vals[idx];
        }
    }
}

template <typename KeyT, typename ValT, int THREADS, int IPT, int VALS, typename ST, typename Gen>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? 3 : (THREADS <= 512 ? 2 : 1)))
sort_pass_kernel(const KeyT *__restrict__ kin, const ValT *__restrict__ vin,
                 KeyT *__restrict__ kout, ValT *__restrict__ vout, u64 n,
                 int shift, u32 dmask, const u64 *__restrict__ base,
                 ST *status, u32 *ticket, u32 *err, const Gen gen, const int use_bulk)
{
    typedef PassSmem<KeyT, ValT, THREADS, IPT, VALS> Smem;
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * IPT;
    static_assert(THREADS >= kRadixSize && THREADS % 32 == 0, "one thread per digit is assumed");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x;

    if (tid == 0) {
        sm.tile = atomicAdd(ticket, 1u);
        mbar_init(&sm.mbar[0], 1); mbar_init(&sm.mbar[1], 1);
        mbar_fence_init();
    }
    {
        uint4 *z = reinterpret_cast<uint4 *>(sm.whist);
        for (int i = tid; i < WARPS * kRadixSize / 4; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
        if (tid < kRadixSize) sm.cnt[tid] = 0;
    }
    __syncthreads();
    const u32 tile = sm.tile;
    const u64 tile_base = (u64)tile * TILE;
    const u32 count = (u32)((n - tile_base) < (u64)TILE ? (n - tile_base) : (u64)TILE);
    if (count == TILE)
        sort_pass_tile<KeyT, ValT, THREADS, IPT, VALS, ST, true, Gen>(sm, gen, kin, vin, kout, vout, tile, count, shift, dmask, base, status, err, use_bulk != 0);
    else
        sort_pass_tile<KeyT, ValT, THREADS, IPT, VALS, ST, false, Gen>(sm, gen, kin, vin, kout, vout, tile, count, shift, dmask, base, status, err, false);
}

// Tile shapes of the pass kernel; LIBSAIS_CUDA_SORT_VARIANT selects one for experiments.
// Measured on B200, 2^28 (u64,u32) pairs, 8-bit digits (profiles/sort_pass_variants_r1.md):
// two to three independent CTAs per SM beat one large CTA; 384 x 12 is the fastest.
struct PassVariant { int threads, ipt, vals; };
static const PassVariant kPassVariants[] = { {384, 12, 2}, {512, 8, 1}, {256, 16, 2} };
static const int kNumPassVariants = sizeof(kPassVariants) / sizeof(kPassVariants[0]);
static const int kDefaultPassVariant = 0;

static inline int pass_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LIBSAIS_CUDA_SORT_VARIANT");
        v = (e && *e) ? atoi(e) : kDefaultPassVariant;
        if (v < 0 || v >= kNumPassVariants) v = kDefaultPassVariant;
    }
    return v;
}

template <typename KeyT, typename ValT>
struct RadixSort {
    static const int HIST_THREADS = 512;

    static int tile_elems() { if (sizeof(ValT) == 8) return 256 * 16; const PassVariant &pv = kPassVariants[pass_variant()]; return pv.threads * pv.ipt; }
    static u64 tiles(u64 n) { return ceil_div(n, (u64)tile_elems()); }

    static size_t temp_bytes(u64 n)
    {
        return 2 * kMaxPasses * kRadixSize * sizeof(u64)      // hist + base
             + 256                                            // tickets (u32[kMaxPasses])
             + ceil_div(n, 3072) * kRadixSize * sizeof(u64);  // status of one pass (smallest tile of any variant)
    }

    template <int THREADS, int IPT, int VALS, typename ST, typename Gen>
    static void launch_pass_st(Ctx &c, const Gen &gen, const KeyT *kin, const ValT *vin, KeyT *kout, ValT *vout, u64 n,
                               int shift, u32 dmask, const u64 *base, u64 *status_raw, u32 *ticket, u32 *err)
    {
        typedef PassSmem<KeyT, ValT, THREADS, IPT, VALS> Smem;
        auto kern = sort_pass_kernel<KeyT, ValT, THREADS, IPT, VALS, ST, Gen>;
        ST *status = reinterpret_cast<ST *>(status_raw);
        // TMA bulk loads need 16-byte aligned sources (tile offsets are multiples of 16 bytes by construction)
        static const bool bulk_env = [] { const char *e = getenv("LIBSAIS_CUDA_TMA"); return !(e && *e && atoi(e) == 0); }();
        const int use_bulk = bulk_env && !Gen::kActive && (((uintptr_t)kin | (uintptr_t)vin) & 15) == 0
                             && ((size_t)THREADS * IPT * sizeof(KeyT)) % 16 == 0 && ((size_t)THREADS * IPT * sizeof(ValT)) % 16 == 0;
        c.check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        u64 nt = ceil_div(n, (u64)THREADS * IPT);
        const double in_bytes = Gen::kActive ? 2.0 : (double)(sizeof(KeyT) + sizeof(ValT));   // generator: packed text + bwt byte
        const int kc = c.pass_class_override >= 0 ? c.pass_class_override : (Gen::kActive ? KC_SORT_PASS_GEN : KC_SORT_PASS);
        LSC_LAUNCH(c, kc, (double)n * (in_bytes + sizeof(KeyT) + sizeof(ValT)), kern, (u32)nt, THREADS, sizeof(Smem),
                   kin, vin, kout, vout, n, shift, dmask, base, status, ticket, err, gen, use_bulk);
    }

    template <int THREADS, int IPT, int VALS, typename Gen>
    static void launch_pass(Ctx &c, const Gen &gen, const KeyT *kin, const ValT *vin, KeyT *kout, ValT *vout, u64 n,
                            int shift, u32 dmask, const u64 *base, u64 *status, u32 *ticket, u32 *err)
    {
        if (n < (1ull << 30)) launch_pass_st<THREADS, IPT, VALS, u32, Gen>(c, gen, kin, vin, kout, vout, n, shift, dmask, base, status, ticket, err);
        else                  launch_pass_st<THREADS, IPT, VALS, u64, Gen>(c, gen, kin, vin, kout, vout, n, shift, dmask, base, status, ticket, err);
    }

    template <typename Gen>
    static void pass(Ctx &c, const Gen &gen, const KeyT *kin, const ValT *vin, KeyT *kout, ValT *vout, u64 n,
                     int shift, u32 dmask, const u64 *base, u64 *status, u32 *ticket, u32 *err)
    {
        if (sizeof(ValT) == 8) {            // 64-bit values (distributed path): the 4096-pair tile keeps two CTAs per SM
            launch_pass<256, 16, 2, Gen>(c, gen, kin, vin, kout, vout, n, shift, dmask, base, status, ticket, err);
            return;
        }
        switch (pass_variant()) {
        case 1:  launch_pass<512, 8, 1, Gen>(c, gen, kin, vin, kout, vout, n, shift, dmask, base, status, ticket, err); break;
        case 2:  launch_pass<256, 16, 2, Gen>(c, gen, kin, vin, kout, vout, n, shift, dmask, base, status, ticket, err); break;
        default: launch_pass<384, 12, 2, Gen>(c, gen, kin, vin, kout, vout, n, shift, dmask, base, status, ticket, err); break;
        }
    }

    // Sort n pairs on key bits [lo_bit, hi_bit).  Input in (ka, va) -- or produced by `gen` when
    // Gen::kActive, in which case (ka, va) is only the alternate buffer; (kb, vb) is the other
    // buffer.  Returns 0 when the result is in (ka, va), 1 when in (kb, vb), -1 on error.
    template <typename Gen>
    static int sort_from(Ctx &c, const Gen &gen, KeyT *ka, ValT *va, KeyT *kb, ValT *vb, u64 n, int lo_bit, int hi_bit,
                         void *temp, u32 *err, int *passes_out = nullptr, bool hist_ready = false)
    {
        // hist_ready: the caller already wrote the digit histograms (u64[passes][256]) at the start of `temp`
        SortPlan plan = make_sort_plan(lo_bit, hi_bit);
        if (passes_out) *passes_out = plan.passes;
        if (n == 0 || plan.passes == 0) return Gen::kActive ? -1 : 0;
        char *t = (char *)temp;
        u64 *hist = (u64 *)t;               t += kMaxPasses * kRadixSize * sizeof(u64);
        u64 *base = (u64 *)t;               t += kMaxPasses * kRadixSize * sizeof(u64);
        u32 *tickets = (u32 *)t;            t += 256;
        u64 *status = (u64 *)t;
        const u64 nt = tiles(n);

        if (hist_ready) c.check(cudaMemsetAsync(tickets, 0, 256, c.stream));
        else c.check(cudaMemsetAsync(hist, 0, 2 * kMaxPasses * kRadixSize * sizeof(u64) + 256, c.stream));
        if (!hist_ready) {
            u64 want = ceil_div(n, (u64)HIST_THREADS * 8);
            u32 grid = (u32)(want < (u64)c.sm_count * 4 ? (want ? want : 1) : (u64)c.sm_count * 4);
            LSC_LAUNCH(c, c.pass_class_override >= 0 ? c.pass_class_override : KC_SORT_HIST, (double)n * (Gen::kActive ? 2.0 : (double)sizeof(KeyT)), (sort_hist_kernel<KeyT, HIST_THREADS, Gen>),
                       grid, HIST_THREADS, 0, ka, n, plan, hist, gen);
        }
        LSC_LAUNCH(c, c.pass_class_override >= 0 ? c.pass_class_override : KC_SORT_SCAN, 0.0, sort_scan_kernel, plan.passes, kRadixSize, 0, hist, base);

        KeyT *kin = ka, *kout = kb; ValT *vin = va, *vout = vb;
        int where = 0;
        for (int p = 0; p < plan.passes; ++p) {
            c.check(cudaMemsetAsync(status, 0, nt * kRadixSize * (n < (1ull << 30) ? sizeof(u32) : sizeof(u64)), c.stream));   // u32 status words below 2^30 (launch_pass)
            const int shift = plan.shift[p]; const u32 dmask = (1u << plan.nbits[p]) - 1;
            if (p == 0 && Gen::kActive) pass<Gen>(c, gen, kin, vin, kout, vout, n, shift, dmask, base, status, tickets, err);
            else pass<NoGen>(c, NoGen(), kin, vin, kout, vout, n, shift, dmask, base + p * kRadixSize, status, tickets + p, err);
            KeyT *tk2 = kin; kin = kout; kout = tk2;
            ValT *tv = vin; vin = vout; vout = tv;
            where ^= 1;
        }
        return c.failed() ? -1 : where;
    }

    static int sort(Ctx &c, KeyT *ka, ValT *va, KeyT *kb, ValT *vb, u64 n, int lo_bit, int hi_bit,
                    void *temp, u32 *err, int *passes_out = nullptr)
    {
        return sort_from<NoGen>(c, NoGen(), ka, va, kb, vb, n, lo_bit, hi_bit, temp, err, passes_out);
    }
};

}  // namespace lsc
