// common.cuh -- shared device/host helpers for libsais_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace lsc {

typedef uint8_t  u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t  i64;

// Largest text length the single-GPU core handles: positions, slots and ranks are u32
// (rank+1 and the sentinel n must still fit).
static const u64 kMaxN = 0xFFFFFFF0ull;

static inline u64 ceil_div(u64 a, u64 b) { return (a + b - 1) / b; }
static inline int bits_for(u64 maxval) { int b = 0; while (maxval) { ++b; maxval >>= 1; } return b ? b : 1; }

// Kernel classes, for launch counting and the per-class device-time / algorithmic-byte report.
enum KernelClass {
    KC_HIST_SYM = 0,   // byte histogram (freq) / max symbol
    KC_PACK,           // symbol -> b-bit code bitstream
    KC_MAKE_KEYS,      // round-0 k-mer key build
    KC_SORT_HIST,      // onesweep upfront digit histograms
    KC_SORT_SCAN,      // digit-base exclusive scan
    KC_SORT_PASS,      // onesweep digit pass (the dominant kernel)
    KC_SORT_PASS_GEN,  // first digit pass fused with k-mer key generation (reads the packed text, not a key array)
    KC_RANK_INIT,      // round-0 head flags + rank + ISA scatter + compaction
    KC_RANK_SCAN,      // scan of the rank stage's tile aggregates
    KC_ROUND_KEYS,     // round>=1 key build (ISA gather)
    KC_RANK_UPDATE,    // round>=1 rank update + ISA scatter + compaction
    KC_SCATTER,        // locality-partitioned scatter of (index, value) pairs (ISA, phi): partition pass + streaming scatter
    KC_BWT,            // BWT gather (+ primary, aux)
    KC_PHI,            // phi scatter
    KC_PLCP,           // PLCP compare
    KC_LCP,            // LCP permute
    KC_UNBWT_PREP,     // unBWT: L' build / psi via counting sort helper kernels
    KC_UNBWT_WALK,     // unBWT: splitter walks
    KC_UNBWT_RANK,     // unBWT: splitter list ranking
    KC_CONVERT,        // widen / narrow / copy helpers
    KC_LOCAL_SORT,     // round>=1 in-shared-memory sort of small groups (key build + sort in one kernel)
    KC_PART_PASS,      // unstable digit partition pass (round-0 MSD levels; the stable pass is KC_SORT_PASS)
    KC_BUCKET_SORT,    // round-0 MSD finish: every 16-bit bucket sorted inside shared memory
    KC_COUNT
};

static const char *const kKernelClassName[KC_COUNT] = {
    "hist_sym", "pack", "make_keys", "sort_hist", "sort_scan", "sort_pass", "sort_pass_gen", "rank_init", "rank_scan",
    "round_keys", "rank_update", "scatter", "bwt", "phi", "plcp", "lcp",
    "unbwt_prep", "unbwt_walk", "unbwt_rank", "convert", "local_sort", "part_pass", "bucket_sort"
};

#ifdef __CUDACC__
// ---- status-word loads/stores for chained scans: bypass L1, no fence needed because value
// ---- and flag travel in one 64-bit word.
__device__ __forceinline__ u64 ld_relaxed(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u64 *p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u32 ld_relaxed(const u32 *p)
{
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u32 *p, u32 v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// streaming (read-once) loads that do not allocate in L1
__device__ __forceinline__ u64 ld_stream(const u64 *p)
{
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 ld_stream(const u32 *p)
{
    u32 v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
#endif

}  // namespace lsc
