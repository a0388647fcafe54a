// scatter.cuh -- dst[idx[i]] = val[i] for N pairs, with locality.
//
// A random 4-byte scatter over a multi-GB array runs at ~20 G elements/s on B200 (every store is
// a partial-sector write that misses L2 and costs a DRAM read-modify-write).  Partitioning the
// pairs first by the top 8 bits of the destination index -- one onesweep digit pass, streaming --
// makes the following scatter walk the destination window by window (dst_len*4/256 bytes each):
// the partial writes merge in L2 and leave as full sectors.  Used for the ISA updates of the
// prefix-doubling rounds and for the phi array (reference compute_phi, src/libsais.c:8116-8142).
#pragma once
#include "radix_sort.cuh"
#include "partition.cuh"

namespace lsc {

static __global__ void __launch_bounds__(256)
scatter_pairs_kernel(const u32 *__restrict__ idx, const u32 *__restrict__ val, u64 N, u32 *__restrict__ dst, u64 dst_len)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    u32 p = ld_stream(idx + i);
    if ((u64)p < dst_len) dst[p] = ld_stream(val + i);
}

template <typename Gen>
static __global__ void __launch_bounds__(256)
scatter_gen_kernel(const Gen gen, u64 N, u32 *__restrict__ dst, u64 dst_len)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    u32 p = (u32)gen.key(i);
    if ((u64)p < dst_len) dst[p] = gen.val(i);
}

static inline bool scatter_is_direct(u64 N, u64 dst_len) { return dst_len <= (8ull << 20) || N < (1ull << 20); }

// The producer of the pairs can take the histogram of the partition digit itself (one shared-memory add per
// pair while the position is in a register) and save the partition pass its read of the index array: zeroes
// the histogram slot of `sort_temp` and returns where to count and the digit's shift; false when the scatter
// will be direct and needs no histogram.
static inline bool scatter_hist_prepare(Ctx &c, u64 N, u64 dst_len, void *sort_temp, u64 **hist, int *shift)
{
    *hist = nullptr; *shift = 0;
    if (N == 0 || scatter_is_direct(N, dst_len)) return false;
    const int bits = bits_for(dst_len - 1);
    *shift = bits > kRadixBits ? bits - kRadixBits : 0;
    *hist = (u64 *)sort_temp;
    c.check(cudaMemsetAsync(sort_temp, 0, kRadixSize * sizeof(u64), c.stream));
    return true;
}

// Pairs come from (ia, va), or from `gen` when Gen::kActive.  (ib, vb): N-element scratch.
// Small problems (destination fits L2 comfortably, or few pairs) scatter directly.
// hist_ready: the digit histogram is already in `sort_temp` (scatter_hist_prepare + the producer).
template <typename Gen>
static int partitioned_scatter(Ctx &c, const Gen &gen, u32 *ia, u32 *va, u32 *ib, u32 *vb, u64 N, u64 dst_len,
                               u32 *dst, void *sort_temp, u32 *err, bool hist_ready = false)
{
    if (N == 0) return 0;
    const bool direct = scatter_is_direct(N, dst_len) || ib == nullptr || vb == nullptr;
    if (direct) {
        const int kc = c.pass_class_override >= 0 ? c.pass_class_override : KC_SCATTER;
        if (Gen::kActive) LSC_LAUNCH(c, kc, (double)N * 12, scatter_gen_kernel<Gen>, (u32)ceil_div(N, 256), 256, 0, gen, N, dst, dst_len);
        else LSC_LAUNCH(c, kc, (double)N * 12, scatter_pairs_kernel, (u32)ceil_div(N, 256), 256, 0, ia, va, N, dst, dst_len);
        return c.failed() ? -2 : 0;
    }
    const int bits = bits_for(dst_len - 1);
    const int lo = bits > kRadixBits ? bits - kRadixBits : 0;
    const int saved_class = c.pass_class_override;
    const int kc = saved_class >= 0 ? saved_class : KC_SCATTER;
    static const bool stable_env = [] { const char *e = getenv("LIBSAIS_CUDA_SCATTER_STABLE"); return e && *e && atoi(e) != 0; }();
    if (stable_env) {
        c.pass_class_override = kc;
        int where = RadixSort<u32, u32>::template sort_from<Gen>(c, gen, ia, va, ib, vb, N, lo, bits, sort_temp, err, nullptr, hist_ready);
        c.pass_class_override = saved_class;
        if (where != 1) return -2;
    } else {
        // order inside a window is irrelevant: the unstable partition pass (partition.cuh), same temp layout as the sort
        char *t = (char *)sort_temp;
        u64 *hist = (u64 *)t;               t += kMaxPasses * kRadixSize * sizeof(u64);
        u64 *base = (u64 *)t;               t += kMaxPasses * kRadixSize * sizeof(u64);
        u32 *tickets = (u32 *)t;            t += 256;
        void *status = (void *)t;
        const u64 nt = ceil_div(N, (u64)part_tile());
        const u32 dmask = (1u << (bits - lo)) - 1;
        if (!hist_ready) {
            c.check(cudaMemsetAsync(hist, 0, kRadixSize * sizeof(u64), c.stream));
            SortPlan plan = make_sort_plan(lo, bits);
            u64 want = ceil_div(N, (u64)512 * 8);
            u32 grid = (u32)(want < (u64)c.sm_count * 4 ? (want ? want : 1) : (u64)c.sm_count * 4);
            LSC_LAUNCH(c, kc, (double)N * 4.0, (sort_hist_kernel<u32, 512, Gen>), grid, 512, 0, ia, N, plan, hist, gen);
        }
        c.check(cudaMemsetAsync(tickets, 0, 256, c.stream));
        LSC_LAUNCH(c, kc, 0.0, sort_scan_kernel, 1, kRadixSize, 0, hist, base);
        c.check(cudaMemsetAsync(status, 0, nt * kRadixSize * (N < (1ull << 30) ? sizeof(u32) : sizeof(u64)), c.stream));
        const double ab = (double)N * ((Gen::kActive ? 2.0 : 8.0) + 8.0);
        PartArgs pa; pa.n = N; pa.shift = lo; pa.dmask = dmask; pa.base = base; pa.cp = nullptr; pa.nseg = 1; pa.tpc = (u32)nt;
        pa.boff = nullptr; pa.tstart = nullptr; pa.tinfo = nullptr; pa.ticket = tickets; pa.err = err; pa.use_bulk = 0; pa.kptr = nullptr; pa.vptr = nullptr;
        if constexpr (Gen::kActive) {
            FuncSrc<Gen> src; src.f = gen;
            launch_part_pass<u32, u32, FuncSrc<Gen>, false>(c, kc, ab, src, (const u32 *)nullptr, (const u32 *)nullptr, ib, vb, pa, nt, status);
        } else {
            launch_part_pass<u32, u32, ArraySrc, false>(c, kc, ab, ArraySrc(), ia, va, ib, vb, pa, nt, status);
        }
        if (c.failed()) return -2;
    }
    LSC_LAUNCH(c, kc, (double)N * 12, scatter_pairs_kernel, (u32)ceil_div(N, 256), 256, 0, ib, vb, N, dst, dst_len);
    return c.failed() ? -2 : 0;
}

}  // namespace lsc
