// po_rounds.cuh -- rounds >= 1 of the prefix doubling on a POSITION-ORDERED active list: one fused kernel per round.
//
// After round 0 the unresolved suffixes lie in slot order, group after group, and neighbouring groups have nothing
// to do with each other in the text: a tile's look-ups ISA[p + h] and its rank updates ISA[p] touch as many DRAM
// granules as it has elements (ncu, round 1: 118 bytes fetched per 4-byte look-up).  But a doubling round never
// moves anything between groups, so the ORDER OF THE GROUPS in the active list is free.  Once, after round 0, the
// groups are sorted by the text position of their head suffix (members keep their slot order).  In a repetitive
// text the group of the copies of offset x is then followed by the group of the copies of x + 1: a tile of ~20
// groups of ~100 copies reads and updates ~100 runs of ~20 consecutive ISA entries instead of 2000 scattered ones, and
// tiles taken in list order walk those runs forward through L2.  The active record is (position, rank of the
// group) -- a group occupies the slots [rank, rank + size), so slots and group boundaries need no arrays of
// their own.
//
// po_round_kernel, one tile = the groups that start in a window of the list (whole groups, <= kPoCap elements):
//   gather  r = ISA[p + h] + 1
//   order   runs of neighbours with the same r keep their relative order, so only the head of a run is ordered by counting
//           the members of its group in front of it (ties included) and behind it (smaller only), 8 lanes per head; tiles
//           with many runs order every element, a thread taking 8 consecutive members of ONE group and sharing each
//           shared-memory load between them (one integer-pipe compare + one FMA-pipe predicated add per pair)
//   rank    new sub-group heads by neighbour compare in sorted order; new rank = old rank + offset of the head
//   emit    final suffixes (sub-group of one): SA[slot], BWT row, primary index, aux sample -- each written once,
//           when it is known; changed ranks as (position, rank) pairs, grouped by text neighbourhood and staged through
//           shared memory (applied to ISA by po_apply_kernel AFTER the round: a round must read the ranks of the round
//           before); the suffixes that stay active, compacted in list order through a chained scan over the tiles (atomic
//           tickets, warp-wide decoupled look-back on one status word per tile)
// Replaces local_count / local_sort + rank_flags + rank_scan + rank_apply + the partition pass and scatter of the
// ISA update: 28 instead of ~300 bytes of DRAM traffic per active suffix and round (profiles/po_rounds_r2.md).
// (Reference counterpart: none -- libsais is SA-IS, src/libsais.c:2157-4101; this is the B200-native SA core.)
#pragma once
#include "radix_sort.cuh"

namespace lsc {

static const int kPoCap = 2048;                  // elements a tile can hold
static const int kPoThreads = 256;
static const int kPoIPT = kPoCap / kPoThreads;   // 8
static const int kPoWarps = kPoThreads / 32;
static const u32 kPoNone = 0xFFFFFFFFu;          // pair slot without an update
static const u32 kPoMaxGroup = 1024;             // larger groups after round 0: the slot-ordered path of round 1 handles the text

struct PoArgs {
    const u32 *a_pos, *a_rank;                   // the active list: position, rank of the suffix's group
    u64 N, n, h; u32 C;                          // list length, text length, sorted prefix length, window of a tile
    int bin_shift;                               // rank updates of a tile are grouped by (position >> bin_shift) & 255
    u32 ntiles;                                  // tiles of the round (= CTAs)
    u32 run_div;                                 // a tile orders run heads only when runs * run_div <= elements
    const u32 *ISA;
    u32 *o_pos, *o_rank;                         // suffixes that stay active
    u32 *pair_pos, *pair_rank;                   // [N] rank updates (kPoNone: none)
    u32 *SA; u8 *rows; const u8 *text; u64 aux_mask; int aux_shift; u32 *aux_I; u64 *primary;
    u64 *status; u32 *ticket; u64 *out_counts;   // chained scan; out_counts[0] = active suffixes, [1] = active groups after the round
    u32 *err;
};

// exclusive max-scan and sum-scan over the threads of the CTA (one barrier; the scratch must not be reused before the next barrier)
__device__ __forceinline__ void po_block_scan(u32 mymax, u32 mysum, u32 &exmax, u32 &exsum, u32 &total, u32 (*s_scan)[kPoWarps], int lane, int warp)
{
    u32 m = mymax, s = mysum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u32 om = __shfl_up_sync(0xffffffffu, m, off), os = __shfl_up_sync(0xffffffffu, s, off);
        if (lane >= off) { m = om > m ? om : m; s += os; }
    }
    if (lane == 31) { s_scan[0][warp] = m; s_scan[1][warp] = s; }
    u32 em = __shfl_up_sync(0xffffffffu, m, 1), es = __shfl_up_sync(0xffffffffu, s, 1);
    if (lane == 0) { em = 0; es = 0; }
    __syncthreads();
    u32 cm = 0, cs = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kPoWarps; ++w) {
        const u32 x = s_scan[0][w], y = s_scan[1][w];
        if (w < warp) { cm = x > cm ? x : cm; cs += y; }
        tot += y;
    }
    exmax = em > cm ? em : cm; exsum = es + cs; total = tot;       // (no trailing barrier: every call of a tile has its own scratch)
}

// c += (x <= r) / (x < r), as one compare on the integer pipe and one predicated add on the FMA pipe (float counter,
// exact below 2^24).  Left to itself the compiler spends two integer-pipe instructions and a move per pair, and the
// integer pipe is what bounds the ordering step (one warp instruction per two cycles: profiles/po_rounds_r2.md).
__device__ __forceinline__ void po_cnt_le(float &c, u32 x, u32 r)
{
    asm("{\n\t.reg .pred p;\n\tsetp.le.u32 p, %1, %2;\n\t@p add.f32 %0, %0, 0f3F800000;\n\t}" : "+f"(c) : "r"(x), "r"(r));
}
__device__ __forceinline__ void po_cnt_lt(float &c, u32 x, u32 r)
{
    asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.f32 %0, %0, 0f3F800000;\n\t}" : "+f"(c) : "r"(x), "r"(r));
}

static const int kPoItems = kPoCap / kPoIPT + kPoCap / 2;   // work items of the ordering step: <= cnt/8 + (groups <= cnt/2)

// the same with two maxima (latest group head, latest run head)
__device__ __forceinline__ void po_block_scan3(u32 max1, u32 max2, u32 mysum, u32 &ex1, u32 &ex2, u32 &exsum, u32 &total, u32 (*s_scan)[kPoWarps], int lane, int warp)
{
    u32 m1 = max1, m2 = max2, s = mysum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u32 o1 = __shfl_up_sync(0xffffffffu, m1, off), o2 = __shfl_up_sync(0xffffffffu, m2, off), os = __shfl_up_sync(0xffffffffu, s, off);
        if (lane >= off) { m1 = o1 > m1 ? o1 : m1; m2 = o2 > m2 ? o2 : m2; s += os; }
    }
    if (lane == 31) { s_scan[0][warp] = m1; s_scan[1][warp] = s; s_scan[2][warp] = m2; }
    u32 e1 = __shfl_up_sync(0xffffffffu, m1, 1), e2 = __shfl_up_sync(0xffffffffu, m2, 1), es = __shfl_up_sync(0xffffffffu, s, 1);
    if (lane == 0) { e1 = 0; e2 = 0; es = 0; }
    __syncthreads();
    u32 c1 = 0, c2 = 0, cs = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kPoWarps; ++w) {
        const u32 x = s_scan[0][w], y = s_scan[1][w], z = s_scan[2][w];
        if (w < warp) { c1 = x > c1 ? x : c1; c2 = z > c2 ? z : c2; cs += y; }
        tot += y;
    }
    ex1 = e1 > c1 ? e1 : c1; ex2 = e2 > c2 ? e2 : c2; exsum = es + cs; total = tot;
}

__global__ void __launch_bounds__(kPoThreads, 4)
po_round_kernel(const PoArgs a)
{
    __shared__ __align__(16) u32 s_r[kPoCap];    // gathered rank + 1 by list index; after the ranking: new rank by sorted index
    __shared__ __align__(16) u32 s_p[kPoCap];    // position by list index
    __shared__ __align__(16) u32 s_rk[kPoCap];   // rank of the element's group (the same for every index of a group)
    __shared__ __align__(16) unsigned short s_gs[kPoCap];   // index of the first element of the group
    __shared__ __align__(16) unsigned short s_ge[kPoCap];   // at a group's first index: one past its last; after the ranking: output offset | flags
    __shared__ __align__(16) unsigned short s_src[kPoCap];  // sorted index -> list index
    __shared__ __align__(16) unsigned short s_item[kPoItems];   // work item -> first list index of its (up to) 8 elements; later s_bin
    __shared__ __align__(16) u32 s_out[kPoCap];  // rank updates of the tile, staged in bin order (positions reuse s_rk)
    __shared__ u32 s_scan[4][3][kPoWarps];       // scratch of the four block scans of a tile
    __shared__ u32 bounds[2];
    __shared__ u32 s_tile;
    __shared__ u64 s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // Four CTAs per SM (64 registers, 47.7 KB).  Five -- positions re-read from the list instead of kept in shared memory,
    // 48 registers with a few spills -- were slower: 268 -> 282 ms over the rounds of config 3.
    // One tile per CTA.  (A persistent variant -- CTAs drawing tickets until they run out, so that a tile's stores drain under
    // the next tile's gather instead of on EXIT, where a fifth of the warp samples sit -- was slower: 58 -> 63 ms per round.)
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
    if (tid < 2) bounds[tid] = 0xFFFFFFFFu;
    __syncthreads();
    const u32 tile = s_tile;
    const u64 N = a.N, t0 = (u64)tile * a.C;

    // ---- the tile: from the first group head at or after t0 to the first at or after t0 + C
    bool ok = true;
    {
        // both ends in one loop: a step looks kPoThreads elements further behind each target; one barrier per step
        const u64 tg0 = t0, tg1 = t0 + a.C;
        if (tid == 0) { if (tg0 >= N) bounds[0] = (u32)(N - t0); if (tg1 >= N) bounds[1] = (u32)(N - t0); }
        for (u32 off = 0;; off += kPoThreads) {
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                const u64 target = which ? tg1 : tg0;
                if (target >= N) continue;
                const u64 j = target + off + tid;
                const bool hd = j < N ? (j == 0 || a.a_rank[j] != a.a_rank[j - 1]) : j == N;
                if (hd) atomicMin(&bounds[which], (u32)(j - t0));
            }
            __syncthreads();
            if (bounds[0] != 0xFFFFFFFFu && bounds[1] != 0xFFFFFFFFu) break;          // uniform: shared memory after the barrier
            if (off > (u32)kPoCap) { ok = false; break; }
            __syncthreads();                          // (rare: a second step) everyone has read bounds[] before the next atomics
        }
    }
    ok = ok && bounds[1] >= bounds[0] && bounds[1] - bounds[0] <= (u32)kPoCap;
    if (!ok && tid == 0) *a.err = 2;             // a group larger than promised
    const u32 cnt = ok ? bounds[1] - bounds[0] : 0;
    const u64 s = t0 + (ok ? bounds[0] : 0);

    u32 total = 0, act_groups = 0, nchanged = 0;
    u32 lr[kPoIPT];                              // rank of my (strided) elements inside their update bin
    u32 *s_bin = reinterpret_cast<u32 *>(s_item);
    if (cnt) {                                   // uniform over the CTA
        // ---- gather (strided: coalesced list loads, 8 look-ups in flight per thread)
        {
            u32 p[kPoIPT], rk[kPoIPT], r[kPoIPT];
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) {
                const u32 idx = i * kPoThreads + tid;
                p[i] = idx < cnt ? a.a_pos[s + idx] : 0;
                rk[i] = idx < cnt ? a.a_rank[s + idx] : 0;
            }
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) {
                const u32 idx = i * kPoThreads + tid;
                const u64 q = (u64)p[i] + a.h;
                r[i] = (idx < cnt && q < a.n) ? a.ISA[q] + 1 : 0;
            }
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) {
                const u32 idx = i * kPoThreads + tid;
                if (idx < cnt) { s_p[idx] = p[i]; s_rk[idx] = rk[i]; s_r[idx] = r[i]; }
            }
        }
        __syncthreads();
        const u32 idx0 = (u32)tid * kPoIPT;
        const int nv = idx0 < cnt ? (cnt - idx0 < (u32)kPoIPT ? (int)(cnt - idx0) : kPoIPT) : 0;
        // ---- group starts and ends, and RUNS: maximal stretches of neighbours of one group with the same gathered rank.
        // Only the head of a run has to be ordered by counting; the sorted index of a follower is its head's plus the
        // distance (equal keys keep their list order).  The groups of a repetitive text split off a few members per round,
        // so most of their elements are followers.  (blocked: a thread owns 8 consecutive elements)
        unsigned short *s_rs = reinterpret_cast<unsigned short *>(s_out);      // head of the element's run   } s_out is idle
        unsigned short *s_dh = s_rs + kPoCap;                                  // sorted index of a run head  } until the emit step
        u32 hm = 0;                              // which of my elements head a group
        u32 nrun = 0;                            // runs of the tile
        {
            u32 rk[kPoIPT], rr[kPoIPT];
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) { rk[i] = i < nv ? s_rk[idx0 + i] : 0; rr[i] = i < nv ? s_r[idx0 + i] : 0; }
            const u32 rprev = (nv && idx0 > 0) ? s_rk[idx0 - 1] : 0, rrprev = (nv && idx0 > 0) ? s_r[idx0 - 1] : 0;
            u32 last = 0, lastrun = 0, rm = 0;
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) {
                const bool hd = i < nv && (idx0 + i == 0 || rk[i] != (i ? rk[i - 1] : rprev));
                const bool rh = i < nv && (hd || rr[i] != (i ? rr[i - 1] : rrprev));
                if (hd) { hm |= 1u << i; last = idx0 + i; }
                if (rh) { rm |= 1u << i; lastrun = idx0 + i; }
            }
            u32 carry, carryrun, before;
            po_block_scan3(last, lastrun, (u32)__popc(rm), carry, carryrun, before, nrun, s_scan[0], lane, warp);
            u32 cur = carry, currun = carryrun;
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) {
                if (i < nv) {
                    const u32 idx = idx0 + i;
                    if ((hm >> i) & 1u) { if (idx > 0) s_ge[cur] = (unsigned short)idx; cur = idx; }
                    if ((rm >> i) & 1u) {
                        currun = idx;
                        const u32 k = before + (u32)__popc(rm & ((1u << i) - 1u));
                        if (k < (u32)kPoItems) s_item[k] = (unsigned short)idx;         // list of the run heads (used when they are few)
                    }
                    s_gs[idx] = (unsigned short)cur;
                    s_rs[idx] = (unsigned short)currun;
                    if (idx == cnt - 1) s_ge[cur] = (unsigned short)cnt;
                }
            }
        }
        __syncthreads();
        if (nrun * a.run_div <= cnt) {
            // ---- few runs: one thread per run head counts the members of its group in front of it (ties included) and
            // behind it (smaller ranks only)
            // (8 lanes share a run head and split its group between them: with one thread per head a few warps walked whole
            // groups while the rest of the CTA waited at the barrier)
            const u32 sub = (u32)lane & 7u;
            for (u32 k0 = 0; k0 < nrun; k0 += kPoThreads / 8) {
                const u32 k = k0 + ((u32)tid >> 3);
                u32 c = 0, e = 0, g0 = 0;
                if (k < nrun) {
                    e = s_item[k]; g0 = s_gs[e];
                    const u32 g1 = s_ge[g0], r = s_r[e];
                    for (u32 j = g0 + sub; j < g1; j += 8) {
                        const u32 x = s_r[j];
                        c += (u32)((j < e) ? (x <= r) : (j > e && x < r));
                    }
                }
                c += __shfl_xor_sync(0xffffffffu, c, 1); c += __shfl_xor_sync(0xffffffffu, c, 2); c += __shfl_xor_sync(0xffffffffu, c, 4);
                if (k < nrun && sub == 0) s_dh[e] = (unsigned short)(g0 + c);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) {
                const u32 e = i * kPoThreads + tid;
                if (e < cnt) { const u32 h0 = s_rs[e]; s_src[(u32)s_dh[h0] + (e - h0)] = (unsigned short)e; }
            }
        } else {
        // ---- work items of the ordering step: a group of s elements is cut into ceil(s / 8) runs of consecutive
        // elements, one item each, so that the 8 elements of an item always share their group (a thread that owned 8
        // consecutive elements of the TILE would straddle groups and drag its whole warp through both of them)
        u32 nitems;
        {
            u32 mine = 0;
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i)
                if ((hm >> i) & 1u) mine += ((u32)s_ge[idx0 + i] - (idx0 + i) + kPoIPT - 1) / kPoIPT;
            u32 dummy, off;
            po_block_scan(0, mine, dummy, off, nitems, s_scan[1], lane, warp);
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) {
                if ((hm >> i) & 1u) {
                    const u32 g0 = idx0 + i, g1 = s_ge[g0];
                    for (u32 e = g0; e < g1; e += kPoIPT) s_item[off++] = (unsigned short)e;
                }
            }
        }
        __syncthreads();
        // ---- order: sorted index of an element = group start + members with a smaller rank + members with the same rank in
        // front of it.  32-bit compares; every loaded rank is compared with the item's 8 elements.
        for (u32 k = tid; k < nitems; k += kPoThreads) {
            const u32 e0 = s_item[k], g0 = s_gs[e0], g1 = s_ge[g0];
            const u32 ne = g1 - e0 < (u32)kPoIPT ? g1 - e0 : (u32)kPoIPT;
            u32 r[kPoIPT]; float c[kPoIPT];
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) { r[i] = (u32)i < ne ? s_r[e0 + i] : 0xFFFFFFFFu; c[i] = 0.f; }
#pragma unroll 4
            for (u32 j = g0; j < e0; ++j) {                       // in front of the item: ties count
                const u32 x = s_r[j];
#pragma unroll
                for (int i = 0; i < kPoIPT; ++i) po_cnt_le(c[i], x, r[i]);
            }
#pragma unroll
            for (int jj = 0; jj < kPoIPT; ++jj) {                 // the item itself
#pragma unroll
                for (int i = 0; i < kPoIPT; ++i) {
                    if (jj < i) po_cnt_le(c[i], r[jj], r[i]);
                    if (jj > i) po_cnt_lt(c[i], r[jj], r[i]);
                }
            }
#pragma unroll 4
            for (u32 j = e0 + ne; j < g1; ++j) {                  // behind the item: only smaller ranks count
                const u32 x = s_r[j];
#pragma unroll
                for (int i = 0; i < kPoIPT; ++i) po_cnt_lt(c[i], x, r[i]);
            }
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) if ((u32)i < ne) s_src[g0 + (u32)c[i]] = (unsigned short)(e0 + i);
        }
        }
        __syncthreads();
        // ---- rank (blocked over the sorted order): sub-group heads, finals, new ranks, output offsets
        s_bin[tid] = 0;                          // (the items are done with)
        {
            u32 r2[kPoIPT + 1]; u32 gs[kPoIPT + 1];
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) { r2[i] = i < nv ? s_r[s_src[idx0 + i]] : 0; gs[i] = i < nv ? s_gs[idx0 + i] : 0; }
            const bool more = nv == kPoIPT && idx0 + kPoIPT < cnt;
            r2[kPoIPT] = more ? s_r[s_src[idx0 + kPoIPT]] : 0; gs[kPoIPT] = more ? s_gs[idx0 + kPoIPT] : 0;
            const u32 r2prev = (nv && idx0 > 0) ? s_r[s_src[idx0 - 1]] : 0;
            u32 nhm = 0, actm = 0, last = 0;
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) {
                if (i < nv) {
                    const u32 dd = idx0 + i;
                    const bool nh = dd == gs[i] || r2[i] != (i ? r2[i - 1] : r2prev);
                    const bool last_of_tile = dd + 1 >= cnt;
                    const bool nn = last_of_tile || gs[i + 1] == dd + 1 || r2[i + 1] != r2[i];
                    if (nh) { nhm |= 1u << i; last = dd; }
                    if (!(nh && nn)) actm |= 1u << i;
                }
            }
            u32 carry, excl;
            po_block_scan(last, (u32)__popc(actm), carry, excl, total, s_scan[2], lane, warp);   // (its barriers: every gathered rank has been read)
            u32 cur = carry;
#pragma unroll
            for (int i = 0; i < kPoIPT; ++i) {
                if (i < nv) {
                    const u32 dd = idx0 + i;
                    if ((nhm >> i) & 1u) cur = dd;
                    const bool act = (actm >> i) & 1u;
                    const u32 newrank = s_rk[dd] + (cur - gs[i]);
                    act_groups += (u32)(act && ((nhm >> i) & 1u));
                    // offset among the tile's actives (< 2048: 11 bits) | active << 14 | rank changed << 15
                    s_ge[dd] = (unsigned short)((excl + (u32)__popc(actm & ((1u << i) - 1u))) | (act ? 1u << 14 : 0u) | (cur != gs[i] ? 1u << 15 : 0u));
                    s_r[dd] = newrank;
                }
            }
        }
        __syncthreads();
        // ---- the rank updates leave the tile grouped by text neighbourhood: neighbouring groups of a repetitive text
        // update neighbouring ISA entries (copy by copy), and a warp of po_apply_kernel that holds 32 updates of one
        // neighbourhood stores into a few sectors instead of 32 (the scattered 4-byte stores, not the bytes, bound that
        // kernel: 50 G stores/s, profiles/po_rounds_r2.md).  One shared atomic per update.
#pragma unroll
        for (int i = 0; i < kPoIPT; ++i) {
            const u32 dd = i * kPoThreads + tid;
            lr[i] = 0;
            if (dd < cnt && (s_ge[dd] >> 15)) lr[i] = atomicAdd(&s_bin[(s_p[s_src[dd]] >> a.bin_shift) & 255u], 1u);
        }
        __syncthreads();
        {
            const u32 c = s_bin[tid];
            u32 dummy, ex;
            po_block_scan(0, c, dummy, ex, nchanged, s_scan[3], lane, warp);
            s_bin[tid] = ex;
        }
    }
    // ---- chained scan over the tiles: where this tile's active suffixes go
    act_groups = __reduce_add_sync(0xffffffffu, act_groups);
    if (lane == 0 && act_groups) atomicAdd((unsigned long long *)(a.out_counts + 1), (unsigned long long)act_groups);
    if (warp == 0) {
        typedef StWord<u64> W;
        if (lane == 0) st_relaxed(a.status + tile, tile == 0 ? W::inc(total) : W::agg(total));
        u64 excl = 0;
        if (tile > 0) {
            i64 look = (i64)tile - 1;
            u32 spins = 0;
            for (;;) {
                const i64 idx = look - lane;
                const u64 w = idx >= 0 ? ld_relaxed(a.status + idx) : W::inc(0);
                const u32 f = W::flag(w);
                const u32 incm = __ballot_sync(0xffffffffu, f == 2);
                const int first = incm ? __ffs(incm) - 1 : 31;                 // nearest tile with an inclusive prefix
                const u32 need = first == 31 ? 0xffffffffu : ((2u << first) - 1u);
                if (__ballot_sync(0xffffffffu, f == 0) & need) {                // someone on the way has not published yet
                    if (++spins > kSpinLimit) { if (lane == 0) *a.err = 1; break; }
                    continue;
                }
                u64 v = lane <= first ? W::val(w) : 0;
#pragma unroll
                for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                excl += v;
                if (incm) break;
                look -= 32;
            }
            if (lane == 0) st_relaxed(a.status + tile, W::inc(excl + total));
        }
        if (lane == 0) {
            s_base = excl;
            if (tile == a.ntiles - 1) a.out_counts[0] = excl + total;
        }
    }
    __syncthreads();
    if (!cnt) return;
    // ---- emit (strided over the sorted order)
    const u64 base = s_base;
#pragma unroll
    for (int i = 0; i < kPoIPT; ++i) {
        const u32 dd = i * kPoThreads + tid;
        if (dd < cnt) {
            const u32 p = s_p[s_src[dd]], meta = s_ge[dd], newrank = s_r[dd];
            const bool act = (meta >> 14) & 1u, changed = (meta >> 15) & 1u;
            if (changed) {
                const u32 slot = s_bin[(p >> a.bin_shift) & 255u] + lr[i];
                s_rk[slot] = p; s_out[slot] = newrank;           // (the group ranks in s_rk are done with)
            }
            if (act) {
                const u64 o = base + (meta & 0x3FFFu);
                a.o_pos[o] = p; a.o_rank[o] = newrank;
            } else {
                // final: a sub-group of one; its rank is its slot
                if (a.SA) a.SA[newrank] = p;
                if (p == 0) *a.primary = (u64)newrank + 1;
                else if (a.rows) a.rows[newrank] = a.text[p - 1];
                if (a.aux_I && ((u64)p & a.aux_mask) == 0) a.aux_I[p >> a.aux_shift] = newrank + 1;
            }
        }
    }
    __syncthreads();
    for (u32 idx = tid; idx < cnt; idx += kPoThreads) {           // coalesced: the tile's updates, then "none"
        const bool have = idx < nchanged;
        a.pair_pos[s + idx] = have ? s_rk[idx] : kPoNone;
        if (have) a.pair_rank[s + idx] = s_out[idx];
    }
}

// ISA[p] = rank for the pairs of the round (after the round kernel: the round reads the previous ranks).
// CTA b takes the 2048 consecutive pairs [2048 b, ...): CTAs are dispatched in index order, so the updates of neighbouring
// tiles -- neighbouring ISA entries -- are in flight together and merge in L2 (a grid-stride variant that spread a CTA's
// pairs over the whole list lost that: 26 -> 38 ms per round).  Both 16-byte loads of a quad are issued before its stores.
static __global__ void __launch_bounds__(256)
po_apply_kernel(const u32 *__restrict__ pair_pos, const u32 *__restrict__ pair_rank, u64 N, u32 *__restrict__ ISA)
{
    const u64 base = (u64)blockIdx.x * 2048;
    const bool vec = (((uintptr_t)pair_pos | (uintptr_t)pair_rank) & 15) == 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const u64 j = base + ((u64)k * 256 + threadIdx.x) * 4;
        if (j >= N) return;
        if (vec && j + 4 <= N) {
            const uint4 p = *reinterpret_cast<const uint4 *>(pair_pos + j), r = *reinterpret_cast<const uint4 *>(pair_rank + j);
            if (p.x != kPoNone) ISA[p.x] = r.x;
            if (p.y != kPoNone) ISA[p.y] = r.y;
            if (p.z != kPoNone) ISA[p.z] = r.z;
            if (p.w != kPoNone) ISA[p.w] = r.w;
        } else {
            for (u64 i = j; i < j + 4 && i < N; ++i) {
                const u32 p = pair_pos[i];
                if (p != kPoNone) ISA[p] = pair_rank[i];
            }
        }
    }
}

// ---- the one-time reordering of the GROUPS by the text position of their head suffix.  Whole groups move, so the
// groups are sorted (one (head position, group) pair per group -- 1/100 of the suffixes in a 100-copy collection),
// their new starts are the exclusive scan of their sizes in sorted order, and one pass copies every member to
// new start + offset in the group.  (A sort of the member records themselves cost 4 digit passes over 12-byte
// records: 87 ms of config 3's 567.)
// table of the groups from the slot-ordered active list of round 0 (dense, ascending group ids): first list index,
// head position (the sort key), head slot (= rank of the group), identity (the sort's value)
static __global__ void __launch_bounds__(256)
po_group_table_kernel(const u32 *__restrict__ a_pos, const u32 *__restrict__ a_grp, const u32 *__restrict__ a_slot, u64 N,
                      u32 *__restrict__ gstart, u32 *__restrict__ ghead_pos, u32 *__restrict__ ghead_slot, u32 *__restrict__ gid)
{
    const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j >= N) return;
    const u32 g = a_grp[j];
    if (j == 0 || a_grp[j - 1] != g) { gstart[g] = (u32)j; ghead_pos[g] = a_pos[j]; ghead_slot[g] = a_slot[j]; gid[g] = g; }
    if (j == N - 1) gstart[g + 1] = (u32)N;
}

static const int kPoScanThreads = 512;
static const int kPoScanChunk = kPoScanThreads * 8;
// sum of the sizes of the groups sg[i], i in chunk b
static __global__ void __launch_bounds__(kPoScanThreads)
po_scan_sums_kernel(const u32 *__restrict__ sg, const u32 *__restrict__ gstart, u64 G, u32 *__restrict__ bsum)
{
    __shared__ u32 s_w[kPoScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 base = (u64)blockIdx.x * kPoScanChunk;
    u32 sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const u64 i = base + (u64)k * kPoScanThreads + tid;
        if (i < G) { const u32 g = sg[i]; sum += gstart[g + 1] - gstart[g]; }
    }
    sum = __reduce_add_sync(0xffffffffu, sum);
    if (lane == 0) s_w[warp] = sum;
    __syncthreads();
    if (tid == 0) { u32 t = 0; for (int w = 0; w < kPoScanThreads / 32; ++w) t += s_w[w]; bsum[blockIdx.x] = t; }
}
// one CTA: exclusive scan of the chunk sums in place
static __global__ void __launch_bounds__(1024)
po_scan_top_kernel(u32 *__restrict__ bsum, u64 nb)
{
    __shared__ u32 s_w[32];
    __shared__ u32 s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (u64 base = 0; base < nb; base += 1024) {
        const u64 i = base + tid;
        const u32 v = i < nb ? bsum[i] : 0;
        u32 x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        u32 before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (i < nb) bsum[i] = before + x - v;
        __syncthreads();
        if (tid == 1023) s_carry = before + x;
        __syncthreads();
    }
}
// newstart[sg[i]] = exclusive prefix of the sizes in sorted order (thread: 8 consecutive groups)
static __global__ void __launch_bounds__(kPoScanThreads)
po_scan_apply_kernel(const u32 *__restrict__ sg, const u32 *__restrict__ gstart, u64 G, const u32 *__restrict__ bsum, u32 *__restrict__ newstart)
{
    __shared__ u32 s_w[kPoScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 i0 = (u64)blockIdx.x * kPoScanChunk + (u64)tid * 8;
    u32 g[8], sz[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        g[k] = 0; sz[k] = 0;
        if (i0 + k < G) { g[k] = sg[i0 + k]; sz[k] = gstart[g[k] + 1] - gstart[g[k]]; }
        sum += sz[k];
    }
    u32 x = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    u32 run = bsum[blockIdx.x] + x - sum;
    for (int w = 0; w < warp; ++w) run += s_w[w];
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (i0 + k < G) newstart[g[k]] = run; run += sz[k]; }
}
// every member to its group's new start + its offset in the group; the record becomes (position, rank of the group)
static __global__ void __launch_bounds__(256)
po_move_kernel(const u32 *__restrict__ a_pos, const u32 *__restrict__ a_grp, u64 N, const u32 *__restrict__ gstart,
               const u32 *__restrict__ newstart, const u32 *__restrict__ ghead_slot, u32 *__restrict__ o_pos, u32 *__restrict__ o_rank)
{
    const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j >= N) return;
    const u32 g = a_grp[j];
    const u64 dst = (u64)newstart[g] + (j - (u64)gstart[g]);
    o_pos[dst] = a_pos[j];
    o_rank[dst] = ghead_slot[g];
}

}  // namespace lsc
