// sa_core.cu -- suffix array + inverse suffix array of a device-resident text by prefix
// doubling with discarding.  This is the B200-native replacement of the reference's SA-IS
// core (reference src/libsais.c:6879-6923 libsais_main_8u and :6668-6877
// libsais_main_32s_recursion); it computes the same, unique, SA.
//
// Pipeline:
//   hist_sym      byte histogram (freq[], alphabet map)                      [~ :1398-1427]
//   pack          symbols -> order-preserving b-bit codes, big-endian bitstream
//   round 0       key(p) = first k codes of suffix p (K = k*b <= 64 bits), elements in DESCENDING p.
//                 The key array is never materialised: KmerGen feeds the first onesweep pass, and the
//                 digit histograms of all passes come from the text's s-gram / byte histogram.
//   onesweep      stable LSD radix sort of (key, p)                          radix_sort.cuh
//   rank stage    rank_flags -> rank_scan -> rank_apply: head flags, rank = slot of the group head,
//                 SA / BWT rows / primary / aux outputs, compaction of the suffixes whose group is
//                 not yet a singleton ("active"); ranks reach ISA through the locality-partitioned
//                 scatter (scatter.cuh) -- or not at all when few suffixes stay active (lazy ISA)
//   repeat while active: keys (g, ISA[p+h]+1), sorted -> rank stage; h doubles.  Rounds whose groups are
//                 all small gather and sort them inside shared memory (local_sort.cuh: counting rank or
//                 in-tile radix passes); the others run round_keys (+ digit histograms) -> onesweep
//   (bottom of the file: building blocks of the distributed variant, libsais_b200/dist.py)
//
// End-of-text rule ("a suffix that is a prefix of another sorts first", reference
// include/libsais.h:76-84): keys are zero padded, the initial sort is stable over elements
// laid out in descending position, and the < k suffixes that run past the end are forced to
// be singleton groups -- so a short suffix precedes every longer suffix sharing its padded
// key, exactly as an implicit smallest sentinel would order them.
#include "core.h"
#include "radix_sort.cuh"
#include "scatter.cuh"
#include "lazy_rank.cuh"
#include "local_sort.cuh"
#include "partition.cuh"
#include "po_rounds.cuh"
#include <cmath>
#include <cstdlib>

namespace lsc {

// ---------------------------------------------------------------------------------------------
// symbol statistics
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
byte_hist_kernel(const u8 *__restrict__ T, u64 n, u64 *__restrict__ freq)
{
    __shared__ u32 sh[8][256];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    // body: 16 bytes per thread per step (T from cudaMalloc / arena: 16-B aligned), tail: bytes
    const u64 nvec = ((uintptr_t)T & 15) == 0 ? n / 16 : 0;
    const uint4 *T4 = reinterpret_cast<const uint4 *>(T);
    for (u64 v = (u64)blockIdx.x * 256 + tid; v < nvec; v += (u64)gridDim.x * 256) {
        uint4 q = T4[v];
        u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&sh[warp][w[j] & 255], 1u);
            atomicAdd(&sh[warp][(w[j] >> 8) & 255], 1u);
            atomicAdd(&sh[warp][(w[j] >> 16) & 255], 1u);
            atomicAdd(&sh[warp][w[j] >> 24], 1u);
        }
    }
    for (u64 i = nvec * 16 + (u64)blockIdx.x * 256 + tid; i < n; i += (u64)gridDim.x * 256)
        atomicAdd(&sh[warp][T[i]], 1u);
    __syncthreads();
    u32 s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w][tid];
    if (s) atomicAdd((unsigned long long *)&freq[tid], (unsigned long long)s);
}

void run_byte_histogram(Ctx &c, const u8 *d_T, u64 n)
{
    c.check(cudaMemsetAsync(c.d_scalars + S_FREQ, 0, 256 * sizeof(u64), c.stream));
    if (n == 0) return;
    u64 want = ceil_div(n, 256 * 64);
    u32 grid = (u32)(want < (u64)c.sm_count * 8 ? want : (u64)c.sm_count * 8);
    LSC_LAUNCH(c, KC_HIST_SYM, (double)n, byte_hist_kernel, grid, 256, 0, d_T, n, c.d_scalars + S_FREQ);
}

template <typename SymT>
__global__ void __launch_bounds__(256)
max_sym_kernel(const SymT *__restrict__ T, u64 n, u64 *__restrict__ out)
{
    u64 m = 0;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
        u64 v = (u64)T[i];
        m = v > m ? v : m;
    }
    for (int off = 16; off; off >>= 1) { u64 o = __shfl_xor_sync(0xffffffffu, m, off); m = o > m ? o : m; }
    if ((threadIdx.x & 31) == 0 && m) atomicMax((unsigned long long *)out, (unsigned long long)m);
}

// ---------------------------------------------------------------------------------------------
// pack: word j of the stream holds bits [64j, 64j+64); symbol s occupies bits [s*b, s*b+b),
// most significant bit first, so a 64-bit window at bit p*b is the big-endian k-mer of suffix p.
// ---------------------------------------------------------------------------------------------
template <typename SymT, bool USE_LUT>
__global__ void __launch_bounds__(256)
pack_kernel(const SymT *__restrict__ T, u64 n, int b, u64 *__restrict__ words, u64 nwords,
            const u8 *__restrict__ lut)
{
    u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j >= nwords) return;
    const u64 bit0 = j * 64;
    u64 w = 0;
    for (u64 s = bit0 / (u64)b;; ++s) {
        u64 sb = s * (u64)b;
        if (sb >= bit0 + 64) break;
        u64 code = 0;
        if (s < n) code = USE_LUT ? (u64)lut[(u32)T[s] & 255] : (u64)T[s];
        i64 sh = (i64)(bit0 + 64) - (i64)(sb + (u64)b);
        w |= sh >= 0 ? (code << sh) : (code >> (-sh));
    }
    words[j] = w;
}

// Byte texts whose code width divides 8 (b = 1, 2, 4, 8: every word is a whole number of bytes of
// text): 8-byte loads, codes from a shared-memory table (IDENT: the table is the identity -- all
// 256 byte values occur -- and a word is just 8 text bytes in big-endian order).
template <int B, bool IDENT>
__global__ void __launch_bounds__(256)
pack_bytes_kernel(const u8 *__restrict__ T, u64 n, u64 *__restrict__ words, u64 nwords, const u8 *__restrict__ lut)
{
    __shared__ u8 s_lut[256];
    if (!IDENT) { s_lut[threadIdx.x] = lut[threadIdx.x]; __syncthreads(); }
    const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j >= nwords) return;
    constexpr int SPW = 64 / B;                    // symbols (= text bytes) per word
    const u64 s0 = j * SPW;
    u64 w = 0;
    if (s0 + SPW <= n && ((uintptr_t)T & 7) == 0) {
        const u64 *T8 = reinterpret_cast<const u64 *>(T + s0);
#pragma unroll
        for (int q = 0; q < SPW / 8; ++q) {
            const u64 x = T8[q];
            if (IDENT) {
                w = (u64)__byte_perm((u32)x, 0, 0x0123) << 32 | (u64)__byte_perm((u32)(x >> 32), 0, 0x0123);
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) w = (w << B) | (u64)s_lut[(u32)(x >> (8 * t)) & 255];
            }
        }
    } else {
        for (int t = 0; t < SPW; ++t) {
            const u64 sidx = s0 + t;
            const u64 code = sidx < n ? (IDENT ? (u64)T[sidx] : (u64)s_lut[T[sidx]]) : 0;
            w = (w << B) | code;
        }
    }
    words[j] = w;
}

// Round-0 element generator: element i <-> position p = n-1-i (descending positions: see the
// end-of-text rule above), key = k-mer of suffix p (<< key_shift, | preceding byte in BWT mode),
// value = p.  Evaluated by the histogram kernel and by the first digit pass, so the initial
// (key, position) array is never written or read (saves 30 of ~150 bytes of traffic per suffix).
struct KmerGen {
    static const bool kActive = true;
    const u64 *words; const u8 *text; u64 n; int b, K, key_shift;
    __device__ __forceinline__ u64 key(u64 i) const
    {
        u64 p = n - 1 - i;
        u64 k = kmer_at(words, p, b, K) << key_shift;
        if (text != nullptr && p > 0) k |= (u64)text[p - 1];
        return k;
    }
    __device__ __forceinline__ u32 val(u64 i) const { return (u32)(n - 1 - i); }
};

// Materialising variant (LIBSAIS_CUDA_FUSE_KEYS=0), kept for A/B measurements.
// element i <-> position p = n-1-i.
// BWT mode (text != nullptr): the byte preceding the suffix rides in the low 8 key bits,
// below the sorted bit range, so the BWT falls out of the sort without a gather.
__global__ void __launch_bounds__(256)
make_keys_kernel(const u64 *__restrict__ words, u64 n, int b, int K, int key_shift,
                 const u8 *__restrict__ text, u64 *__restrict__ keys, u32 *__restrict__ pos)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    u64 p = n - 1 - i;
    u64 key = kmer_at(words, p, b, K) << key_shift;
    if (text != nullptr && p > 0) key |= (u64)text[p - 1];
    keys[i] = key;
    pos[i] = (u32)p;
}

// ---------------------------------------------------------------------------------------------
// rank stage: sorted (key, pos) -> head flags, rank (= slot of the group head), SA / ISA
// scatter, compaction of non-singleton ("active") suffixes with their slot and dense group id,
// and -- for suffixes that became singletons ("final") -- the fused outputs: BWT row byte,
// primary index, aux samples.  Three kernels, no inter-CTA waiting:
//   rank_flags  streaming, 4 consecutive elements per thread (16-byte loads; neighbour compares stay
//               in the thread, one shuffle per thread for the element before / after): head and
//               active bits (one byte per thread), per-warp and per-tile aggregates, and everything
//               that needs no prefix: SA[slot], BWT rows, primary, aux samples
//   rank_scan   three CTAs: exclusive scan of the tile aggregates
//               [0] max : slot of the last group head   [1] sum : active suffixes   [2] sum : active groups
//   rank_apply  streaming over the tiles that contain work: rank -> ISA, compaction of the actives
// ---------------------------------------------------------------------------------------------
static const int kRankThreads = 256;
static const int kRankIPT = 4;
static const int kRankTile = kRankThreads * kRankIPT;
static const int kRankWarps = kRankThreads / 32;
__device__ __forceinline__ u64 ceil_div_dev(u64 a, u64 b) { return (a + b - 1) / b; }

struct RankArgs {
    const u64 *keys; const u32 *pos; const u32 *slot_in;
    u64 N; u64 tail_start; int key_shift;
    u32 slot_base;           // global slot of element 0 (0 on one GPU; a rank's offset in the distributed sort)
    u32 *SA;                 // nullable: SA[slot - slot_base] = pos
    u32 *ISA; int isa_all;   // isa_all: write every element's rank, else only the active ones
    u32 *pair_idx, *pair_val; // non-null: emit (position, rank) pairs in element order instead of scattering into ISA
    u64 *phist; int pshift;   // non-null: histogram of (position >> pshift) & 255 over the pairs, for the partition pass of the scatter
    u8 *rows; const u8 *text;            // BWT mode: rows[slot] = byte preceding the suffix
    u64 aux_mask; int aux_shift; u32 *aux_I;
    u64 *primary;
    u32 *a_pos, *a_slot, *a_grp;
    u32 *masks;              // one byte per 4 elements: head bits | active bits << 4
    u32 *wagg;               // [ntiles * kRankWarps][3]  per-warp aggregates
    u32 *tagg;               // [3][ntiles] per-tile aggregates, overwritten by their exclusive scan
    u64 nchunks;             // ceil(N / 32)
    u64 ntiles; u64 *out_counts;
};

// 4 consecutive elements of an array: one 16-byte access per 4 words when the address allows it
__device__ __forceinline__ void load4(const u64 *__restrict__ p, u64 j0, int nv, u64 (&o)[4])
{
    if (nv == 4 && ((uintptr_t)(p + j0) & 15) == 0) {
        const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(p + j0), y = *reinterpret_cast<const ulonglong2 *>(p + j0 + 2);
        o[0] = x.x; o[1] = x.y; o[2] = y.x; o[3] = y.y;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = i < nv ? p[j0 + i] : 0;
    }
}
__device__ __forceinline__ void load4(const u32 *__restrict__ p, u64 j0, int nv, u32 (&o)[4])
{
    if (nv == 4 && ((uintptr_t)(p + j0) & 15) == 0) {
        const uint4 x = *reinterpret_cast<const uint4 *>(p + j0);
        o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = i < nv ? p[j0 + i] : 0;
    }
}
__device__ __forceinline__ void store4(u32 *__restrict__ p, u64 j0, int nv, const u32 (&v)[4])
{
    if (nv == 4 && ((uintptr_t)(p + j0) & 15) == 0) {
        *reinterpret_cast<uint4 *>(p + j0) = make_uint4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i < nv) p[j0 + i] = v[i];
    }
}
// element of a 4-array picked by the highest set bit of a nibble (selects, no local memory)
__device__ __forceinline__ u32 top_of(u32 nib, const u32 (&v)[4])
{
    return (nib & 8u) ? v[3] : (nib & 4u) ? v[2] : (nib & 2u) ? v[1] : (nib & 1u) ? v[0] : 0u;
}

template <bool ROUND0>
__global__ void __launch_bounds__(kRankThreads)
rank_flags_kernel(const RankArgs a)
{
    __shared__ u32 s_wagg[3][kRankWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 N = a.N;
    const u64 tile = blockIdx.x;
    const u64 j0 = tile * kRankTile + (u64)tid * 4, jn = j0 + 4;       // my elements: [j0, j0 + nv)
    const int nv = j0 < N ? (N - j0 < 4 ? (int)(N - j0) : 4) : 0;
    const u32 vb = (1u << nv) - 1u;

    u64 kk[4]; u32 p[4], slot[4];
    load4(a.keys, j0, nv, kk);
    load4(a.pos, j0, nv, p);
    if (ROUND0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) slot[i] = a.slot_base + (u32)j0 + (u32)i;
    } else load4(a.slot_in, j0, nv, slot);
    // the elements just outside the warp's 128: lane 0 reads element j0 - 1, lane 31 element j0 + 4
    u64 kbefore = 0, kafter = 0; u32 pbefore = 0, pafter = 0;
    if (lane == 0 && j0 > 0 && nv > 0) { kbefore = a.keys[j0 - 1] >> a.key_shift; if (ROUND0) pbefore = a.pos[j0 - 1]; }
    if (lane == 31 && jn < N) { kafter = a.keys[jn] >> a.key_shift; if (ROUND0) pafter = a.pos[jn]; }

    u32 prevbytes = 0;
    if (ROUND0) prevbytes = (u32)(kk[0] & 255) | (u32)(kk[1] & 255) << 8 | (u32)(kk[2] & 255) << 16 | (u32)(kk[3] & 255) << 24;
#pragma unroll
    for (int i = 0; i < 4; ++i) kk[i] >>= a.key_shift;
    u64 kprev = __shfl_up_sync(0xffffffffu, kk[3], 1);
    if (lane == 0) kprev = kbefore;
    u32 h = (u32)(j0 == 0 || kprev != kk[0]) | (u32)(kk[0] != kk[1]) << 1 | (u32)(kk[1] != kk[2]) << 2 | (u32)(kk[2] != kk[3]) << 3;
    h &= vb;
    u32 t = 0;
    if (ROUND0) {
        // tail suffixes are singletons: they are heads, and so is the element after one
        t = ((u32)((u64)p[0] >= a.tail_start) | (u32)((u64)p[1] >= a.tail_start) << 1
             | (u32)((u64)p[2] >= a.tail_start) << 2 | (u32)((u64)p[3] >= a.tail_start) << 3) & vb;
        u32 tprev = __shfl_up_sync(0xffffffffu, t >> 3, 1);
        if (lane == 0) tprev = (u32)(j0 > 0 && nv > 0 && (u64)pbefore >= a.tail_start);
        h |= t | (((t << 1) | tprev) & vb);
    }
    // head flag of the element after my last one; an element beyond N counts as a head
    u32 hnext = __shfl_down_sync(0xffffffffu, (h & 1u) | (u32)(nv == 0), 1);
    if (lane == 31) {
        if (jn >= N) hnext = 1;
        else {
            hnext = (u32)(kafter != kk[3]);
            if (ROUND0) hnext |= (u32)((u64)pafter >= a.tail_start) | (t >> 3);
        }
    }
    const u32 nh = (((h | ~vb) >> 1) & 7u) | (hnext << 3);
    const u32 am = vb & ~(h & nh);          // active: not (head followed by a head)
    const u32 gm = am & h;                  // heads of active groups

    if (nv) {
        reinterpret_cast<u8 *>(a.masks)[j0 >> 2] = (u8)(h | (am << 4));
        if (a.SA) {
            if (ROUND0) store4(a.SA, j0, nv, p);
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (i < nv) a.SA[slot[i] - a.slot_base] = p[i];
            }
        }
        if (a.rows) {
            if (ROUND0) {
                u8 *r = a.rows + slot[0];
                if (nv == 4 && ((uintptr_t)r & 3) == 0) *reinterpret_cast<u32 *>(r) = prevbytes;
                else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (i < nv) r[i] = (u8)(prevbytes >> (8 * i));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (i < nv && !((am >> i) & 1u) && p[i] != 0) a.rows[slot[i]] = a.text[p[i] - 1];
            }
        }
        const u32 fin = vb & ~am;           // final: the slot of this suffix will never change again
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if ((fin >> i) & 1u) {
                if (p[i] == 0) *a.primary = (u64)slot[i] + 1;
                if (a.aux_I && ((u64)p[i] & a.aux_mask) == 0) a.aux_I[p[i] >> a.aux_shift] = slot[i] + 1;
            }
        }
    }
    // slots ascend with the element index: the latest head of the warp is the largest head slot
    const u32 w_head = __reduce_max_sync(0xffffffffu, top_of(h, slot));
    const u32 w_act = __reduce_add_sync(0xffffffffu, (u32)__popc(am));
    const u32 w_grp = __reduce_add_sync(0xffffffffu, (u32)__popc(gm));
    if (lane == 0) {
        u32 *wa = a.wagg + (tile * kRankWarps + warp) * 3;
        wa[0] = w_head; wa[1] = w_act; wa[2] = w_grp;
        s_wagg[0][warp] = w_head; s_wagg[1][warp] = w_act; s_wagg[2][warp] = w_grp;
    }
    __syncthreads();
    if (tid < 3) {
        u32 r = 0;
#pragma unroll
        for (int w = 0; w < kRankWarps; ++w) { u32 v = s_wagg[tid][w]; r = tid == 0 ? (v > r ? v : r) : r + v; }
        a.tagg[(u64)tid * a.ntiles + tile] = r;
    }
}

// Round 0 of the fused MSD path: bucket_sort_kernel (partition.cuh, BucketFuse) left one flag byte per slot (bit 0 head,
// bit 1 active).  Packs them into rank_flags' mask bytes (heads | actives << 4 per 4 slots) and takes the per-warp and
// per-tile aggregates: slot of the last head, active suffixes, active groups.
static const int kAggTiles = 8;          // tiles per CTA of rank_agg_kernel (one tile per CTA: 262 144 CTAs for 2^28 slots, 1.26 TB/s)
__global__ void __launch_bounds__(kRankThreads)
rank_agg_kernel(const u8 *__restrict__ flags, u32 *__restrict__ masks, u64 N, u32 *__restrict__ wagg, u32 *__restrict__ tagg, u64 ntiles)
{
    __shared__ u32 s_wagg[kAggTiles][3][kRankWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tile0 = (u64)blockIdx.x * kAggTiles;
    const bool vec = ((uintptr_t)flags & 3) == 0;
    u32 f4[kAggTiles];
#pragma unroll
    for (int t = 0; t < kAggTiles; ++t) {                  // all loads first
        const u64 j0 = (tile0 + t) * kRankTile + (u64)tid * 4;
        f4[t] = 0;
        if (tile0 + t < ntiles) {
            if (vec && j0 + 4 <= N) f4[t] = *reinterpret_cast<const u32 *>(flags + j0);
            else { for (int i = 0; i < 4; ++i) if (j0 + i < N) f4[t] |= (u32)flags[j0 + i] << (8 * i); }
        }
    }
#pragma unroll
    for (int t = 0; t < kAggTiles; ++t) {
        const u64 tile = tile0 + t;
        if (tile >= ntiles) break;                         // uniform
        const u64 j0 = tile * kRankTile + (u64)tid * 4;
        const u32 f = f4[t];
        const u32 h = (f & 1u) | ((f >> 7) & 2u) | ((f >> 14) & 4u) | ((f >> 21) & 8u);
        const u32 am = ((f >> 1) & 1u) | ((f >> 8) & 2u) | ((f >> 15) & 4u) | ((f >> 22) & 8u);
        const u32 gm = am & h;
        if (j0 < N) reinterpret_cast<u8 *>(masks)[j0 >> 2] = (u8)(h | (am << 4));
        const u32 w_head = __reduce_max_sync(0xffffffffu, h ? (u32)j0 + (u32)(31 - __clz(h)) : 0u);
        const u32 w_act = __reduce_add_sync(0xffffffffu, (u32)__popc(am));
        const u32 w_grp = __reduce_add_sync(0xffffffffu, (u32)__popc(gm));
        if (lane == 0) {
            u32 *wa = wagg + (tile * kRankWarps + warp) * 3;
            wa[0] = w_head; wa[1] = w_act; wa[2] = w_grp;
            s_wagg[t][0][warp] = w_head; s_wagg[t][1][warp] = w_act; s_wagg[t][2][warp] = w_grp;
        }
    }
    __syncthreads();
    if (tid < 3 * kAggTiles) {
        const int t = tid / 3, q = tid % 3;
        if (tile0 + t < ntiles) {
            u32 r = 0;
#pragma unroll
            for (int w = 0; w < kRankWarps; ++w) { u32 v = s_wagg[t][q][w]; r = q == 0 ? (v > r ? v : r) : r + v; }
            tagg[(u64)q * ntiles + tile0 + t] = r;
        }
    }
}

// Three CTAs (one per aggregate): exclusive scan of the tile aggregates in place (SoA:
// tagg[q * ntiles + tile]); totals -> out_counts[0..1].  Each warp owns a contiguous segment of
// tiles and walks it 32 tiles at a time (coalesced loads, shuffle scan, running carry): reduce,
// combine, re-walk.
__device__ __forceinline__ u32 warp_incl_scan(u32 v, bool is_max, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        u32 o = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v = is_max ? (o > v ? o : v) : v + o;
    }
    return v;
}

static const int kRankScanCtas = 3;
__global__ void __launch_bounds__(1024)
rank_scan_kernel(u32 *__restrict__ tagg, u64 ntiles, u64 *__restrict__ out_counts)
{
    __shared__ u32 s_tot[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int q = blockIdx.x;
    const bool is_max = q == 0;
    u32 *v = tagg + (u64)q * ntiles;
    const u64 per = ((ntiles + 31) / 32 + 31) / 32 * 32;          // tiles per warp, multiple of 32
    const u64 lo = (u64)warp * per, hi = lo + per < ntiles ? lo + per : ntiles;
    {
        u32 r = 0;
        for (u64 i = lo + lane; i < hi; i += 32) { u32 x = v[i]; r = is_max ? (x > r ? x : r) : r + x; }
#pragma unroll
        for (int off = 16; off; off >>= 1) { u32 o = __shfl_xor_sync(0xffffffffu, r, off); r = is_max ? (o > r ? o : r) : r + o; }
        if (lane == 0) s_tot[warp] = r;
    }
    __syncthreads();
    if (warp == 0) {
        u32 x = s_tot[lane];
        u32 inc = warp_incl_scan(x, is_max, lane);
        u32 ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = 0;
        s_tot[lane] = ex;
        if (lane == 31 && !is_max) out_counts[q - 1] = inc;
    }
    __syncthreads();
    u32 carry = s_tot[warp];
    for (u64 base = lo; base < hi; base += 32) {
        const u64 i = base + lane;
        u32 x = i < hi ? v[i] : 0;
        u32 inc = warp_incl_scan(x, is_max, lane);
        u32 ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = 0;
        ex = is_max ? (ex > carry ? ex : carry) : ex + carry;
        if (i < hi) v[i] = ex;
        u32 tot = __shfl_sync(0xffffffffu, inc, 31);
        carry = is_max ? (tot > carry ? tot : carry) : carry + tot;
    }
}

void run_rank_scan(Ctx &c, u32 *tagg, u64 ntiles, u64 *out_counts)
{
    LSC_LAUNCH(c, KC_RANK_SCAN, (double)ntiles * 24, rank_scan_kernel, kRankScanCtas, 1024, 0, tagg, ntiles, out_counts);
}

static const int kApplyTiles = 4;
template <bool ROUND0>
__global__ void __launch_bounds__(kRankThreads)
rank_apply_kernel(const RankArgs a)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 N = a.N;
    __shared__ u32 s_ph[kRadixSize];
    const bool ph = a.phist != nullptr && a.pair_idx != nullptr;
    if (ph) { s_ph[tid] = 0; __syncthreads(); }
    // a CTA walks kApplyTiles consecutive tiles (no block-level synchronisation inside: warps are independent)
    for (u64 tile = (u64)blockIdx.x * kApplyTiles; tile < a.ntiles && tile < ((u64)blockIdx.x + 1) * kApplyTiles; ++tile) {
    // tiles without active suffixes have nothing to do unless every rank is wanted
    u32 t_head = a.tagg[tile], t_act = a.tagg[a.ntiles + tile], t_grp = a.tagg[2 * a.ntiles + tile];
    if (!a.isa_all) {
        const u32 next_act = tile + 1 < a.ntiles ? a.tagg[a.ntiles + tile + 1] : (u32)a.out_counts[0];
        if (next_act == t_act) continue;
    }
    // exclusive prefix of this warp inside the tile from the per-warp aggregates
    u32 c_head = t_head, c_act = t_act, c_grp = t_grp;
    {
        const u32 *wa = a.wagg + tile * kRankWarps * 3;
        u32 h = lane < warp ? wa[lane * 3] : 0, x = lane < warp ? wa[lane * 3 + 1] : 0, g = lane < warp ? wa[lane * 3 + 2] : 0;
#pragma unroll
        for (int off = 4; off; off >>= 1) {
            u32 oh = __shfl_xor_sync(0xffffffffu, h, off); h = oh > h ? oh : h;
            x += __shfl_xor_sync(0xffffffffu, x, off);
            g += __shfl_xor_sync(0xffffffffu, g, off);
        }
        h = __shfl_sync(0xffffffffu, h, 0); x = __shfl_sync(0xffffffffu, x, 0); g = __shfl_sync(0xffffffffu, g, 0);
        c_head = h > c_head ? h : c_head; c_act += x; c_grp += g;
    }
    const u64 j0 = tile * kRankTile + (u64)tid * 4;
    const int nv = j0 < N ? (N - j0 < 4 ? (int)(N - j0) : 4) : 0;
    const u32 mb = nv ? reinterpret_cast<const u8 *>(a.masks)[j0 >> 2] : 0u;
    const u32 h = mb & 15u, am = mb >> 4, gm = am & h;
    u32 p[4], slot[4];
    load4(a.pos, j0, nv, p);
    if (ROUND0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) slot[i] = a.slot_base + (u32)j0 + (u32)i;
    } else load4(a.slot_in, j0, nv, slot);
    // actives / active groups in the lanes before mine (counts <= 128: both in one word)
    const u32 mine = (u32)__popc(am) | (u32)__popc(gm) << 16;
    u32 inc = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { u32 o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
    const u32 e_act = c_act + ((inc - mine) & 0xFFFFu), e_grp = c_grp + ((inc - mine) >> 16);
    // rank = slot of the nearest head at or before the element: in my nibble, else in the nearest lane below
    // that has one, else carried into the warp
    const u32 hb = __ballot_sync(0xffffffffu, h != 0) & lanemask_lt();
    const u32 hs = __shfl_sync(0xffffffffu, top_of(h, slot), hb ? 31 - __clz(hb) : 0);
    u32 cur = hb ? hs : c_head;
    u32 rk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { if ((h >> i) & 1u) cur = slot[i]; rk[i] = cur; }
    if (nv) {
        if (a.pair_idx) {
            store4(a.pair_idx, j0, nv, p); store4(a.pair_val, j0, nv, rk);
            if (ph) {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (i < nv) atomicAdd(&s_ph[(p[i] >> a.pshift) & 255u], 1u);
            }
        }
        else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (i < nv && (a.isa_all || ((am >> i) & 1u))) a.ISA[p[i]] = rk[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if ((am >> i) & 1u) {
                const u32 o = e_act + (u32)__popc(am & ((1u << i) - 1u));
                a.a_pos[o] = p[i];
                a.a_slot[o] = slot[i];
                a.a_grp[o] = e_grp + (u32)__popc(gm & ((2u << i) - 1u)) - 1;
            }
        }
    }
    }
    if (ph) {
        __syncthreads();
        const u32 v = s_ph[tid];
        if (v) atomicAdd((unsigned long long *)&a.phist[tid], (unsigned long long)v);
    }
}

// round >= 1 keys: (dense group id, rank of the suffix h positions further + 1).  The ISA gather
// leaves the issue slots idle, so the digit histograms of the sort that follows are taken here,
// while the keys are in registers (hist[pass][256], zeroed by the caller): the sort needs no
// histogram pass of its own.  Grid-stride over tiles of 1024 elements, 4 gathers in flight per thread.
static const int kKeyPasses = 8;        // a 64-bit key has at most 8 digits
template <bool LAZY>
__global__ void __launch_bounds__(256)
round_keys_kernel(const u32 *__restrict__ a_pos, const u32 *__restrict__ a_grp,
                  const u32 *__restrict__ ISA, u64 N, u64 n, u64 h, int rank_bits,
                  u64 *__restrict__ keys, u32 *__restrict__ pos, const LazyArgs la,
                  const SortPlan plan, u64 *__restrict__ hist)
{
    __shared__ u32 sh[kKeyPasses * kRadixSize];
    const int tid = threadIdx.x, lane = tid & 31, P = plan.passes;
    for (int i = tid; i < P * kRadixSize; i += 256) sh[i] = 0;
    __syncthreads();
    for (u64 base = (u64)blockIdx.x * 1024; base < N; base += (u64)gridDim.x * 1024) {
        u32 p[4], g[4], r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u64 j = base + (u64)i * 256 + tid;
            p[i] = j < N ? a_pos[j] : 0;
            g[i] = j < N ? a_grp[j] : 0;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u64 j = base + (u64)i * 256 + tid, q = (u64)p[i] + h;
            r[i] = (j < N && q < n) ? ISA[q] : 0;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u64 j = base + (u64)i * 256 + tid, q = (u64)p[i] + h;
            const bool valid = j < N;
            u64 key = 0;
            if (valid) {
                u64 k2 = 0;
                if (q < n) {
                    u32 rr = r[i];
                    if (LAZY && rr == kIsaInvalid) rr = lazy_rank(la, q);
                    k2 = (u64)rr + 1;
                }
                key = ((u64)g[i] << rank_bits) | k2;
                keys[j] = key;
                pos[j] = p[i];
            }
            for (int d = 0; d < P; ++d)
                warp_hist_add(sh + d * kRadixSize, digit_of(key, plan.shift[d], (1u << plan.nbits[d]) - 1), valid, lane);
        }
    }
    __syncthreads();
    for (int i = tid; i < P * kRadixSize; i += 256) {
        const u32 cnt = sh[i];
        if (cnt) atomicAdd((unsigned long long *)&hist[i], (unsigned long long)cnt);
    }
}

// ---------------------------------------------------------------------------------------------
// Round-0 digit histograms without touching the keys.  When a digit is a whole number of symbols
// (b divides 8) digit j of suffix p is the s-gram (s = 8/b symbols) at text position
// p + s*(P-1-j): every pass's histogram is the s-gram histogram H of the whole text, minus the
// s-grams before that offset, plus the zero padding past the end.  For bytes (b = 8) H is the
// byte histogram that was needed anyway (freq); for smaller alphabets one pass over the packed text.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 sgram_at(const u64 *__restrict__ words, u64 q, int b)
{
    u64 bit = q * (u64)b;
    u64 w = bit >> 6; int off = (int)(bit & 63);
    u64 hi = words[w], lo = words[w + 1];
    u64 x = off ? ((hi << off) | (lo >> (64 - off))) : hi;
    return (u32)(x >> 56);
}

__global__ void __launch_bounds__(256)
sgram_hist_kernel(const u64 *__restrict__ words, u64 n, int b, u64 *__restrict__ H)
{
    __shared__ u32 sh[8][256];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int per = 64 / b;                                   // positions whose s-gram starts in one word
    const u64 nw = ceil_div_dev(n * (u64)b, 64);
    for (u64 w = (u64)blockIdx.x * 256 + tid; w < nw; w += (u64)gridDim.x * 256) {
        u64 hi = words[w], lo = words[w + 1];
        u64 q0 = w * (u64)per;
        for (int i = 0; i < per; ++i) {
            if (q0 + i >= n) break;
            int off = i * b;
            u64 x = off ? ((hi << off) | (lo >> (64 - off))) : hi;
            atomicAdd(&sh[warp][(u32)(x >> 56)], 1u);
        }
    }
    __syncthreads();
    u32 t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][tid];
    if (t) atomicAdd((unsigned long long *)&H[tid], (unsigned long long)t);
}

// b = 8: H[code] = freq[symbol]
__global__ void sgram_from_freq_kernel(const u64 *__restrict__ freq, const u8 *__restrict__ lut, u64 *__restrict__ H)
{
    int s = threadIdx.x;
    u64 f = freq[s];
    if (f) H[lut[s]] = f;          // codes of present symbols are distinct
}

// grid = passes: hist[j] = H - (s-grams at positions < off_j) + off_j zero-padded windows past the end
__global__ void __launch_bounds__(256)
digit_hist_from_sgram_kernel(const u64 *__restrict__ H, const u64 *__restrict__ words, u64 n, int b, int passes,
                             u64 *__restrict__ hist)
{
    __shared__ unsigned long long sh[256];
    const int j = blockIdx.x, t = threadIdx.x;
    const int s = 8 / b;
    const u64 off = (u64)s * (u64)(passes - 1 - j);
    sh[t] = H[t];
    __syncthreads();
    // windows at positions [0, off) leave; `off` windows that start past the end (all padding, value 0) enter
    for (u64 q = t; q < off; q += 256) {
        if (q < n) atomicAdd(&sh[sgram_at(words, q, b)], (unsigned long long)-1ll);
        atomicAdd(&sh[0], 1ull);
    }
    __syncthreads();
    hist[j * 256 + t] = sh[t];
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
size_t sa_workspace_bytes(u64 n, int sym_bytes)
{
    u64 nw = ceil_div(n * 64, 64) + 4;                       // worst-case bitstream words (b <= 64)
    if (sym_bytes == 1) nw = ceil_div(n * 8, 64) + 4;
    else if (sym_bytes == 4) nw = ceil_div(n * 32, 64) + 4;
    size_t per = (size_t)n * (8 + 8 + 4 + 4      /* keys, vals (ping-pong) */
                              + 4 + 4            /* SA, ISA */
                              + 4 + 4 + 4 + 4    /* a_pos, a_grp, a_slot x2 */
                              + 1);              /* lazy mode: separate small round buffers */
    size_t st = 3 * ceil_div(n, 32) * sizeof(u32) + ceil_div(n, kRankTile) * (kRankWarps + 1) * 3 * sizeof(u32) + 1024;
    size_t msd = (size_t)(65536 + 65540 + 257 + 1024 * 256) * 4 + 256 * 8 + (ceil_div(n, 3072) + 2048) * kRadixSize * 8
               + (ceil_div(n, 1536) + 2) * 32 + (ceil_div(n, 3072) + 2048) * 8 + 12 * 256;    // round-0 MSD path: prefix histogram, offsets, chunk prefixes, tile status, tile table
    return per + nw * 8 + RadixSort<u64, u32>::temp_bytes(n) + st + msd + 256 + 32 * 256 + 4096;
}

static int choose_key_symbols(u64 n, int b, double entropy_bits, int max_key_bits)
{
    int kmax = max_key_bits / b;
    if (kmax < 1) kmax = 1;
    const char *env = getenv("LIBSAIS_CUDA_KEY_SYMBOLS");
    if (env && *env) { int k = atoi(env); if (k >= 1) return k < kmax ? k : kmax; }
    // enough symbols that an iid source of this entropy leaves ~2^-10 of the suffixes tied
    double need = std::log2((double)(n < 2 ? 2 : n)) + 10.0;
    if (entropy_bits < 0.05) entropy_bits = 0.05;
    double kk = std::ceil(need / entropy_bits);
    int k = kk > (double)kmax ? kmax : (int)kk;
    if (k < 1) k = 1;
    // round the key up to whole digit passes: extra symbols are free inside a pass.  (Cutting a key that overshoots a pass
    // boundary by a bit or two back instead saves config 3 one of six LSD passes, 16 ms -- and costs it 30 ms: the rounds
    // start at h = 20 instead of 24 and every one of them keeps more suffixes active.)
    int passes = (k * b + kRadixBits - 1) / kRadixBits;
    int k2 = (passes * kRadixBits) / b;
    if (k2 > kmax) k2 = kmax;
    return k2 > k ? k2 : k;
}

static bool read_round_scalars(Ctx &c)
{
    c.check(cudaMemcpyAsync(c.h_scalars + S_ERR, c.d_scalars + S_ERR, (S_BIGGRP - S_ERR + 1) * sizeof(u64),
                            cudaMemcpyDeviceToHost, c.stream));
    if (!c.sync()) return false;
    if (c.h_scalars[S_ERR] != 0) { c.last_error = cudaErrorLaunchTimeout; return false; }
    return true;
}

int build_sa(Ctx &c, const void *d_T, int sym_bytes, u64 n, const SAOptions &opt, SAResult *out)
{
    if (n == 0) return 0;
    if (n > kMaxN) return -2;
    cudaStream_t st = c.stream;
    const bool bwt_mode = opt.bwt_rows != nullptr && sym_bytes == 1;

    // ---- alphabet: bits per symbol, order-preserving code map (bytes), key width
    int b = 8; double entropy = 8.0;
    u8 *d_lut = nullptr;
    u8 lut[256];
    bool lut_identity = false;
    if (sym_bytes == 1) {
        run_byte_histogram(c, (const u8 *)d_T, n);
        c.check(cudaMemcpyAsync(c.h_scalars + S_FREQ, c.d_scalars + S_FREQ, 256 * sizeof(u64),
                                cudaMemcpyDeviceToHost, st));
        if (!c.sync()) return -2;
        int sigma = 0; entropy = 0;
        for (int s = 0; s < 256; ++s) {
            u64 f = c.h_scalars[S_FREQ + s];
            lut[s] = (u8)(sigma ? sigma : 0);
            if (f) { lut[s] = (u8)sigma; ++sigma; double pr = (double)f / (double)n; entropy -= pr * std::log2(pr); }
        }
        b = bits_for((u64)(sigma > 1 ? sigma - 1 : 1));
        lut_identity = sigma == 256;
        d_lut = c.alloc_n<u8>(256);
        if (!d_lut) return -2;
        // h_scalars tail as pinned staging for the LUT
        u8 *h_lut = (u8 *)(c.h_scalars + S_MISC);
        for (int s = 0; s < 256; ++s) h_lut[s] = lut[s];
        c.check(cudaMemcpyAsync(d_lut, h_lut, 256, cudaMemcpyHostToDevice, st));
    } else {
        c.check(cudaMemsetAsync(c.d_scalars + S_MAXSYM, 0, sizeof(u64), st));
        u32 grid = (u32)(ceil_div(n, 256 * 16) < (u64)c.sm_count * 8 ? ceil_div(n, 256 * 16) : (u64)c.sm_count * 8);
        if (sym_bytes == 4) LSC_LAUNCH(c, KC_HIST_SYM, (double)n * 4, max_sym_kernel<u32>, grid, 256, 0, (const u32 *)d_T, n, c.d_scalars + S_MAXSYM);
        else                LSC_LAUNCH(c, KC_HIST_SYM, (double)n * 8, max_sym_kernel<u64>, grid, 256, 0, (const u64 *)d_T, n, c.d_scalars + S_MAXSYM);
        c.check(cudaMemcpyAsync(c.h_scalars + S_MAXSYM, c.d_scalars + S_MAXSYM, sizeof(u64), cudaMemcpyDeviceToHost, st));
        if (!c.sync()) return -2;
        b = bits_for(c.h_scalars[S_MAXSYM]);
        entropy = (double)b;            // unknown distribution: assume dense
    }
    const int key_shift = bwt_mode ? 8 : 0;
    const int k = choose_key_symbols(n, b, entropy, 64 - key_shift);
    const int K = k * b;

    // ---- arena
    const u64 nwords = ceil_div(n * (u64)b, 64) + 2;
    u64 *words = c.alloc_n<u64>(nwords);
    u64 *keyA = c.alloc_n<u64>(n), *keyB = c.alloc_n<u64>(n);
    u32 *valA = c.alloc_n<u32>(n), *valB = c.alloc_n<u32>(n);
    u32 *SA = opt.want_sa ? (opt.sa_out ? opt.sa_out : c.alloc_n<u32>(n)) : nullptr;
    u32 *ISA = c.alloc_n<u32>(n);
    u32 *a_pos = c.alloc_n<u32>(n), *a_grp = c.alloc_n<u32>(n);
    u32 *a_slot0 = c.alloc_n<u32>(n), *a_slot1 = c.alloc_n<u32>(n);
    const u64 rank_tiles = ceil_div(n, kRankTile);
    u32 *rmasks = c.alloc_n<u32>(3 * ceil_div(n, 32));
    u32 *rwagg = c.alloc_n<u32>(rank_tiles * kRankWarps * 3);
    u32 *rtagg = c.alloc_n<u32>(rank_tiles * 3);
    void *sort_temp = c.alloc(RadixSort<u64, u32>::temp_bytes(n));
    if (!sort_temp || !rmasks || !rwagg || !rtagg || !a_slot1 || (opt.want_sa && !SA)) return -2;
    u32 *err = (u32 *)(c.d_scalars + S_ERR);
    c.check(cudaMemsetAsync(c.d_scalars + S_ERR, 0, (S_MISC - S_ERR) * sizeof(u64), st));
    c.check(cudaMemsetAsync(ISA, 0xFF, n * sizeof(u32), st));                  // kIsaInvalid everywhere

    // ---- pack + initial keys
    {
        u32 grid = (u32)ceil_div(nwords, 256);
        double ab = (double)n * sym_bytes + (double)nwords * 8;
        if (sym_bytes == 1 && b == 8 && lut_identity) LSC_LAUNCH(c, KC_PACK, ab, (pack_bytes_kernel<8, true>), grid, 256, 0, (const u8 *)d_T, n, words, nwords, d_lut);
        else if (sym_bytes == 1 && b == 8) LSC_LAUNCH(c, KC_PACK, ab, (pack_bytes_kernel<8, false>), grid, 256, 0, (const u8 *)d_T, n, words, nwords, d_lut);
        else if (sym_bytes == 1 && b == 4) LSC_LAUNCH(c, KC_PACK, ab, (pack_bytes_kernel<4, false>), grid, 256, 0, (const u8 *)d_T, n, words, nwords, d_lut);
        else if (sym_bytes == 1 && b == 2) LSC_LAUNCH(c, KC_PACK, ab, (pack_bytes_kernel<2, false>), grid, 256, 0, (const u8 *)d_T, n, words, nwords, d_lut);
        else if (sym_bytes == 1 && b == 1) LSC_LAUNCH(c, KC_PACK, ab, (pack_bytes_kernel<1, false>), grid, 256, 0, (const u8 *)d_T, n, words, nwords, d_lut);
        else if (sym_bytes == 1) LSC_LAUNCH(c, KC_PACK, ab, (pack_kernel<u8, true>), grid, 256, 0, (const u8 *)d_T, n, b, words, nwords, d_lut);
        else if (sym_bytes == 4) LSC_LAUNCH(c, KC_PACK, ab, (pack_kernel<u32, false>), grid, 256, 0, (const u32 *)d_T, n, b, words, nwords, (const u8 *)nullptr);
        else LSC_LAUNCH(c, KC_PACK, ab, (pack_kernel<u64, false>), grid, 256, 0, (const u64 *)d_T, n, b, words, nwords, (const u8 *)nullptr);
    }

    // ---- round 0: sort by the k-mer, rank, compact
    RoundStat rs; rs.h = 0; rs.n_active = n; rs.key_bits = K; rs.passes = 0; rs.n_groups = 0;
    bool fuse_keys = true;
    { const char *env = getenv("LIBSAIS_CUDA_FUSE_KEYS"); if (env && *env) fuse_keys = atoi(env) != 0; }
    int where;
    // ---- MSD path (partition.cuh): when the 16-bit key prefixes spread the suffixes over small buckets (random
    // bytes, iid DNA: any near-uniform source), two UNSTABLE partition passes on the top 16 key bits and an
    // in-shared-memory finish of every bucket replace the K/8 stable LSD passes.  Same sorted arrays.
    bool msd = false, msd_fused = false, stream_rows = false;
    u32 *m_boff = nullptr, *m_tstart = nullptr, *m_H = nullptr; u64 *m_base = nullptr; u64 m_maxb = 0, m_maxnt = 0;
    // LIBSAIS_CUDA_PART_PIPE: bit 0 = persistent double-buffered kernel for the first MSD level, bit 1 = for the second.
    // Measured on B200 (profiles/part_pass_r2.md): it pays for the array source (TMA prefetch: 2.15 -> 1.99 ms) and not for the
    // k-mer source (cp.async staging costs more issue slots than it hides: 1.73 -> 2.05 ms), hence the default 2.
    static const int pipe_env = [] { const char *e = getenv("LIBSAIS_CUDA_PART_PIPE"); return (e && *e) ? atoi(e) : 2; }();
    const int pipe = part_tile() == kPipeTile ? pipe_env : 0;       // both kernels of a run must agree on the tile (chunk / cell grid)
    const u32 ptile = part_tile();
    const u64 nt1 = ceil_div(n, (u64)ptile);
    u64 want_seg = 16;                                                              // chunks of the first pass (profiles/part_pass_r2.md)
    { const char *env = getenv("LIBSAIS_CUDA_PART_NSEG"); if (env && *env && atoi(env) > 0) want_seg = (u64)atoi(env); }
    if (want_seg > 1024) want_seg = 1024;
    u32 nseg1 = (u32)(nt1 < want_seg ? nt1 : want_seg);
    const u32 tpc = (u32)ceil_div(nt1, (u64)nseg1);
    nseg1 = (u32)ceil_div(nt1, (u64)tpc);
    {
        int mode = 1;
        { const char *env = getenv("LIBSAIS_CUDA_MSD"); if (env && *env) mode = atoi(env); }
        const u64 min_n = mode >= 2 ? 64 : ((u64)1 << 16);
        if (fuse_keys && mode > 0 && (8 % b) == 0 && K >= 16 && n >= min_n) {
            u32 *h16 = c.alloc_n<u32>(65536);
            m_boff = c.alloc_n<u32>(65537 + 3); m_tstart = c.alloc_n<u32>(kRadixSize + 1); m_base = c.alloc_n<u64>(kRadixSize);
            m_H = c.alloc_n<u32>((size_t)nseg1 * kRadixSize);
            if (!h16 || !m_boff || !m_tstart || !m_base || !m_H) return -2;
            c.check(cudaMemsetAsync(h16, 0, 65536 * sizeof(u32), st));
            c.check(cudaMemsetAsync(c.d_scalars + S_MSD, 0, 4 * sizeof(u64), st));
            // histogram units: every chunk is cut into upc units of ut tiles so that ~2 units per SM exist
            const u32 q = (u32)ceil_div((u64)c.sm_count * 2, (u64)nseg1);
            const u32 ut = (u32)ceil_div((u64)tpc, (u64)q), upc = (u32)ceil_div((u64)tpc, (u64)ut);
            const u32 nunits = nseg1 * upc;
            const u32 grid = nunits < (u32)c.sm_count ? nunits : (u32)c.sm_count;
            c.check(cudaMemsetAsync(m_H, 0, (size_t)nseg1 * kRadixSize * sizeof(u32), st));
            const size_t hsm = kHist16Words * sizeof(u32);
            const double hb = (double)ceil_div(n * (u64)b, 64) * 8;
            u64 *flag = c.d_scalars + S_MSD + 2;
            if (b == 8)      { c.check(cudaFuncSetAttribute(hist16_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
                               LSC_LAUNCH(c, KC_SORT_HIST, hb, hist16_kernel<8>, grid, kHist16Threads, hsm, words, n, h16, m_H, nunits, upc, ut, tpc, ptile, flag); }
            else if (b == 4) { c.check(cudaFuncSetAttribute(hist16_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
                               LSC_LAUNCH(c, KC_SORT_HIST, hb, hist16_kernel<4>, grid, kHist16Threads, hsm, words, n, h16, m_H, nunits, upc, ut, tpc, ptile, flag); }
            else if (b == 2) { c.check(cudaFuncSetAttribute(hist16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
                               LSC_LAUNCH(c, KC_SORT_HIST, hb, hist16_kernel<2>, grid, kHist16Threads, hsm, words, n, h16, m_H, nunits, upc, ut, tpc, ptile, flag); }
            else             { c.check(cudaFuncSetAttribute(hist16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
                               LSC_LAUNCH(c, KC_SORT_HIST, hb, hist16_kernel<1>, grid, kHist16Threads, hsm, words, n, h16, m_H, nunits, upc, ut, tpc, ptile, flag); }
            LSC_LAUNCH(c, KC_SORT_SCAN, 65536.0 * 8, scan16_kernel, 1, 1024, 0, h16, m_boff, m_base, m_tstart, m_H, nseg1, c.d_scalars + S_MSD, ptile);
            c.check(cudaMemcpyAsync(c.h_scalars + S_MSD, c.d_scalars + S_MSD, 3 * sizeof(u64), cudaMemcpyDeviceToHost, st));
            if (opt.h_U && opt.h_T && bwt_mode) {
                // bucket of suffix 0 (the row the BWT drops) from the first 16 / b symbols of the text: its slot range
                // [p0_lo, p0_hi) tells, before the sort, on which side of the dropped row every other slot lies
                u32 pre = 0;
                for (int i = 0; i < 16 / b; ++i) pre = (pre << b) | ((u64)i < n ? (u32)lut[opt.h_T[i]] : 0u);
                c.check(cudaMemcpyAsync(c.h_scalars + S_MSD + 3, m_boff + pre, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
            }
            if (!c.sync()) return -2;
            m_maxb = c.h_scalars[S_MSD]; m_maxnt = c.h_scalars[S_MSD + 1];
            msd = c.h_scalars[S_MSD + 2] == 0 && m_maxb <= (u64)kBucketMaxBucket;
        }
    }
    if (msd) {
        const u64 grid1 = (u64)nseg1 * tpc, grid2 = (u64)kRadixSize * m_maxnt;
        const u64 gmax = grid1 > grid2 ? grid1 : grid2;
        const size_t stw = n < (1ull << 30) ? sizeof(u32) : sizeof(u64);
        void *status = c.alloc(gmax * kRadixSize * stw);
        u32 Cw = (u32)kBucketCap - (u32)m_maxb;
        if (Cw > 6144) Cw = 6144;
        const u64 btiles = ceil_div(n, (u64)Cw);
        uint4 *tb = c.alloc_n<uint4>(2 * (btiles + 1));
        uint2 *tinfo = c.alloc_n<uint2>(grid2 + 1);
        if (!status || !tb || !tinfo) return -2;
        u32 *tickets = (u32 *)(c.d_scalars + S_TICKET);
        c.check(cudaMemsetAsync(tickets, 0, 8 * sizeof(u64), st));
        static const bool atomic_tickets = [] { const char *e = getenv("LIBSAIS_CUDA_TICKETS"); return e && *e && atoi(e) != 0; }();
        KmerSrc src; src.words = words; src.nwords = nwords; src.text = bwt_mode ? (const u8 *)d_T : nullptr; src.n = n; src.b = b; src.K = K; src.key_shift = key_shift;
        PartArgs pa; pa.n = n; pa.dmask = 255u; pa.err = err; pa.use_bulk = 0; pa.kptr = nullptr; pa.vptr = nullptr;
        pa.shift = key_shift + K - 8; pa.base = m_base; pa.cp = m_H; pa.nseg = nseg1; pa.tpc = tpc; pa.boff = nullptr; pa.tstart = nullptr; pa.tinfo = nullptr;
        pa.ticket = atomic_tickets ? tickets : nullptr;
        c.check(cudaMemsetAsync(status, 0, grid1 * kRadixSize * stw, st));
        if (pipe & 1) launch_part_pipe<KmerSrc, false>(c, KC_SORT_PASS_GEN, (double)n * (2.0 + 12.0), src, (const u64 *)nullptr, (const u32 *)nullptr, keyA, valA, pa, grid1, status);
        else launch_part_pass<u64, u32, KmerSrc, false>(c, KC_SORT_PASS_GEN, (double)n * (2.0 + 12.0), src, (const u64 *)nullptr, (const u32 *)nullptr, keyA, valA, pa, grid1, status);
        pa.shift = key_shift + K - 16; pa.base = nullptr; pa.cp = nullptr; pa.nseg = kRadixSize; pa.tpc = 0; pa.boff = m_boff; pa.tstart = m_tstart; pa.tinfo = tinfo;
        pa.ticket = atomic_tickets ? tickets + 1 : nullptr;
        LSC_LAUNCH(c, KC_SORT_HIST, 0.0, seg_tiles_kernel, (u32)ceil_div(grid2, 256), 256, 0, m_boff, m_tstart, ptile, (u32)grid2, tinfo);
        c.check(cudaMemsetAsync(status, 0, grid2 * kRadixSize * stw, st));
        // second level: only the key bits below the bucket prefix are written when they fit 32 bits (bucket_load_key)
        int k32 = (pipe & 2) && K - 16 + key_shift <= 32;
        { const char *env = getenv("LIBSAIS_CUDA_MSD_K32"); if (env && *env && atoi(env) == 0) k32 = 0; }
        pa.out32 = k32;
        if (pipe & 2) launch_part_pipe<ArraySrc, true>(c, KC_PART_PASS, (double)n * (k32 ? 20.0 : 24.0), ArraySrc(), keyA, valA, keyB, valB, pa, grid2, status);
        else launch_part_pass<u64, u32, ArraySrc, true>(c, KC_PART_PASS, (double)n * 24.0, ArraySrc(), keyA, valA, keyB, valB, pa, grid2, status);
        LSC_LAUNCH(c, KC_SORT_HIST, 0.0, bucket_tiles_kernel, (u32)ceil_div(btiles, 256), 256, 0, m_boff, btiles, Cw, tb);
        // fused rank stage: the sorted keys are never written; positions go straight to the suffix array (or valA)
        { const char *env = getenv("LIBSAIS_CUDA_MSD_FUSE"); msd_fused = !(env && *env && atoi(env) == 0); }
        BucketFuse fz; fz.on = msd_fused ? 1 : 0; fz.flags = (u8 *)a_slot1; fz.rows = bwt_mode ? opt.bwt_rows : nullptr;      // a_slot1 is idle until the second round
        fz.tail_start = n >= (u64)k ? n - (u64)k + 1 : 0;
        fz.aux_I = opt.aux_I; fz.aux_mask = opt.aux_I ? opt.aux_r - 1 : 0; fz.aux_shift = opt.aux_I ? bits_for(opt.aux_r) - 1 : 0;
        fz.primary = c.d_scalars + S_PRIMARY; fz.big_flag = c.d_scalars + S_BIGGRP;
        {
            const double ab = (double)n * ((k32 ? 8.0 : 12.0) + (msd_fused ? 4.0 + (bwt_mode ? 1.0 : 0.0) + 1.0 : 12.0));
            u32 *vout = (msd_fused && SA) ? SA : valA;
            // Streamed rows: the tiles are launched in chunks; when a chunk is done every slot below its last window is final
            // for round 0 and its row bytes start their way to the caller's pinned buffer on the copy stream, shifted by one
            // below the dropped row.  The bucket of suffix 0 itself waits for the primary index (api.cu bwt_body).
            const u32 *h_b0 = (const u32 *)(c.h_scalars + S_MSD + 3);
            const u64 p0_lo = h_b0[0], p0_hi = h_b0[1];
            stream_rows = msd_fused && bwt_mode && opt.h_U && opt.h_T && btiles >= 64 && p0_lo < p0_hi && p0_hi <= n && c.ensure_copy_stream();
            { const char *env = getenv("LIBSAIS_CUDA_STREAM_ROWS"); if (env && *env && atoi(env) == 0) stream_rows = false; }
            const int nchunk = stream_rows ? 8 : 1;
            out->p0_lo = p0_lo; out->p0_hi = p0_hi;
#define LSC_BUCKET_SORT(IN32, FUSED)                                                                                                        \
            do {                                                                                                                            \
                c.check(cudaFuncSetAttribute(bucket_sort_kernel<IN32, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BucketSmem))); \
                LSC_LAUNCH(c, KC_BUCKET_SORT, ab * (double)(t1 - t0) / (double)btiles, (bucket_sort_kernel<IN32, FUSED>), (u32)(t1 - t0), kBucketThreads, sizeof(BucketSmem), \
                           keyB, valB, tb, n, Cw, key_shift, K - 16, keyA, vout, err, fz, m_boff, (u32)t0);                                 \
            } while (0)
            for (int ck = 0; ck < nchunk; ++ck) {
                const u64 t0 = btiles * (u64)ck / nchunk, t1 = btiles * (u64)(ck + 1) / nchunk;
                if (t1 == t0) continue;
                if (k32) { if (msd_fused) LSC_BUCKET_SORT(true, true); else LSC_BUCKET_SORT(true, false); }
                else     { if (msd_fused) LSC_BUCKET_SORT(false, true); else LSC_BUCKET_SORT(false, false); }
                if (stream_rows) {
                    c.check(cudaEventRecord(c.chunk_ev[ck], st));
                    c.check(cudaStreamWaitEvent(c.copy_stream, c.chunk_ev[ck], 0));
                    const u64 lo = t0 * (u64)Cw, hi = ck == nchunk - 1 ? n : t1 * (u64)Cw;      // slots that are final now
                    const u64 a1 = hi < p0_lo ? hi : p0_lo;                                      // [lo, a1): below the dropped row -> U[slot + 1]
                    if (lo < a1) c.check(cudaMemcpyAsync(opt.h_U + lo + 1, opt.bwt_rows + lo, a1 - lo, cudaMemcpyDeviceToHost, c.copy_stream));
                    const u64 b0s = lo > p0_hi ? lo : p0_hi;                                     // [b0s, hi): above it -> U[slot]
                    if (b0s < hi) c.check(cudaMemcpyAsync(opt.h_U + b0s, opt.bwt_rows + b0s, hi - b0s, cudaMemcpyDeviceToHost, c.copy_stream));
                }
            }
#undef LSC_BUCKET_SORT
        }
        rs.passes = 2;
        where = c.failed() ? -1 : 0;
    } else if (fuse_keys) {
        KmerGen gen; gen.words = words; gen.text = bwt_mode ? (const u8 *)d_T : nullptr; gen.n = n; gen.b = b; gen.K = K; gen.key_shift = key_shift;
        // digit histograms straight from the s-gram histogram of the text when digits are symbol aligned
        bool hist_ready = false;
        const int passes0 = (K + kRadixBits - 1) / kRadixBits;
        bool sgram = (8 % b) == 0 && (K % 8) == 0 && n > 4096;
        { const char *env = getenv("LIBSAIS_CUDA_SGRAM_HIST"); if (env && *env) sgram = sgram && atoi(env) != 0; }
        if (sgram) {
            u64 *H = c.d_scalars + S_SGRAM;
            u64 *hist = (u64 *)sort_temp;
            c.check(cudaMemsetAsync(H, 0, 256 * sizeof(u64), st));
            if (b == 8 && sym_bytes == 1) {
                LSC_LAUNCH(c, KC_SORT_HIST, 0.0, sgram_from_freq_kernel, 1, 256, 0, c.d_scalars + S_FREQ, d_lut, H);
            } else {
                u64 nw = ceil_div(n * (u64)b, 64);
                u32 grid = (u32)(ceil_div(nw, 256 * 8) < (u64)c.sm_count * 8 ? ceil_div(nw, 256 * 8) : (u64)c.sm_count * 8);
                LSC_LAUNCH(c, KC_SORT_HIST, (double)nw * 8, sgram_hist_kernel, grid ? grid : 1, 256, 0, words, n, b, H);
            }
            LSC_LAUNCH(c, KC_SORT_HIST, 0.0, digit_hist_from_sgram_kernel, passes0, 256, 0, H, words, n, b, passes0, hist);
            hist_ready = true;
        }
        where = RadixSort<u64, u32>::sort_from<KmerGen>(c, gen, keyA, valA, keyB, valB, n, key_shift, key_shift + K, sort_temp, err, &rs.passes, hist_ready);
    } else {
        LSC_LAUNCH(c, KC_MAKE_KEYS, (double)nwords * 8 + (double)n * (12 + (bwt_mode ? 1 : 0)), make_keys_kernel,
                   (u32)ceil_div(n, 256), 256, 0, words, n, b, K, key_shift, bwt_mode ? (const u8 *)d_T : (const u8 *)nullptr, keyA, valA);
        where = RadixSort<u64, u32>::sort(c, keyA, valA, keyB, valB, n, key_shift, key_shift + K, sort_temp, err, &rs.passes);
    }
    if (where < 0) return -2;
    u64 *ks = where ? keyB : keyA; u32 *vs = where ? valB : valA;      // S0: sorted round-0 (key, pos)
    u32 *const vbuf = vs;                                               // the arena buffer (ping-pong half of the later rounds)
    if (msd_fused && SA) vs = SA;                                       // fused MSD path: the positions in slot order went straight to SA
    u64 *ko = where ? keyA : keyB; u32 *vo = where ? valA : valB;
    u32 *slot_cur = a_slot0, *slot_nxt = a_slot1;
    const u64 tail_start = n >= (u64)k ? n - (u64)k + 1 : 0;

    RankArgs ra;
    ra.keys = ks; ra.pos = vs; ra.slot_in = nullptr; ra.N = n; ra.tail_start = tail_start; ra.key_shift = key_shift;
    ra.slot_base = 0;
    ra.SA = SA; ra.ISA = ISA; ra.isa_all = 0;
    ra.rows = bwt_mode ? opt.bwt_rows : nullptr; ra.text = bwt_mode ? (const u8 *)d_T : nullptr;
    ra.aux_I = opt.aux_I; ra.aux_mask = opt.aux_I ? opt.aux_r - 1 : 0; ra.aux_shift = opt.aux_I ? bits_for(opt.aux_r) - 1 : 0;
    ra.primary = c.d_scalars + S_PRIMARY;
    ra.a_pos = a_pos; ra.a_slot = slot_cur; ra.a_grp = a_grp;
    ra.masks = rmasks; ra.wagg = rwagg; ra.tagg = rtagg; ra.nchunks = ceil_div(n, 32);
    ra.ntiles = rank_tiles; ra.out_counts = c.d_scalars + S_NACT;
    ra.pair_idx = nullptr; ra.pair_val = nullptr; ra.phist = nullptr; ra.pshift = 0;
    if (msd_fused) LSC_LAUNCH(c, KC_RANK_INIT, (double)n * 1.25, rank_agg_kernel, (u32)ceil_div(rank_tiles, (u64)kAggTiles), kRankThreads, 0, (const u8 *)a_slot1, rmasks, n, rwagg, rtagg, rank_tiles);
    else LSC_LAUNCH(c, KC_RANK_INIT, (double)n * (12 + (SA ? 4 : 0) + (bwt_mode ? 1 : 0)), rank_flags_kernel<true>, (u32)rank_tiles, kRankThreads, 0, ra);
    LSC_LAUNCH(c, KC_RANK_SCAN, (double)rank_tiles * 24, rank_scan_kernel, kRankScanCtas, 1024, 0, rtagg, rank_tiles, c.d_scalars + S_NACT);
    if (!read_round_scalars(c)) return -2;
    u64 N = c.h_scalars[S_NACT], G = c.h_scalars[S_NGRP];
    rs.n_groups = G;
    // fused MSD path: the bucket sort saw every group's size -- when none exceeds 128 the lazy rounds order the groups where
    // they lie (local_count_kernel<LAZY>: one kernel instead of key build + histogram + six digit passes over a few thousand pairs)
    bool lazy_local = msd_fused && !(c.h_scalars[S_BIGGRP] & 1);
    { const char *env = getenv("LIBSAIS_CUDA_LAZY_LOCAL"); if (env && *env && atoi(env) == 0) lazy_local = false; }

    // ---- lazy ISA: with few active suffixes the ranks of round-0 singletons are never scattered;
    // the rare look-ups that hit one recompute it by binary search in the sorted round-0 keys.
    bool lazy = N * 32 <= n;
    { const char *env = getenv("LIBSAIS_CUDA_LAZY_ISA"); if (env && *env) lazy = atoi(env) != 0; }
    u64 *rk0 = ks, *rk1 = ko; u32 *rv0 = vbuf, *rv1 = vo;
    if (N == 0 && stream_rows) out->u_streamed = true;          // everything settled in round 0: nothing to send again
    if (N > 0) {
        if (lazy) {
            rk0 = c.alloc_n<u64>(N); rk1 = c.alloc_n<u64>(N); rv0 = c.alloc_n<u32>(N); rv1 = c.alloc_n<u32>(N);
            if (!rv1 || !rk1 || !rk0 || !rv0) {                 // no room for separate buffers: fall back to the full ISA
                lazy = false; rk0 = ks; rk1 = ko; rv0 = vbuf; rv1 = vo; c.last_error = cudaSuccess;
            }
        }
        if (lazy) {
            ra.isa_all = 0;
            LSC_LAUNCH(c, KC_RANK_INIT, (double)N * 16, rank_apply_kernel<true>, (u32)ceil_div(rank_tiles, kApplyTiles), kRankThreads, 0, ra);
            if (stream_rows && N <= ((u64)1 << 20)) {
                // streamed rows: the slots that are still open keep the list of rows to send again when they are final
                u32 *saved = c.alloc_n<u32>(N);
                if (saved) {
                    c.check(cudaMemcpyAsync(saved, slot_cur, N * sizeof(u32), cudaMemcpyDeviceToDevice, st));
                    out->patch_slots = saved; out->n_patch = N; out->u_streamed = true;
                } else c.last_error = cudaSuccess;
            }
        } else {
            // every rank is needed: (position, rank) pairs in slot order, then a locality-partitioned scatter
            ra.isa_all = 1; ra.pair_idx = (u32 *)ko; ra.pair_val = (u32 *)ko + n;
            const bool fused_hist = scatter_hist_prepare(c, n, n, sort_temp, &ra.phist, &ra.pshift);
            LSC_LAUNCH(c, KC_RANK_INIT, (double)n * (4 + 8) + (double)N * 12, rank_apply_kernel<true>, (u32)ceil_div(rank_tiles, kApplyTiles), kRankThreads, 0, ra);
            if (partitioned_scatter<NoGen>(c, NoGen(), ra.pair_idx, ra.pair_val, (u32 *)ks, (u32 *)ks + n, n, n, ISA, sort_temp, err, fused_hist) != 0) return -2;
            ra.phist = nullptr;
        }
    }
    c.rounds.push_back(rs);             // (here: the rank scatter above belongs to round 0 in the per-round accounting)
    LazyArgs la; la.words = words; la.b = b; la.K = K; la.s0_keys = msd_fused ? nullptr : ks; la.s0_pos = vs; la.key_shift = key_shift;
    la.tail_start = tail_start; la.n = n; la.boff16 = msd ? m_boff : nullptr;

    // ---- doubling rounds on the active suffixes
    const int rank_bits = bits_for(n);                 // k2 = ISA+1 <= n
    u64 h = (u64)k;
    int round = 1;
    // Rounds whose groups are all small are sorted where they lie (local_sort.cuh); a round with a large
    // group, few active suffixes, or the lazy ISA takes the global onesweep.
    bool local_on = true;
    { const char *env = getenv("LIBSAIS_CUDA_LOCAL_SORT"); if (env && *env) local_on = atoi(env) != 0; }
    local_on = local_on && !lazy;
    u64 kLocalMin = (u64)1 << 18;
    { const char *env = getenv("LIBSAIS_CUDA_LOCAL_MIN"); if (env && *env && atoi(env) > 0) kLocalMin = (u64)atoi(env); }      // tests: small texts through the local paths
    auto launch_big_check = [&](u64 n_upper) {
        c.check(cudaMemsetAsync(c.d_scalars + S_BIGGRP, 0, sizeof(u64), st));
        const u64 want = ceil_div(n_upper, 256 * 8);
        const u32 grid = (u32)(want < (u64)c.sm_count * 16 ? (want ? want : 1) : (u64)c.sm_count * 16);
        LSC_LAUNCH(c, KC_LOCAL_SORT, 0.0, big_group_kernel, grid, 256, 0, a_grp, c.d_scalars + S_NACT, c.d_scalars + S_BIGGRP);
    };
    // window of a local-sort tile from the group-size flags (0: some group is too large for a tile)
    auto local_window = [&]() -> u32 {          // 1: every group fits the counting kernel
        const u64 f = c.h_scalars[S_BIGGRP];
        if (!(f & 1)) return 1;
        for (int i = 0; i < 3; ++i) if (!((f >> (i + 1)) & 1)) return (u32)kLocalCap - kLocalLimits[i];
        return 0;
    };
    u32 win = 0;
    if (local_on && N >= kLocalMin) {
        c.check(cudaFuncSetAttribute(local_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LocalSmem)));
        launch_big_check(N);
        if (!read_round_scalars(c)) return -2;
        win = local_window();
    }
    // ---- position-ordered rounds (po_rounds.cuh): no group larger than kPoMaxGroup and enough suffixes left to pay for
    // the one-time reordering of the groups.  From here on the active record is (position, rank of the group).
    // The test is repeated before every slot-ordered round: a text whose round-0 groups are too large (natural language: the
    // suffixes of a frequent word) runs slot-ordered rounds until they have split far enough, then switches over.
    int po_env = 1;
    { const char *env = getenv("LIBSAIS_CUDA_PO"); if (env && *env) po_env = atoi(env); }
    bool flags_valid = local_on && N >= kLocalMin;              // S_BIGGRP describes the current active list
    while (N > 0) {
    if (po_env && local_on && flags_valid && N >= kLocalMin && !(c.h_scalars[S_BIGGRP] & 4)) {
        const u64 f = c.h_scalars[S_BIGGRP];
        const u32 limit = !(f & 1) ? kLocalCountLimit : !(f & 2) ? 512u : kPoMaxGroup;
        const u32 C = (u32)kPoCap - limit;
        // groups sorted by the position of their head suffix; members keep their slot order (po_rounds.cuh)
        u32 *ghead_pos = (u32 *)keyA, *gid = ghead_pos + G, *ghead_slot = gid + G;            // 3 G     <= 1.5 n words of keyA's 2 n
        u32 *tmpk = (u32 *)keyB, *tmpv = tmpk + G, *gstart = tmpv + G;                        // 3 G + 1 <= 1.5 n + 1 words of keyB's 2 n
        LSC_LAUNCH(c, KC_ROUND_KEYS, (double)N * 4 + (double)G * 24, po_group_table_kernel, (u32)ceil_div(N, 256), 256, 0,
                   a_pos, a_grp, slot_cur, N, gstart, ghead_pos, ghead_slot, gid);
        const int pos_bits = bits_for(n - 1);
        int bin_shift = 7;
        { const char *env = getenv("LIBSAIS_CUDA_PO_BIN"); if (env && *env) { bin_shift = atoi(env); if (bin_shift < 0) bin_shift = 0; if (bin_shift > 24) bin_shift = 24; } }
        u32 run_div = 4;
        { const char *env = getenv("LIBSAIS_CUDA_PO_RUNS"); if (env && *env && atoi(env) > 0) run_div = (u32)atoi(env); }
        RoundStat r0; r0.h = h; r0.n_active = N; r0.n_groups = G; r0.passes = 0; r0.key_bits = pos_bits;
        c.pass_class_override = KC_ROUND_KEYS;
        where = RadixSort<u32, u32>::sort(c, ghead_pos, gid, tmpk, tmpv, G, 0, pos_bits, sort_temp, err, &r0.passes);
        c.pass_class_override = -1;
        if (where < 0) return -2;
        const u32 *sg = where ? tmpv : gid;                  // group ids by ascending head position
        u32 *newstart = where ? ghead_pos : tmpk;            // the sort's other key buffer is free
        {
            const u64 nb = ceil_div(G, (u64)kPoScanChunk);
            u32 *bsum = (u32 *)sort_temp;
            LSC_LAUNCH(c, KC_ROUND_KEYS, (double)G * 12, po_scan_sums_kernel, (u32)nb, kPoScanThreads, 0, sg, gstart, G, bsum);
            LSC_LAUNCH(c, KC_ROUND_KEYS, (double)nb * 8, po_scan_top_kernel, 1, 1024, 0, bsum, nb);
            LSC_LAUNCH(c, KC_ROUND_KEYS, (double)G * 16, po_scan_apply_kernel, (u32)nb, kPoScanThreads, 0, sg, gstart, G, bsum, newstart);
        }
        LSC_LAUNCH(c, KC_ROUND_KEYS, (double)N * 16 + (double)G * 12, po_move_kernel, (u32)ceil_div(N, 256), 256, 0,
                   a_pos, a_grp, N, gstart, newstart, ghead_slot, valA, valB);
        u32 *lp[2] = {a_pos, valA}, *lr[2] = {a_grp, valB};    // lists: out = [cur], in = [cur ^ 1]
        u32 *pair_pos = a_slot0, *pair_rank = a_slot1;
        int cur = 0; bool first = true;
        while (N > 0) {
            if (round > 80) { c.last_error = cudaErrorUnknown; return -2; }
            RoundStat r; r.h = h; r.n_active = N; r.key_bits = rank_bits + 11; r.passes = first ? r0.passes : 0; r.n_groups = 0;
            const u64 tiles = ceil_div(N, (u64)C);
            c.check(cudaMemsetAsync(sort_temp, 0, tiles * sizeof(u64), st));
            c.check(cudaMemsetAsync(c.d_scalars + S_TICKET, 0, sizeof(u64), st));
            c.check(cudaMemsetAsync(c.d_scalars + S_NACT, 0, 2 * sizeof(u64), st));
            PoArgs pa; pa.a_pos = lp[cur ^ 1]; pa.a_rank = lr[cur ^ 1];
            pa.N = N; pa.n = n; pa.h = h; pa.C = C; pa.bin_shift = bin_shift; pa.run_div = run_div; pa.ISA = ISA; pa.o_pos = lp[cur]; pa.o_rank = lr[cur];
            pa.pair_pos = pair_pos; pa.pair_rank = pair_rank;                        // both 16-byte aligned (po_apply_kernel)
            pa.SA = SA; pa.rows = bwt_mode ? opt.bwt_rows : nullptr; pa.text = bwt_mode ? (const u8 *)d_T : nullptr;
            pa.aux_mask = ra.aux_mask; pa.aux_shift = ra.aux_shift; pa.aux_I = opt.aux_I; pa.primary = c.d_scalars + S_PRIMARY;
            pa.status = (u64 *)sort_temp; pa.ticket = (u32 *)(c.d_scalars + S_TICKET); pa.out_counts = c.d_scalars + S_NACT; pa.err = err;
            pa.ntiles = (u32)tiles;
            LSC_LAUNCH(c, KC_LOCAL_SORT, (double)N * (8 + 4 + 8 + 8), po_round_kernel, (u32)tiles, kPoThreads, 0, pa);
            LSC_LAUNCH(c, KC_SCATTER, (double)N * 12, po_apply_kernel, (u32)ceil_div(N, 2048), 256, 0, pa.pair_pos, pa.pair_rank, N, ISA);
            if (!read_round_scalars(c)) return -2;
            N = c.h_scalars[S_NACT]; G = c.h_scalars[S_NGRP];
            r.n_groups = G;
            c.rounds.push_back(r);
            first = false; cur ^= 1; h *= 2; ++round;
        }
        break;
    }
    {
        if (round > 80) { c.last_error = cudaErrorUnknown; return -2; }
        const int grp_bits = bits_for(G > 1 ? G - 1 : 1);
        RoundStat r; r.h = h; r.n_active = N; r.key_bits = rank_bits + grp_bits; r.passes = 0; r.n_groups = 0;
        const bool local = (local_on && win != 0 && N >= kLocalMin) || (lazy && lazy_local);
        if (local) {
            // key build + sort in one kernel, in place of round_keys + onesweep (r.passes stays 0)
            if (lazy) LSC_LAUNCH(c, KC_LOCAL_SORT, (double)N * (4 + 4 + 4 + 12), local_count_kernel<true>, (u32)ceil_div(N, kCountWindow), kCountThreads, 0,
                                 a_pos, a_grp, ISA, N, n, h, rank_bits, rk1, rv1, err, la);
            else if (win == 1) LSC_LAUNCH(c, KC_LOCAL_SORT, (double)N * (4 + 4 + 4 + 12), local_count_kernel<false>, (u32)ceil_div(N, kCountWindow), kCountThreads, 0,
                                     a_pos, a_grp, ISA, N, n, h, rank_bits, rk1, rv1, err, la);
            else LSC_LAUNCH(c, KC_LOCAL_SORT, (double)N * (4 + 4 + 4 + 12), local_sort_kernel, (u32)ceil_div(N, win), kLocalThreads, sizeof(LocalSmem),
                            a_pos, a_grp, ISA, N, n, h, rank_bits, win, rk1, rv1, err);
            where = 1;
        } else {
            const int key_bits = rank_bits + grp_bits;
            const SortPlan plan = make_sort_plan(0, key_bits);
            c.check(cudaMemsetAsync(sort_temp, 0, (size_t)kMaxPasses * kRadixSize * sizeof(u64), st));
            const u64 want = ceil_div(N, 1024);
            const u32 grid = (u32)(want < (u64)c.sm_count * 8 ? want : (u64)c.sm_count * 8);
            if (lazy) LSC_LAUNCH(c, KC_ROUND_KEYS, (double)N * (4 + 4 + 4 + 12), round_keys_kernel<true>, grid, 256, 0,
                                 a_pos, a_grp, ISA, N, n, h, rank_bits, rk0, rv0, la, plan, (u64 *)sort_temp);
            else      LSC_LAUNCH(c, KC_ROUND_KEYS, (double)N * (4 + 4 + 4 + 12), round_keys_kernel<false>, grid, 256, 0,
                                 a_pos, a_grp, ISA, N, n, h, rank_bits, rk0, rv0, la, plan, (u64 *)sort_temp);
            where = RadixSort<u64, u32>::sort_from<NoGen>(c, NoGen(), rk0, rv0, rk1, rv1, N, 0, key_bits, sort_temp, err, &r.passes, true);
        }
        if (where < 0) return -2;
        const u64 tiles = ceil_div(N, kRankTile);
        ra.keys = where ? rk1 : rk0; ra.pos = where ? rv1 : rv0; ra.slot_in = slot_cur; ra.N = N; ra.tail_start = 0; ra.key_shift = 0;
        ra.isa_all = 1; ra.a_slot = slot_nxt; ra.ntiles = tiles; ra.nchunks = ceil_div(N, 32);
        u64 *other_k = where ? rk0 : rk1;                     // the ping-pong half not holding the sorted result
        ra.pair_idx = (u32 *)other_k; ra.pair_val = (u32 *)other_k + N;
        const bool fused_hist = scatter_hist_prepare(c, N, n, sort_temp, &ra.phist, &ra.pshift);
        LSC_LAUNCH(c, KC_RANK_UPDATE, (double)N * (12 + 4 + 4), rank_flags_kernel<false>, (u32)tiles, kRankThreads, 0, ra);
        LSC_LAUNCH(c, KC_RANK_SCAN, (double)tiles * 24, rank_scan_kernel, kRankScanCtas, 1024, 0, rtagg, tiles, c.d_scalars + S_NACT);
        LSC_LAUNCH(c, KC_RANK_UPDATE, (double)N * (4 + 4 + 8 + 12), rank_apply_kernel<false>, (u32)ceil_div(tiles, kApplyTiles), kRankThreads, 0, ra);
        {   // new ranks -> ISA; the sorted (key, pos) buffer is dead now and serves as partition scratch
            u64 *sorted_k = where ? rk1 : rk0;
            if (partitioned_scatter<NoGen>(c, NoGen(), ra.pair_idx, ra.pair_val, (u32 *)sorted_k, (u32 *)sorted_k + N, N, n, ISA, sort_temp, err, fused_hist) != 0) return -2;
        }
        ra.phist = nullptr;
        const bool checked = local_on && N >= kLocalMin;
        if (checked) launch_big_check(N);
        if (!read_round_scalars(c)) return -2;
        N = c.h_scalars[S_NACT]; G = c.h_scalars[S_NGRP];
        win = checked ? local_window() : 0;
        flags_valid = checked;
        r.n_groups = G;
        c.rounds.push_back(r);
        u32 *t = slot_cur; slot_cur = slot_nxt; slot_nxt = t;
        h *= 2;
        ++round;
    }
    }
    out->SA = SA; out->ISA = ISA; out->isa_complete = !lazy;
    out->primary = c.h_scalars[S_PRIMARY];
    out->scratch = keyA; out->scratch_bytes = (size_t)n * 8;
    return c.failed() ? -2 : 0;
}


// ---------------------------------------------------------------------------------------------
// Building blocks of the distributed prefix doubling (libsais_b200/dist.py): the
// text is replicated, every rank owns a range of positions (ISA slice) and, after the sample
// sort, a range of keys (a slice of the sorted order).  The kernels are the single-GPU ones; the
// exchanges between them are NCCL all-to-alls issued from the host side.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dist_keys_kernel(const u64 *__restrict__ words, u64 n, int b, int k, int K, int len_bits, u64 lo, u64 count,
                 u64 *__restrict__ keys, u32 *__restrict__ pos)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    u64 p = lo + i;
    // the length field (all ones = full length) makes suffixes that run past the end unique and orders them
    // before every longer suffix with the same zero-padded k-mer: no stability is asked of the distributed sort
    u64 len = p + (u64)k <= n ? (((u64)1 << len_bits) - 1) : n - p;
    keys[i] = (kmer_at(words, p, b, K) << len_bits) | len;
    pos[i] = (u32)p;
}

int dist_prepare(Ctx &c, const u8 *d_T, u64 n, int *k_out, int *K_out)
{
    if (n == 0 || n > kMaxN) return -2;
    cudaStream_t st = c.stream;
    run_byte_histogram(c, d_T, n);
    c.check(cudaMemcpyAsync(c.h_scalars + S_FREQ, c.d_scalars + S_FREQ, 256 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    if (!c.sync()) return -2;
    u8 *h_lut = (u8 *)(c.h_scalars + S_MISC);
    int sigma = 0; double entropy = 0;
    for (int s = 0; s < 256; ++s) {
        u64 f = c.h_scalars[S_FREQ + s];
        h_lut[s] = (u8)sigma;
        if (f) { ++sigma; double pr = (double)f / (double)n; entropy -= pr * std::log2(pr); }
    }
    const int b = bits_for((u64)(sigma > 1 ? sigma - 1 : 1));
    // <= 48 key bits + <= 6 length bits leave the top byte of the u64 free for the destination rank of the sample
    // sort.  Unlike the single-GPU heuristic the key is NOT rounded up to whole digit passes: the length field
    // shares the last digit.
    int k;
    {
        int kmax = 48 / b; if (kmax < 1) kmax = 1;
        double need = std::log2((double)(n < 2 ? 2 : n)) + 10.0;
        double kk = std::ceil(need / (entropy < 0.05 ? 0.05 : entropy));
        k = kk > (double)kmax ? kmax : (int)kk;
        if (k < 1) k = 1;
        const char *env = getenv("LIBSAIS_CUDA_KEY_SYMBOLS");
        if (env && *env) { int kv = atoi(env); if (kv >= 1) k = kv < kmax ? kv : kmax; }
    }
    const u64 nwords = ceil_div(n * (u64)b, 64) + 2;
    if (c.dist_words) { cudaFree(c.dist_words); c.dist_words = nullptr; }
    if (cudaMalloc(&c.dist_words, nwords * 8 + 256) != cudaSuccess) { cudaGetLastError(); return -2; }
    u8 *d_lut = (u8 *)(c.dist_words + nwords);
    c.check(cudaMemcpyAsync(d_lut, h_lut, 256, cudaMemcpyHostToDevice, st));
    LSC_LAUNCH(c, KC_PACK, (double)n + (double)nwords * 8, (pack_kernel<u8, true>), (u32)ceil_div(nwords, 256), 256, 0, d_T, n, b, c.dist_words, nwords, d_lut);
    c.dist_n = n; c.dist_b = b; c.dist_k = k;
    *k_out = k; *K_out = k * b;
    return c.sync() && !c.failed() ? 0 : -2;
}

int dist_keys(Ctx &c, u64 lo, u64 count, u64 *d_keys, u32 *d_pos)
{
    if (!c.dist_words || lo + count > c.dist_n) return -1;
    if (count) LSC_LAUNCH(c, KC_MAKE_KEYS, (double)count * 13, dist_keys_kernel, (u32)ceil_div(count, 256), 256, 0,
                          c.dist_words, c.dist_n, c.dist_b, c.dist_k, c.dist_k * c.dist_b, bits_for((u64)c.dist_k), lo, count, d_keys, d_pos);
    return c.failed() ? -2 : 0;
}

size_t rank_stage_workspace_bytes(u64 count)
{
    const u64 tiles = ceil_div(count, kRankTile);
    return 3 * ceil_div(count, 32) * 4 + tiles * (kRankWarps + 1) * 3 * 4 + 4096;
}

// Rank stage on a sorted slice: element j sits in global slot slot_base + j (slot_in == nullptr) or
// slot_in[j].  Emits (pos, rank) pairs for every element, the slice of the suffix array, and the
// compacted active suffixes.  counts[0..1] (host) = active suffixes, active groups.
int run_rank_stage(Ctx &c, const u64 *d_keys, const u32 *d_pos, const u32 *d_slot_in, u64 count, u32 slot_base,
                   u32 *d_sa_local, u32 *d_pair_pos, u32 *d_pair_rank, u32 *d_act_pos, u32 *d_act_slot, u32 *d_act_grp,
                   u64 *counts)
{
    counts[0] = counts[1] = 0;
    if (count == 0) return 0;
    const u64 tiles = ceil_div(count, kRankTile);
    u32 *masks = c.alloc_n<u32>(3 * ceil_div(count, 32));
    u32 *wagg = c.alloc_n<u32>(tiles * kRankWarps * 3);
    u32 *tagg = c.alloc_n<u32>(tiles * 3);
    if (!masks || !wagg || !tagg) return -2;
    RankArgs ra;
    ra.keys = d_keys; ra.pos = d_pos; ra.slot_in = d_slot_in; ra.N = count; ra.tail_start = ~0ull; ra.key_shift = 0;
    ra.slot_base = slot_base;
    ra.SA = d_sa_local; ra.ISA = nullptr; ra.isa_all = 1; ra.pair_idx = d_pair_pos; ra.pair_val = d_pair_rank;
    ra.rows = nullptr; ra.text = nullptr; ra.aux_mask = 0; ra.aux_shift = 0; ra.aux_I = nullptr; ra.phist = nullptr; ra.pshift = 0;
    ra.primary = c.d_scalars + S_PRIMARY;
    ra.a_pos = d_act_pos; ra.a_slot = d_act_slot; ra.a_grp = d_act_grp;
    ra.masks = masks; ra.wagg = wagg; ra.tagg = tagg; ra.nchunks = ceil_div(count, 32); ra.ntiles = tiles;
    ra.out_counts = c.d_scalars + S_NACT;
    if (d_slot_in == nullptr) {
        LSC_LAUNCH(c, KC_RANK_INIT, (double)count * 16, rank_flags_kernel<true>, (u32)tiles, kRankThreads, 0, ra);
        LSC_LAUNCH(c, KC_RANK_SCAN, (double)tiles * 24, rank_scan_kernel, kRankScanCtas, 1024, 0, tagg, tiles, c.d_scalars + S_NACT);
        LSC_LAUNCH(c, KC_RANK_INIT, (double)count * 24, rank_apply_kernel<true>, (u32)ceil_div(tiles, kApplyTiles), kRankThreads, 0, ra);
    } else {
        LSC_LAUNCH(c, KC_RANK_UPDATE, (double)count * 20, rank_flags_kernel<false>, (u32)tiles, kRankThreads, 0, ra);
        LSC_LAUNCH(c, KC_RANK_SCAN, (double)tiles * 24, rank_scan_kernel, kRankScanCtas, 1024, 0, tagg, tiles, c.d_scalars + S_NACT);
        LSC_LAUNCH(c, KC_RANK_UPDATE, (double)count * 28, rank_apply_kernel<false>, (u32)ceil_div(tiles, kApplyTiles), kRankThreads, 0, ra);
    }
    c.check(cudaMemcpyAsync(c.h_scalars + S_NACT, c.d_scalars + S_NACT, 2 * sizeof(u64), cudaMemcpyDeviceToHost, c.stream));
    if (!c.sync() || c.failed()) return -2;
    counts[0] = c.h_scalars[S_NACT]; counts[1] = c.h_scalars[S_NGRP];
    return 0;
}

size_t sort_workspace_bytes(u64 count) { return RadixSort<u64, u32>::temp_bytes(count) + 4096; }

int run_sort_pairs(Ctx &c, u64 *ka, u32 *va, u64 *kb, u32 *vb, u64 count, int lo_bit, int hi_bit)
{
    void *temp = c.alloc(RadixSort<u64, u32>::temp_bytes(count));
    if (!temp) return -2;
    u32 *err = (u32 *)(c.d_scalars + S_ERR);
    c.check(cudaMemsetAsync(c.d_scalars + S_ERR, 0, sizeof(u64), c.stream));
    int where = RadixSort<u64, u32>::sort(c, ka, va, kb, vb, count, lo_bit, hi_bit, temp, err);
    return where < 0 ? -2 : where;
}

int run_sort_u32_pairs(Ctx &c, u32 *ka, u32 *va, u32 *kb, u32 *vb, u64 count, int lo_bit, int hi_bit)
{
    void *temp = c.alloc(RadixSort<u32, u32>::temp_bytes(count));
    if (!temp) return -2;
    u32 *err = (u32 *)(c.d_scalars + S_ERR);
    c.check(cudaMemsetAsync(c.d_scalars + S_ERR, 0, sizeof(u64), c.stream));
    c.pass_class_override = KC_SCATTER;
    int where = RadixSort<u32, u32>::sort(c, ka, va, kb, vb, count, lo_bit, hi_bit, temp, err);
    c.pass_class_override = -1;
    return where < 0 ? -2 : where;
}

__global__ void __launch_bounds__(256)
gather_u32_kernel(const u32 *__restrict__ src, u64 src_len, const u32 *__restrict__ idx, u64 count, u32 idx_offset, u32 *__restrict__ out)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    u64 j = (u64)(idx[i] - idx_offset);
    out[i] = j < src_len ? src[j] : 0;
}
__global__ void __launch_bounds__(256)
scatter_u32_kernel(u32 *__restrict__ dst, u64 dst_len, const u32 *__restrict__ idx, const u32 *__restrict__ val, u64 count, u32 idx_offset)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    u64 j = (u64)(idx[i] - idx_offset);
    if (j < dst_len) dst[j] = val[i];
}
void run_gather_u32(Ctx &c, const u32 *src, u64 src_len, const u32 *idx, u64 count, u32 idx_offset, u32 *out)
{
    if (count) LSC_LAUNCH(c, KC_ROUND_KEYS, (double)count * 12, gather_u32_kernel, (u32)ceil_div(count, 256), 256, 0, src, src_len, idx, count, idx_offset, out);
}
size_t scatter_workspace_bytes(u64 count) { return (size_t)count * 8 + RadixSort<u32, u32>::temp_bytes(count) + 4096; }

// dst[idx[i] - idx_offset] = val[i].  Large scatters go through the locality partition (scatter.cuh);
// the partition digit is taken from idx itself, which is fine because idx - idx_offset preserves locality.
void run_scatter_u32(Ctx &c, u32 *dst, u64 dst_len, const u32 *idx, const u32 *val, u64 count, u32 idx_offset)
{
    if (!count) return;
    if (count >= (1ull << 20) && dst_len > (8ull << 20)) {
        u32 *ib = c.alloc_n<u32>(count), *vb = c.alloc_n<u32>(count);
        void *temp = c.alloc(RadixSort<u32, u32>::temp_bytes(count));
        if (ib && vb && temp) {
            u32 *err = (u32 *)(c.d_scalars + S_ERR);
            c.check(cudaMemsetAsync(c.d_scalars + S_ERR, 0, sizeof(u64), c.stream));
            // one digit pass on the top 8 bits of the index range [idx_offset, idx_offset + dst_len)
            const int bits = bits_for((u64)idx_offset + dst_len - 1);
            const int lo = bits > kRadixBits ? bits - kRadixBits : 0;
            c.pass_class_override = KC_SCATTER;
            int where = RadixSort<u32, u32>::sort(c, const_cast<u32 *>(idx), const_cast<u32 *>(val), ib, vb, count, lo, bits, temp, err);
            c.pass_class_override = -1;
            if (where == 1) {
                LSC_LAUNCH(c, KC_SCATTER, (double)count * 12, scatter_u32_kernel, (u32)ceil_div(count, 256), 256, 0, dst, dst_len, ib, vb, count, idx_offset);
                return;
            }
        }
        c.last_error = cudaSuccess;          // no scratch: fall through to the plain scatter
    }
    LSC_LAUNCH(c, KC_SCATTER, (double)count * 12, scatter_u32_kernel, (u32)ceil_div(count, 256), 256, 0, dst, dst_len, idx, val, count, idx_offset);
}


// ---- routing for the all-to-alls: items are grouped by destination rank with ONE onesweep digit pass.
// The destination rides in bits [56, 64) of a u64 key whose low 32 bits carry the first payload word.
static const int kRouteMaxRanks = 64;

__global__ void __launch_bounds__(256)
route_pack_kernel(const u32 *__restrict__ a, u64 count, u64 add, u64 limit, u64 block, u32 world,
                  u64 *__restrict__ keys, u64 *__restrict__ dest_counts)
{
    __shared__ u32 sh[kRouteMaxRanks + 1];
    if (threadIdx.x <= kRouteMaxRanks) sh[threadIdx.x] = 0;
    __syncthreads();
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < count) {
        u64 v = (u64)a[i] + add;                              // position (add = 0) or requested position p + h
        u64 owner = v < limit ? v / block : (u64)world;       // beyond the text: dropped (sorted behind every rank)
        if (owner > (u64)world) owner = world;
        if (v < limit && owner >= world) owner = world - 1;
        keys[i] = (owner << 56) | (v & 0xFFFFFFFFull);
        atomicAdd(&sh[owner], 1u);
    }
    __syncthreads();
    if (threadIdx.x <= world && sh[threadIdx.x]) atomicAdd((unsigned long long *)&dest_counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
route_unpack_kernel(const u64 *__restrict__ keys, u64 count, u32 *__restrict__ a_out)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < count) a_out[i] = (u32)keys[i];
}

// keys |= (number of splitters <= key) << 56, per-destination counts
__global__ void __launch_bounds__(256)
splitter_dest_kernel(u64 *__restrict__ keys, u64 count, const u64 *__restrict__ splitters, u32 nsplit, u64 *__restrict__ dest_counts)
{
    __shared__ u32 sh[kRouteMaxRanks + 1];
    __shared__ u64 sp[kRouteMaxRanks];
    if (threadIdx.x <= kRouteMaxRanks) sh[threadIdx.x] = 0;
    if (threadIdx.x < nsplit) sp[threadIdx.x] = splitters[threadIdx.x];
    __syncthreads();
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < count) {
        u64 k = keys[i];
        u32 d = 0;
        for (u32 j = 0; j < nsplit; ++j) d += sp[j] <= k ? 1u : 0u;
        keys[i] = k | ((u64)d << 56);
        atomicAdd(&sh[d], 1u);
    }
    __syncthreads();
    if (threadIdx.x <= nsplit && sh[threadIdx.x]) atomicAdd((unsigned long long *)&dest_counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

size_t route_workspace_bytes(u64 count) { return (size_t)count * (8 + 8 + 4) + RadixSort<u64, u32>::temp_bytes(count) + 8192; }

// Group (a, b) items by owner((a + add) / block): outputs in destination order, counts_out[0..world) on the host.
int run_route(Ctx &c, const u32 *d_a, const u32 *d_b, u64 count, u64 add, u64 limit, u64 block, u32 world,
              u32 *d_a_out, u32 *d_b_out, u64 *counts_out)
{
    if (world == 0 || world > kRouteMaxRanks) return -1;        // before touching counts_out: callers size it for <= 64 ranks
    for (u32 r = 0; r < world; ++r) counts_out[r] = 0;
    if (count == 0) return 0;
    u64 *keyA = c.alloc_n<u64>(count), *keyB = c.alloc_n<u64>(count);
    u32 *valA = c.alloc_n<u32>(count);
    void *temp = c.alloc(RadixSort<u64, u32>::temp_bytes(count));
    if (!keyA || !keyB || !valA || !temp) return -2;
    u64 *dcnt = c.d_scalars + S_ROUTE;
    u32 *err = (u32 *)(c.d_scalars + S_ERR);
    c.check(cudaMemsetAsync(c.d_scalars + S_ERR, 0, sizeof(u64), c.stream));
    c.check(cudaMemsetAsync(dcnt, 0, (kRouteMaxRanks + 1) * sizeof(u64), c.stream));
    c.check(cudaMemcpyAsync(valA, d_b, count * 4, cudaMemcpyDeviceToDevice, c.stream));
    LSC_LAUNCH(c, KC_SCATTER, (double)count * 12, route_pack_kernel, (u32)ceil_div(count, 256), 256, 0, d_a, count, add, limit, block, world, keyA, dcnt);
    c.pass_class_override = KC_SCATTER;
    int where = RadixSort<u64, u32>::sort(c, keyA, valA, keyB, d_b_out, count, 56, 56 + bits_for(world), temp, err);
    c.pass_class_override = -1;
    if (where != 1) return -2;
    LSC_LAUNCH(c, KC_SCATTER, (double)count * 12, route_unpack_kernel, (u32)ceil_div(count, 256), 256, 0, keyB, count, d_a_out);
    c.check(cudaMemcpyAsync(c.h_scalars + S_ROUTE, dcnt, (kRouteMaxRanks + 1) * sizeof(u64), cudaMemcpyDeviceToHost, c.stream));
    if (!c.sync() || c.failed()) return -2;
    for (u32 r = 0; r < world; ++r) counts_out[r] = c.h_scalars[S_ROUTE + r];
    return 0;
}

// Round-0 sample sort: destination = number of splitters <= key; one digit pass groups (key, pos) by destination.
// Result in (d_keys_out, d_pos_out); the keys keep the destination in their top byte.
int run_partition_by_splitters(Ctx &c, u64 *d_keys, u32 *d_pos, u64 count, const u64 *d_splitters, u32 nsplit,
                               u64 *d_keys_out, u32 *d_pos_out, u64 *counts_out)
{
    if (nsplit >= kRouteMaxRanks) return -1;                    // before touching counts_out (65 entries at the caller)
    for (u32 r = 0; r <= nsplit; ++r) counts_out[r] = 0;
    if (count == 0) return 0;
    void *temp = c.alloc(RadixSort<u64, u32>::temp_bytes(count));
    if (!temp) return -2;
    u64 *dcnt = c.d_scalars + S_ROUTE;
    u32 *err = (u32 *)(c.d_scalars + S_ERR);
    c.check(cudaMemsetAsync(c.d_scalars + S_ERR, 0, sizeof(u64), c.stream));
    c.check(cudaMemsetAsync(dcnt, 0, (kRouteMaxRanks + 1) * sizeof(u64), c.stream));
    LSC_LAUNCH(c, KC_SCATTER, (double)count * 16, splitter_dest_kernel, (u32)ceil_div(count, 256), 256, 0, d_keys, count, d_splitters, nsplit, dcnt);
    int where;
    if (nsplit == 0) {
        c.check(cudaMemcpyAsync(d_keys_out, d_keys, count * 8, cudaMemcpyDeviceToDevice, c.stream));
        c.check(cudaMemcpyAsync(d_pos_out, d_pos, count * 4, cudaMemcpyDeviceToDevice, c.stream));
        where = 1;
    } else {
        c.pass_class_override = KC_SCATTER;
        where = RadixSort<u64, u32>::sort(c, d_keys, d_pos, d_keys_out, d_pos_out, count, 56, 56 + bits_for(nsplit), temp, err);
        c.pass_class_override = -1;
    }
    if (where != 1) return -2;
    c.check(cudaMemcpyAsync(c.h_scalars + S_ROUTE, dcnt, (kRouteMaxRanks + 1) * sizeof(u64), cudaMemcpyDeviceToHost, c.stream));
    if (!c.sync() || c.failed()) return -2;
    for (u32 r = 0; r <= nsplit; ++r) counts_out[r] = c.h_scalars[S_ROUTE + r];
    return 0;
}

}  // namespace lsc
