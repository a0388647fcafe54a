// post.cu -- the stages that consume a suffix array: BWT (+primary, aux), phi / PLCP / LCP,
// and the inverse BWT.  Definitions follow SURVEY.md §8a; each kernel cites the reference
// region whose result it reproduces bit-exactly.
#include "core.h"
#include "radix_sort.cuh"
#include "scatter.cuh"

namespace lsc {

// ---------------------------------------------------------------------------------------------
// BWT assembly.  Reference: final_bwt_scan_* + bwt_copy_8u, src/libsais.c:4777-4797, :5396-5422,
// :6960-7006, assembly :7110-7118.  The SA core already produced rows[i] = T[SA[i]-1] for every
// slot (the byte rides through the sort in the key's low bits); with p0 = primary-1 the slot of
// suffix 0:  U[0] = T[n-1];  slot i != p0 goes to U[i + (i < p0)]  (the "$" row is dropped).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bwt_finish_kernel(const u8 *__restrict__ T, const u8 *__restrict__ rows, u8 *__restrict__ U, u64 n, u64 p0)
{
    // 16 output bytes per thread: U[o] = rows[o - 1] for 1 <= o <= p0, rows[o] for o > p0, U[0] = T[n-1]
    const u64 o0 = ((u64)blockIdx.x * 256 + threadIdx.x) * 16;
    if (o0 >= n) return;
    const bool aligned = (((uintptr_t)U | (uintptr_t)rows) & 15) == 0;
    if (aligned && o0 + 16 <= n && o0 > p0) {                       // whole group after the dropped row: straight copy
        *reinterpret_cast<uint4 *>(U + o0) = *reinterpret_cast<const uint4 *>(rows + o0);
        return;
    }
    if (aligned && o0 + 16 <= n && o0 + 15 <= p0 && o0 >= 16) {     // whole group before it: shifted by one byte
        const uint4 a = *reinterpret_cast<const uint4 *>(rows + o0 - 16), c = *reinterpret_cast<const uint4 *>(rows + o0);
        uint4 r;
        r.x = __funnelshift_l(a.w, c.x, 8); r.y = __funnelshift_l(c.x, c.y, 8);
        r.z = __funnelshift_l(c.y, c.z, 8); r.w = __funnelshift_l(c.z, c.w, 8);
        *reinterpret_cast<uint4 *>(U + o0) = r;
        return;
    }
    for (u64 o = o0; o < o0 + 16 && o < n; ++o)
        U[o] = o == 0 ? T[n - 1] : rows[o <= p0 ? o - 1 : o];
}

// Streamed rows (sa_core.cu): the rows of the slots round 0 settled are already in the caller's pinned buffer.  The slots
// that were still open then (a few thousand on random text) are final now: their bytes go straight into the host buffer
// (zero-copy stores through the buffer's device alias), shifted around the dropped row p0.
__global__ void __launch_bounds__(256)
bwt_patch_kernel(const u32 *__restrict__ slots, u64 count, const u8 *__restrict__ rows, u8 *__restrict__ U_host, u64 p0)
{
    const u64 j = (u64)blockIdx.x * 256 + threadIdx.x;
    if (j >= count) return;
    const u64 s = slots[j];
    if (s != p0) U_host[s + (s < p0 ? 1 : 0)] = rows[s];
}

int run_bwt_patch(Ctx &c, const u32 *d_slots, u64 count, const u8 *d_rows, u8 *U_host_dev, u64 p0)
{
    if (count) LSC_LAUNCH(c, KC_BWT, (double)count * 6, bwt_patch_kernel, (u32)ceil_div(count, 256), 256, 0, d_slots, count, d_rows, U_host_dev, p0);
    return c.failed() ? -2 : 0;
}

int run_bwt_finish(Ctx &c, const u8 *d_T, const u8 *d_rows, u8 *d_U, u64 n, u64 primary)
{
    LSC_LAUNCH(c, KC_BWT, (double)n * 2, bwt_finish_kernel, (u32)ceil_div(ceil_div(n, 16), 256), 256, 0, d_T, d_rows, d_U, n, primary - 1);
    return c.failed() ? -2 : 0;
}

// ---------------------------------------------------------------------------------------------
// phi scatter (reference compute_phi :8116-8142): PLCP[SA[i]] = SA[i-1], PLCP[SA[0]] = n.
// The (SA[i], SA[i-1]) pairs are generated on the fly and scattered through the locality
// partition of scatter.cuh (a plain random scatter of n 4-byte words runs at ~20 G/s).
// ---------------------------------------------------------------------------------------------
struct PhiGen {
    static const bool kActive = true;
    const u32 *SA; u64 n;
    __device__ __forceinline__ u64 key(u64 i) const { return SA[i]; }
    __device__ __forceinline__ u32 val(u64 i) const { return i ? SA[i - 1] : (u32)n; }
};

// ---------------------------------------------------------------------------------------------
// PLCP compare (reference compute_plcp :8167-8190, _int :8263-8286): in text order,
// PLCP[i] = lcp(T[i..], T[phi[i]..]).  The Kasai bound PLCP[i] >= PLCP[i-1] - 1 generalises to
// PLCP[i] >= PLCP[i-s] - s, which is what makes the work O(n log n) in the WORST case with
// n-way parallelism (a chunk that restarted at l = 0 cost O(PLCP) per chunk: quadratic on a^n):
//   1. position 0, then levels s = 2^L .. kPlcpChunk: every odd multiple i of s is computed by one
//      WARP (16 bytes per lane and step) starting from the bound given by the already known
//      PLCP[i - s]; per level the compares telescope to <= 2n symbols.  phi[i] is read from
//      PLCP[i] itself, which is overwritten in place.
//   2. the chunk kernel: every thread walks kPlcpChunk consecutive positions from its (now exact)
//      chunk head with the l - 1 carry; the warp's 1024 phi/PLCP values are staged through shared
//      memory so global accesses are coalesced; bytes are compared 16 at a time through
//      unaligned windows built from aligned 8-byte loads.
// ---------------------------------------------------------------------------------------------
static const int kPlcpChunk = 32;
static const int kPlcpThreads = 128;

// bytes [pos, pos + 16) of the text as two little-endian words (text readable up to 24 bytes past pos)
__device__ __forceinline__ void window128(const u64 *__restrict__ W, u64 pos, u64 &a, u64 &b)
{
    const u64 q = pos >> 3; const int sh = (int)(pos & 7) * 8;
    const u64 w0 = W[q], w1 = W[q + 1];
    if (sh == 0) { a = w0; b = w1; return; }
    const u64 w2 = W[q + 2];
    a = (w0 >> sh) | (w1 << (64 - sh));
    b = (w1 >> sh) | (w2 << (64 - sh));
}
// matching bytes at the start of two 16-byte windows (0..16)
__device__ __forceinline__ u32 match16(u64 a0, u64 a1, u64 b0, u64 b1)
{
    const u64 x0 = a0 ^ b0, x1 = a1 ^ b1;
    if (x0) return (u32)((__ffsll((long long)x0) - 1) >> 3);
    if (x1) return 8u + (u32)((__ffsll((long long)x1) - 1) >> 3);
    return 16u;
}

// one thread extends a match of l symbols between suffixes i and k (at most lim symbols can match)
template <int SYM_BYTES>
__device__ __forceinline__ u64 extend_thread(const void *__restrict__ Tv, u64 i, u64 k, u64 l, u64 lim)
{
    if (SYM_BYTES == 1) {
        const u64 *W = (const u64 *)Tv;
        while (l < lim) {
            u64 a0, a1, b0, b1;
            window128(W, i + l, a0, a1); window128(W, k + l, b0, b1);
            const u32 m = match16(a0, a1, b0, b1);
            l += m;
            if (m < 16) break;
        }
    } else {
        const u32 *S = (const u32 *)Tv;
        while (l < lim && S[i + l] == S[k + l]) ++l;
    }
    return l < lim ? l : lim;
}

// the same by a whole warp: 512 bytes (32 symbols for integer texts) per step
template <int SYM_BYTES>
__device__ __forceinline__ u64 extend_warp(const void *__restrict__ Tv, u64 i, u64 k, u64 l, u64 lim, int lane)
{
    const int per = SYM_BYTES == 1 ? 16 : 1;
    while (l < lim) {
        const u64 off = l + (u64)lane * per;
        const bool valid = off < lim;
        u32 m = 0;
        if (valid) {
            if (SYM_BYTES == 1) {
                u64 a0, a1, b0, b1;
                window128((const u64 *)Tv, i + off, a0, a1); window128((const u64 *)Tv, k + off, b0, b1);
                m = match16(a0, a1, b0, b1);
            } else {
                const u32 *S = (const u32 *)Tv;
                m = S[i + off] == S[k + off] ? 1u : 0u;
            }
        }
        const u32 stop = __ballot_sync(0xffffffffu, !valid || m < (u32)per);
        if (stop == 0) { l += 32ull * per; continue; }
        const int f = __ffs(stop) - 1;
        const u32 mf = __shfl_sync(0xffffffffu, m, f);
        l += (u64)f * per + mf;
        break;
    }
    return l < lim ? l : lim;
}

// level kernel: warp w computes position i = first + w * stride from the bound PLCP[i - back] - back (back == 0: none)
template <int SYM_BYTES>
__global__ void __launch_bounds__(256)
plcp_level_kernel(const void *__restrict__ Tv, u32 *__restrict__ PLCP, u64 n, u64 first, u64 stride, u64 back, u64 count)
{
    const int lane = threadIdx.x & 31;
    const u64 w = ((u64)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (w >= count) return;
    const u64 i = first + w * stride;
    if (i >= n) return;
    const u64 k = PLCP[i];
    u64 l = 0;
    if (k < n) {
        if (back) { const u64 pv = PLCP[i - back]; l = pv > back ? pv - back : 0; }
        const u64 lim = n - (i > k ? i : k);
        l = extend_warp<SYM_BYTES>(Tv, i, k, l < lim ? l : lim, lim, lane);
    }
    __syncwarp();
    if (lane == 0) PLCP[i] = (u32)l;
}

// chunk kernel: thread t of the CTA walks positions [c0 + 32 t, c0 + 32 t + 32); the head of every chunk is final
template <int SYM_BYTES>
__global__ void __launch_bounds__(kPlcpThreads)
plcp_chunk_kernel(const void *__restrict__ Tv, u32 *__restrict__ PLCP, u64 n)
{
    __shared__ u32 sh[kPlcpThreads * (kPlcpChunk + 1)];
    const int tid = threadIdx.x;
    const u64 c0 = (u64)blockIdx.x * (kPlcpThreads * kPlcpChunk);
    for (int j = tid; j < kPlcpThreads * kPlcpChunk; j += kPlcpThreads) {
        const u64 i = c0 + j;
        sh[(j >> 5) * (kPlcpChunk + 1) + (j & 31)] = i < n ? PLCP[i] : 0u;
    }
    __syncthreads();
    const u64 begin = c0 + (u64)tid * kPlcpChunk;
    if (begin < n) {
        u32 *mine = sh + tid * (kPlcpChunk + 1);
        u64 l = mine[0];
        const u64 end = begin + kPlcpChunk < n ? begin + kPlcpChunk : n;
        for (u64 i = begin + 1; i < end; ++i) {
            if (l) --l;
            const u64 k = mine[i - begin];
            if (k >= n) l = 0;
            else {
                const u64 lim = n - (i > k ? i : k);
                l = extend_thread<SYM_BYTES>(Tv, i, k, l < lim ? l : lim, lim);
            }
            mine[i - begin] = (u32)l;
        }
    }
    __syncthreads();
    for (int j = tid; j < kPlcpThreads * kPlcpChunk; j += kPlcpThreads) {
        const u64 i = c0 + j;
        if (i < n) PLCP[i] = sh[(j >> 5) * (kPlcpChunk + 1) + (j & 31)];
    }
}

size_t plcp_workspace_bytes(u64 n) { return (size_t)n * 8 + RadixSort<u32, u32>::temp_bytes(n) + 4096; }

template <int SYM_BYTES>
static void run_plcp_compare(Ctx &c, const void *d_T, u32 *d_PLCP, u64 n)
{
    const double per_level = 12.0;
    // position 0 has no predecessor
    LSC_LAUNCH(c, KC_PLCP, per_level, plcp_level_kernel<SYM_BYTES>, 1, 256, 0, d_T, d_PLCP, n, (u64)0, (u64)1, (u64)0, (u64)1);
    if (n > (u64)kPlcpChunk) {
        u64 s = kPlcpChunk;
        while (s * 2 < n) s *= 2;                                  // largest chunk multiple 2^L < n
        for (; s >= (u64)kPlcpChunk; s >>= 1) {
            const u64 count = (n - 1 - s) / (2 * s) + 1;           // odd multiples of s below n
            LSC_LAUNCH(c, KC_PLCP, (double)count * per_level, plcp_level_kernel<SYM_BYTES>, (u32)ceil_div(count * 32, 256), 256, 0,
                       d_T, d_PLCP, n, s, 2 * s, s, count);
        }
    }
    const double ab = (double)n * (SYM_BYTES == 1 ? 11.0 : 16.0);
    LSC_LAUNCH(c, KC_PLCP, ab, plcp_chunk_kernel<SYM_BYTES>, (u32)ceil_div(n, (u64)kPlcpThreads * kPlcpChunk), kPlcpThreads, 0, d_T, d_PLCP, n);
}

int run_plcp(Ctx &c, const void *d_T, int sym_bytes, const u32 *d_SA, u32 *d_PLCP, u64 n)
{
    {
        u32 *ib = c.alloc_n<u32>(n), *vb = c.alloc_n<u32>(n);
        void *temp = c.alloc(RadixSort<u32, u32>::temp_bytes(n));
        if (!ib || !vb || !temp) return -2;
        u32 *err = (u32 *)(c.d_scalars + S_ERR);
        c.check(cudaMemsetAsync(c.d_scalars + S_ERR, 0, sizeof(u64), c.stream));
        PhiGen gen; gen.SA = d_SA; gen.n = n;
        c.pass_class_override = KC_PHI;
        int rc = partitioned_scatter<PhiGen>(c, gen, nullptr, nullptr, ib, vb, n, n, d_PLCP, temp, err);
        c.pass_class_override = -1;
        if (rc != 0) return -2;
    }
    if (sym_bytes == 1) run_plcp_compare<1>(c, d_T, d_PLCP, n);
    else run_plcp_compare<4>(c, d_T, d_PLCP, n);
    return c.failed() ? -2 : 0;
}

// LCP permute (reference compute_lcp :8311-8338): LCP[i] = PLCP[SA[i]]
__global__ void __launch_bounds__(256)
lcp_kernel(const u32 *__restrict__ PLCP, const u32 *__restrict__ SA, u32 *__restrict__ LCP, u64 n)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < n) { u32 s = ld_stream(SA + i); LCP[i] = s < n ? PLCP[s] : 0; }
}

int run_lcp(Ctx &c, const u32 *d_PLCP, const u32 *d_SA, u32 *d_LCP, u64 n)
{
    LSC_LAUNCH(c, KC_LCP, (double)n * 12, lcp_kernel, (u32)ceil_div(n, 256), 256, 0, d_PLCP, d_SA, d_LCP, n);
    return c.failed() ? -2 : 0;
}

// ---------------------------------------------------------------------------------------------
// Inverse BWT.  Reference: libsais_unbwt_* src/libsais.c:7362-8112 (bigram psi + a sequential
// pointer chase).  Here: rows 0..n are the sorted rotations of T$, the "$" row is `primary`,
// L'[row] = B[row - (row > primary)].  One stable counting-sort pass of the row ids by symbol
// (the onesweep pass) yields psi: psi[j+1] = j-th row in symbol order, F[j+1] = its symbol.
// The text is the walk x0 = primary, x(t+1) = psi[x(t)], T[t] = F[x(t)], ending at row 0.
// The single n-step chain is cut at splitter rows (row % S == 0, plus the start): every
// splitter walks to the next one (phase 1), the splitters are list-ranked by pointer jumping
// (phase 2), and every splitter re-walks its piece writing text at its now-known offset.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
unbwt_rows_kernel(u32 *__restrict__ rows, u64 n, u64 primary)
{
    u64 e = (u64)blockIdx.x * 256 + threadIdx.x;
    if (e < n) rows[e] = (u32)(e + (e >= primary ? 1 : 0));
}

static const u32 kEndMark = 0xFFFFFFFFu;

__device__ __forceinline__ bool unbwt_stop(u64 x, u64 smask, u64 primary)
{
    return x == 0 || (x & smask) == 0 || x == primary;
}

// phase 1: splitter id 0 = the start row `primary`; id s >= 1 = row s*S.
__global__ void __launch_bounds__(128)
unbwt_walk1_kernel(const u32 *__restrict__ psi, u64 n, u64 primary, int logS, u64 nsplit,
                   u32 *__restrict__ nxt, u32 *__restrict__ len)
{
    u64 id = (u64)blockIdx.x * 128 + threadIdx.x;
    if (id >= nsplit) return;
    const u64 smask = ((u64)1 << logS) - 1;
    u64 x = id ? (id << logS) : primary;
    u32 steps = 0;
    do { x = psi[x]; ++steps; } while (!unbwt_stop(x, smask, primary) && steps <= n);
    len[id] = steps;
    nxt[id] = (x == 0) ? kEndMark : (x == primary ? 0u : (u32)(x >> logS));
}

// phase 2: one pointer-jumping round (double buffered): dist = characters from here to the end
__global__ void __launch_bounds__(256)
unbwt_jump_kernel(const u32 *__restrict__ nxt_in, const u64 *__restrict__ dist_in,
                  u32 *__restrict__ nxt_out, u64 *__restrict__ dist_out, u64 nsplit)
{
    u64 id = (u64)blockIdx.x * 256 + threadIdx.x;
    if (id >= nsplit) return;
    u32 nx = nxt_in[id]; u64 d = dist_in[id];
    if (nx != kEndMark) { d += dist_in[nx]; nx = nxt_in[nx]; }
    nxt_out[id] = nx; dist_out[id] = d;
}

__global__ void __launch_bounds__(256)
unbwt_dist_init_kernel(const u32 *__restrict__ len, u64 *__restrict__ dist, u64 nsplit)
{
    u64 id = (u64)blockIdx.x * 256 + threadIdx.x;
    if (id < nsplit) dist[id] = len[id];
}

// phase 3: re-walk, emit text.  cum[c] = 1 + #{symbols < c} (row of the first rotation starting with c)
__global__ void __launch_bounds__(128)
unbwt_walk2_kernel(const u32 *__restrict__ psi, const u64 *__restrict__ base, u64 n, u64 primary,
                   int logS, u64 nsplit, const u64 *__restrict__ dist, u8 *__restrict__ U)
{
    __shared__ u64 cum[257];
    for (int i = threadIdx.x; i < 256; i += 128) cum[i] = base[i] + 1;
    if (threadIdx.x == 0) cum[256] = n + 1;
    __syncthreads();
    u64 id = (u64)blockIdx.x * 128 + threadIdx.x;
    if (id >= nsplit) return;
    const u64 smask = ((u64)1 << logS) - 1;
    u64 x = id ? (id << logS) : primary;
    u64 d = dist[id];
    if (d > n) return;                                   // inconsistent input: never write out of bounds
    u64 t = n - d;
    u64 steps = 0;
    do {
        // F[x]: largest c with cum[c] <= x
        int lo = 0, hi = 256;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cum[mid] <= x) lo = mid; else hi = mid; }
        if (t < n) U[t] = (u8)lo;
        ++t; ++steps;
        x = psi[x];
    } while (!unbwt_stop(x, smask, primary) && steps <= n);
}

// aux samples hand every r-th text position its row for free (I[j] = row of suffix j*r, reference src/libsais.c:7943-7973 decodes
// the blocks independently too): chain j starts at row I[j] and emits U[j*r .. j*r + r) -- one walk, no list ranking
__global__ void __launch_bounds__(128)
unbwt_walk_aux_kernel(const u32 *__restrict__ psi, const u64 *__restrict__ base, u64 n, u64 r, const u32 *__restrict__ I, u64 nchains, u8 *__restrict__ U)
{
    __shared__ u64 cum[257];
    for (int i = threadIdx.x; i < 256; i += 128) cum[i] = base[i] + 1;
    if (threadIdx.x == 0) cum[256] = n + 1;
    __syncthreads();
    const u64 j = (u64)blockIdx.x * 128 + threadIdx.x;
    if (j >= nchains) return;
    u64 x = I[j];
    const u64 t0 = j * r, t1 = t0 + r < n ? t0 + r : n;
    for (u64 t = t0; t < t1; ++t) {
        if (x > n) return;                               // inconsistent samples: never read out of bounds
        int lo = 0, hi = 256;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (cum[mid] <= x) lo = mid; else hi = mid; }
        U[t] = (u8)lo;
        x = psi[x];
    }
}

static int unbwt_log_s(u64 n) { return n >= (1ull << 24) ? 7 : 5; }

size_t unbwt_workspace_bytes(u64 n)
{
    u64 ns = (n >> unbwt_log_s(n)) + 2;
    return (size_t)n * (4 + 4 + 1) + 8 + RadixSort<u8, u32>::temp_bytes(n) + ns * (4 + 4 + 4 + 8 + 8) + 16 * 256;
}

int run_unbwt(Ctx &c, const u8 *d_B, u8 *d_U, u64 n, u64 primary, u64 aux_r, const u32 *d_I, u64 n_aux)
{
    const int logS = unbwt_log_s(n);
    const u64 nsplit = (n >> logS) + 1;
    u32 *rows = c.alloc_n<u32>(n);
    u32 *psi = c.alloc_n<u32>(n + 1);
    u8 *keys_out = c.alloc_n<u8>(n);
    void *temp = c.alloc(RadixSort<u8, u32>::temp_bytes(n));
    u32 *nxt0 = c.alloc_n<u32>(nsplit), *nxt1 = c.alloc_n<u32>(nsplit), *len = c.alloc_n<u32>(nsplit);
    u64 *dist0 = c.alloc_n<u64>(nsplit), *dist1 = c.alloc_n<u64>(nsplit);
    if (!rows || !psi || !keys_out || !temp || !dist1) return -2;
    u32 *err = (u32 *)(c.d_scalars + S_ERR);
    c.check(cudaMemsetAsync(c.d_scalars + S_ERR, 0, sizeof(u64), c.stream));
    c.check(cudaMemsetAsync(psi, 0, sizeof(u32), c.stream));

    LSC_LAUNCH(c, KC_UNBWT_PREP, (double)n * 4, unbwt_rows_kernel, (u32)ceil_div(n, 256), 256, 0, rows, n, primary);
    // one stable 8-bit pass: (symbol, row) -> psi[1..n]; the pass's digit bases are 1-based F boundaries
    int where = RadixSort<u8, u32>::sort(c, const_cast<u8 *>(d_B), rows, keys_out, psi + 1, n, 0, 8, temp, err);
    if (where != 1) return -2;
    const u64 *base = (const u64 *)((char *)temp + kMaxPasses * kRadixSize * sizeof(u64));

    // with dense enough aux samples every block is an independent chain (enough chains to hide the dependent loads)
    if (d_I != nullptr && aux_r >= 2 && n_aux >= 2 && (aux_r <= 512 || n_aux >= 200000 || n < ((u64)1 << 20))) {
        LSC_LAUNCH(c, KC_UNBWT_WALK, (double)n * 5, unbwt_walk_aux_kernel, (u32)ceil_div(n_aux, 128), 128, 0, psi, base, n, aux_r, d_I, n_aux, d_U);
        return c.failed() ? -2 : 0;
    }
    LSC_LAUNCH(c, KC_UNBWT_WALK, (double)n * 4, unbwt_walk1_kernel, (u32)ceil_div(nsplit, 128), 128, 0,
               psi, n, primary, logS, nsplit, nxt0, len);
    LSC_LAUNCH(c, KC_UNBWT_RANK, (double)nsplit * 12, unbwt_dist_init_kernel, (u32)ceil_div(nsplit, 256), 256, 0, len, dist0, nsplit);
    int rounds = bits_for(nsplit) + 1;
    u32 *ni = nxt0, *no = nxt1; u64 *di = dist0, *dout = dist1;
    for (int r = 0; r < rounds; ++r) {
        LSC_LAUNCH(c, KC_UNBWT_RANK, (double)nsplit * 24, unbwt_jump_kernel, (u32)ceil_div(nsplit, 256), 256, 0,
                   ni, di, no, dout, nsplit);
        u32 *tn = ni; ni = no; no = tn;
        u64 *td = di; di = dout; dout = td;
    }
    LSC_LAUNCH(c, KC_UNBWT_WALK, (double)n * 5, unbwt_walk2_kernel, (u32)ceil_div(nsplit, 128), 128, 0,
               psi, base, n, primary, logS, nsplit, di, d_U);
    return c.failed() ? -2 : 0;
}

// ---------------------------------------------------------------------------------------------
// index-width conversion for the libsais64 API (reference converters src/libsais64.c:6638-6701)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) widen_kernel(const u32 *__restrict__ src, i64 *__restrict__ dst, u64 n)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = (i64)src[i];
}
__global__ void __launch_bounds__(256) narrow_kernel(const i64 *__restrict__ src, u32 *__restrict__ dst, u64 n)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = (u32)src[i];
}
void run_widen(Ctx &c, const u32 *src, i64 *dst, u64 n)
{
    if (n) LSC_LAUNCH(c, KC_CONVERT, (double)n * 12, widen_kernel, (u32)ceil_div(n, 256), 256, 0, src, dst, n);
}
void run_narrow(Ctx &c, const i64 *src, u32 *dst, u64 n)
{
    if (n) LSC_LAUNCH(c, KC_CONVERT, (double)n * 12, narrow_kernel, (u32)ceil_div(n, 256), 256, 0, src, dst, n);
}

// ---------------------------------------------------------------------------------------------
// 16-bit symbols (reference src/libsais16.c, include/libsais16.h).  The SA / PLCP cores run on the text widened to 32-bit
// symbols (the integer-alphabet path); this section holds what is specific to uint16_t texts: the 65536-bin frequency
// table, BWT rows gathered from the suffix array, and the inverse BWT over a 65536-symbol alphabet (psi by TWO stable
// onesweep passes of (symbol, row); F[x] is read from the sorted symbols instead of a cumulative table).
// ---------------------------------------------------------------------------------------------
typedef uint16_t u16;

__global__ void __launch_bounds__(256) widen16_kernel(const u16 *__restrict__ src, u32 *__restrict__ dst, u64 n)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = (u32)src[i];
}
void run_widen16(Ctx &c, const u16 *src, u32 *dst, u64 n)
{
    if (n) LSC_LAUNCH(c, KC_CONVERT, (double)n * 6, widen16_kernel, (u32)ceil_div(n, 256), 256, 0, src, dst, n);
}

__global__ void __launch_bounds__(256) hist_u16_kernel(const u16 *__restrict__ T, u64 n, unsigned long long *__restrict__ hist)
{
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) atomicAdd(&hist[T[i]], 1ull);
}
void run_hist_u16(Ctx &c, const u16 *d_T, u64 n, u64 *d_hist)
{
    c.check(cudaMemsetAsync(d_hist, 0, 65536 * sizeof(u64), c.stream));
    if (!n) return;
    const u64 want = ceil_div(n, 256 * 8);
    const u32 grid = (u32)(want < (u64)c.sm_count * 16 ? want : (u64)c.sm_count * 16);
    LSC_LAUNCH(c, KC_HIST_SYM, (double)n * 2, hist_u16_kernel, grid, 256, 0, d_T, n, (unsigned long long *)d_hist);
}

// slot of suffix 0 (+1) and the aux samples I[p / r] = slot + 1 for p % r == 0, from a finished suffix array
__global__ void __launch_bounds__(256)
sa_primary_aux_kernel(const u32 *__restrict__ SA, u64 n, u64 aux_mask, int aux_shift, u32 *__restrict__ aux_I, u64 *__restrict__ primary)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const u32 p = SA[i];
    if (p == 0) *primary = i + 1;
    if (aux_I != nullptr && ((u64)p & aux_mask) == 0) aux_I[p >> aux_shift] = (u32)i + 1;
}
// U[0] = T[n-1]; slot i (suffix p != 0) -> U[i + (i < p0)] = T[p - 1]   (reference assembly src/libsais16.c, same rule as :7110-7118 of libsais.c)
__global__ void __launch_bounds__(256)
bwt16_gather_kernel(const u16 *__restrict__ T, const u32 *__restrict__ SA, u16 *__restrict__ U, u64 n, u64 p0)
{
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (i == 0) U[0] = T[n - 1];
    const u32 p = SA[i];
    if (p != 0) U[i + (i < p0 ? 1 : 0)] = T[p - 1];
}
// returns the primary index (>= 1) in *primary_out; d_U must not alias d_T
int run_bwt16(Ctx &c, const u16 *d_T, const u32 *d_SA, u16 *d_U, u64 n, u64 aux_r, u32 *d_I, u64 *primary_out)
{
    u64 *dp = c.d_scalars + S_PRIMARY;
    c.check(cudaMemsetAsync(dp, 0, sizeof(u64), c.stream));
    LSC_LAUNCH(c, KC_BWT, (double)n * 4, sa_primary_aux_kernel, (u32)ceil_div(n, 256), 256, 0, d_SA, n,
               aux_r ? aux_r - 1 : ~0ull, aux_r ? bits_for(aux_r) - 1 : 0, aux_r ? d_I : (u32 *)nullptr, dp);
    c.check(cudaMemcpyAsync(c.h_scalars + S_PRIMARY, dp, sizeof(u64), cudaMemcpyDeviceToHost, c.stream));
    if (!c.sync()) return -2;
    const u64 primary = c.h_scalars[S_PRIMARY];
    if (primary < 1 || primary > n) return -2;
    LSC_LAUNCH(c, KC_BWT, (double)n * 8, bwt16_gather_kernel, (u32)ceil_div(n, 256), 256, 0, d_T, d_SA, d_U, n, primary - 1);
    *primary_out = primary;
    return c.failed() ? -2 : 0;
}

// inverse BWT walks with F read from the sorted symbols: F[x] = sorted[x - 1] for rows x >= 1
__global__ void __launch_bounds__(128)
unbwt16_walk2_kernel(const u32 *__restrict__ psi, const u16 *__restrict__ sorted, u64 n, u64 primary, int logS, u64 nsplit,
                     const u64 *__restrict__ dist, u16 *__restrict__ U)
{
    u64 id = (u64)blockIdx.x * 128 + threadIdx.x;
    if (id >= nsplit) return;
    const u64 smask = ((u64)1 << logS) - 1;
    u64 x = id ? (id << logS) : primary;
    u64 d = dist[id];
    if (d > n) return;
    u64 t = n - d, steps = 0;
    do {
        if (t < n && x >= 1) U[t] = sorted[x - 1];
        ++t; ++steps;
        x = psi[x];
    } while (!unbwt_stop(x, smask, primary) && steps <= n);
}
__global__ void __launch_bounds__(128)
unbwt16_walk_aux_kernel(const u32 *__restrict__ psi, const u16 *__restrict__ sorted, u64 n, u64 r, const u32 *__restrict__ I, u64 nchains, u16 *__restrict__ U)
{
    const u64 j = (u64)blockIdx.x * 128 + threadIdx.x;
    if (j >= nchains) return;
    u64 x = I[j];
    const u64 t0 = j * r, t1 = t0 + r < n ? t0 + r : n;
    for (u64 t = t0; t < t1; ++t) {
        if (x > n || x < 1) return;                      // inconsistent samples
        U[t] = sorted[x - 1];
        x = psi[x];
    }
}

size_t unbwt16_workspace_bytes(u64 n)
{
    u64 ns = (n >> unbwt_log_s(n)) + 2;
    return (size_t)n * (4 + 4 + 2 + 2) + 64 + RadixSort<u16, u32>::temp_bytes(n) + ns * (4 + 4 + 4 + 8 + 8) + 16 * 256;
}

// d_B is CLOBBERED (it ends up holding the sorted symbols); d_U must not alias it
int run_unbwt16(Ctx &c, u16 *d_B, u16 *d_U, u64 n, u64 primary, u64 aux_r, const u32 *d_I, u64 n_aux)
{
    const int logS = unbwt_log_s(n);
    const u64 nsplit = (n >> logS) + 1;
    u32 *pa = c.alloc_n<u32>(n + 1), *pb = c.alloc_n<u32>(n + 1);
    u16 *kb = c.alloc_n<u16>(n);
    void *temp = c.alloc(RadixSort<u16, u32>::temp_bytes(n));
    u32 *nxt0 = c.alloc_n<u32>(nsplit), *nxt1 = c.alloc_n<u32>(nsplit), *len = c.alloc_n<u32>(nsplit);
    u64 *dist0 = c.alloc_n<u64>(nsplit), *dist1 = c.alloc_n<u64>(nsplit);
    if (!pa || !pb || !kb || !temp || !dist1) return -2;
    u32 *err = (u32 *)(c.d_scalars + S_ERR);
    c.check(cudaMemsetAsync(c.d_scalars + S_ERR, 0, sizeof(u64), c.stream));
    c.check(cudaMemsetAsync(pa, 0, sizeof(u32), c.stream));
    c.check(cudaMemsetAsync(pb, 0, sizeof(u32), c.stream));
    LSC_LAUNCH(c, KC_UNBWT_PREP, (double)n * 4, unbwt_rows_kernel, (u32)ceil_div(n, 256), 256, 0, pa + 1, n, primary);
    // two stable 8-bit passes of (symbol, row): psi[1..n] = rows in symbol order, the keys end up sorted (= F[1..n])
    const int where = RadixSort<u16, u32>::sort(c, d_B, pa + 1, kb, pb + 1, n, 0, 16, temp, err);
    if (where < 0) return -2;
    const u32 *psi = where ? pb : pa;
    const u16 *sorted = where ? kb : d_B;
    if (d_I != nullptr && aux_r >= 2 && n_aux >= 2 && (aux_r <= 512 || n_aux >= 200000 || n < ((u64)1 << 20))) {
        LSC_LAUNCH(c, KC_UNBWT_WALK, (double)n * 8, unbwt16_walk_aux_kernel, (u32)ceil_div(n_aux, 128), 128, 0, psi, sorted, n, aux_r, d_I, n_aux, d_U);
        return c.failed() ? -2 : 0;
    }
    LSC_LAUNCH(c, KC_UNBWT_WALK, (double)n * 4, unbwt_walk1_kernel, (u32)ceil_div(nsplit, 128), 128, 0, psi, n, primary, logS, nsplit, nxt0, len);
    LSC_LAUNCH(c, KC_UNBWT_RANK, (double)nsplit * 12, unbwt_dist_init_kernel, (u32)ceil_div(nsplit, 256), 256, 0, len, dist0, nsplit);
    const int rounds = bits_for(nsplit) + 1;
    u32 *ni = nxt0, *no = nxt1; u64 *di = dist0, *dout = dist1;
    for (int r = 0; r < rounds; ++r) {
        LSC_LAUNCH(c, KC_UNBWT_RANK, (double)nsplit * 24, unbwt_jump_kernel, (u32)ceil_div(nsplit, 256), 256, 0, ni, di, no, dout, nsplit);
        u32 *tn = ni; ni = no; no = tn;
        u64 *td = di; di = dout; dout = td;
    }
    LSC_LAUNCH(c, KC_UNBWT_WALK, (double)n * 8, unbwt16_walk2_kernel, (u32)ceil_div(nsplit, 128), 128, 0, psi, sorted, n, primary, logS, nsplit, di, d_U);
    return c.failed() ? -2 : 0;
}

}  // namespace lsc
