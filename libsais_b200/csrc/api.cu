// api.cu -- the C-ABI of libsais_cuda: every symbol of include/libsais.h and include/libsais64.h
// (host pointers; the drop-in boundary, reference src/libsais.c:7008-7322, :8020-8112,
// :8363-8517 and src/libsais64.c:7058-7250, :8034-8400) plus the device-pointer extras of
// include/libsais_cuda.h.  Argument validation, n <= 1 fast paths and return codes mirror the
// reference; all computing is done by the CUDA pipelines -- there is no CPU
// fallback: without a usable GPU every computing call returns -2.
#include "core.h"
#include "hostcopy.h"
#include <cstring>
#include <cstdlib>
#include <new>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#define LIBSAIS_OPENMP 1      /* export the *_omp symbols unconditionally */
#include "../../include/libsais.h"
#include "../../include/libsais64.h"
#include "../../include/libsais_cuda.h"
#include "../../include/libsais16.h"
#include "../../include/libsais16x64.h"

using namespace lsc;

namespace {

Ctx *new_ctx(int device)
{
    if (device < 0) {
        const char *env = getenv("LIBSAIS_CUDA_DEVICE");
        if (env && *env) device = atoi(env);
        else if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    int prev = -1; cudaGetDevice(&prev);
    Ctx *c = new (std::nothrow) Ctx();
    if (!c) return nullptr;
    bool ok = c->init(device);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (!ok) { c->destroy(); delete c; return nullptr; }
    return c;
}

struct DefaultCtx {
    Ctx *c = nullptr;
    // At process exit the CUDA runtime may already be unloading when thread-local destructors run: touching it
    // then can fail or hang, so the context is simply leaked unless the runtime still answers.
    ~DefaultCtx() {
        if (!c) return;
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) { c->destroy(); delete c; }
        c = nullptr;
    }
};
thread_local DefaultCtx tl_default;

Ctx *default_ctx()
{
    if (!tl_default.c) tl_default.c = new_ctx(-1);
    return tl_default.c;
}

Ctx *as_ctx(const void *p) { return const_cast<Ctx *>(static_cast<const Ctx *>(p)); }

// RAII bracket of one API call on a context: device guard, fresh stats/arena, device timing.
struct Call {
    Ctx &c; DeviceGuard g; cudaEvent_t e0 = nullptr, e1 = nullptr; bool timed = false;
    explicit Call(Ctx &ctx) : c(ctx), g(ctx.device) { c.reset_stats(); c.reset_arena(); }
    void start_timer() {
        if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) { cudaEventRecord(e0, c.stream); timed = true; }
    }
    void stop_timer() { if (timed) cudaEventRecord(e1, c.stream); }
    // sync, fold timings; returns false on any CUDA failure
    bool finish() {
        bool ok = c.sync() && !c.failed();
        if (timed && ok) { float t = 0; if (cudaEventElapsedTime(&t, e0, e1) == cudaSuccess) c.last_device_ms = t; }
        if (ok && c.profiling) c.resolve_profile();
        return ok;
    }
    ~Call() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (c.copy_stream) cudaStreamSynchronize(c.copy_stream);       // nothing of this call may still be writing the caller's buffers
        if (c.failed()) { cudaStreamSynchronize(c.stream); cudaGetLastError(); }
    }
};

const size_t kPad = 64;   // zero bytes after a device text: PLCP's 8-byte windows may over-read

// Copy a host text to a padded, zero-terminated arena buffer.
void *upload_text(Ctx &c, const void *h, size_t bytes)
{
    char *d = (char *)c.alloc(bytes + kPad);
    if (!d) return nullptr;
    c.check(cudaMemsetAsync(d + bytes, 0, kPad, c.stream));
    if (!copy_h2d(c, d, h, bytes)) return nullptr;
    return d;
}

template <typename IDX> void store_freq(Ctx &c, IDX *freq)
{
    if (freq) for (int s = 0; s < 256; ++s) freq[s] = (IDX)c.h_scalars[S_FREQ + s];
}

template <typename IDX> void host_freq(const uint8_t *T, IDX n, IDX *freq)
{
    if (!freq) return;
    for (int s = 0; s < 256; ++s) freq[s] = 0;
    for (IDX i = 0; i < n; ++i) freq[T[i]]++;
}

// Device SA (u32) -> host SA of the API's index width.
template <typename IDX> bool download_indexes(Ctx &c, const u32 *d_src, IDX *h_dst, u64 count, void *scratch8)
{
    if (sizeof(IDX) == 4) {
        return copy_d2h(c, h_dst, d_src, count * 4);
    }
    i64 *wide = (i64 *)scratch8;
    run_widen(c, d_src, wide, count);
    return copy_d2h(c, h_dst, wide, count * 8);
}

// Host index array (API width) -> device u32 array.
template <typename IDX> u32 *upload_indexes(Ctx &c, const IDX *h_src, u64 count)
{
    u32 *d = c.alloc_n<u32>(count);
    if (!d) return nullptr;
    if (sizeof(IDX) == 4) {
        if (!copy_h2d(c, d, h_src, count * 4)) return nullptr;
    } else {
        i64 *wide = c.alloc_n<i64>(count);
        if (!wide) return nullptr;
        if (!copy_h2d(c, wide, h_src, count * 8)) return nullptr;
        run_narrow(c, wide, d, count);
    }
    return d;
}

// ------------------------------------------------------------------------------------------
// generic bodies, IDX = int32_t (libsais_*) or int64_t (libsais64_*)
// ------------------------------------------------------------------------------------------
// libsais64 beyond one GPU: all visible GPUs (or $LIBSAIS_CUDA_DIST of them) run the distributed prefix doubling of dist64.cu
int multi_gpu_count(u64 n)
{
    int want = 0;
    const char *env = getenv("LIBSAIS_CUDA_DIST");
    if (env && *env) want = atoi(env);
    if (want <= 0 && n <= kMaxN) return 0;
    int have = libsais_cuda_device_count();
    if (have <= 0) return -1;
    if (want <= 0) want = have;
    return want;
}

template <typename IDX>
IDX sa_body(Ctx *c, const uint8_t *T, IDX *SA, IDX n, IDX fs, IDX *freq)
{
    if (T == nullptr || SA == nullptr || n < 0 || fs < 0) return -1;
    if (n < 2) { host_freq(T, n, freq); if (n == 1) SA[0] = 0; return 0; }
    if (sizeof(IDX) == 8) {
        const int G = multi_gpu_count((u64)n);
        if (G < 0) return -2;
        if (G > 0) {
            int have = libsais_cuda_device_count();
            std::vector<int> devs;
            for (int i = 0; i < G; ++i) devs.push_back(i % have);          // more ranks than GPUs (testing): ranks share devices
            return (IDX)sa64_multi(T, (i64 *)SA, (u64)n, (i64 *)freq, devs.data(), G, nullptr);
        }
    }
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    Call call(*c);
    if (!c->reserve((size_t)n + kPad + sa_workspace_bytes((u64)n, 1) + 4096)) return -2;
    const u8 *d_T = (const u8 *)upload_text(*c, T, (size_t)n);
    if (!d_T) return -2;
    call.start_timer();
    SAResult res; SAOptions opt;
    if (build_sa(*c, d_T, 1, (u64)n, opt, &res) != 0) return -2;
    call.stop_timer();
    if (!download_indexes<IDX>(*c, res.SA, SA, (u64)n, res.scratch)) return -2;
    if (!call.finish()) return -2;
    store_freq(*c, freq);
    return 0;
}

template <typename IDX, typename SYM>
IDX sa_int_body(Ctx *c, SYM *T, IDX *SA, IDX n, IDX k, IDX fs)
{
    (void)k;                       // the symbol width is measured on the device; k is not trusted
    if (T == nullptr || SA == nullptr || n < 0 || fs < 0) return -1;
    if (n < 2) { if (n == 1) SA[0] = 0; return 0; }
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    Call call(*c);
    const size_t tb = (size_t)n * sizeof(SYM);
    if (!c->reserve(tb + kPad + sa_workspace_bytes((u64)n, (int)sizeof(SYM)) + 4096)) return -2;
    const void *d_T = upload_text(*c, T, tb);
    if (!d_T) return -2;
    call.start_timer();
    SAResult res; SAOptions opt;
    if (build_sa(*c, d_T, (int)sizeof(SYM), (u64)n, opt, &res) != 0) return -2;
    call.stop_timer();
    if (!download_indexes<IDX>(*c, res.SA, SA, (u64)n, res.scratch)) return -2;
    if (!call.finish()) return -2;
    return 0;
}

// Generalized SA (reference libsais_gsa, src/libsais.c:7033-7048): SA of the separator-ranked integer text.
template <typename IDX>
IDX gsa_body(Ctx *c, const uint8_t *T, IDX *SA, IDX n, IDX fs, IDX *freq)
{
    if (T == nullptr || SA == nullptr || n < 0 || (n > 0 && T[n - 1] != 0) || fs < 0) return -1;
    if (n <= 1) { host_freq(T, n, freq); if (n == 1) SA[0] = 0; return 0; }
    if (!c || !c->ok || (u64)n > kMaxN - 512) return -2;
    Call call(*c);
    if (!c->reserve((size_t)n + kPad + gsa_workspace_bytes((u64)n) + sa_workspace_bytes((u64)n, 4) + 8192)) return -2;
    const u8 *d_T = (const u8 *)upload_text(*c, T, (size_t)n);
    if (!d_T) return -2;
    call.start_timer();
    if (freq) run_byte_histogram(*c, d_T, (u64)n);
    if (freq) c->check(cudaMemcpyAsync(c->h_scalars + S_FREQ, c->d_scalars + S_FREQ, 256 * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    const u32 *d_Tint = build_gsa_text(*c, d_T, (u64)n);
    if (!d_Tint) return -2;
    // empty members (T[0] == 0 or "00") are rejected like the reference does (src/libsais.c:6886-6889)
    c->check(cudaMemcpyAsync(c->h_scalars + S_GSA_INVALID, c->d_scalars + S_GSA_INVALID, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    if (!c->sync()) return -2;
    if (c->h_scalars[S_GSA_INVALID] != 0) return -1;
    SAResult res; SAOptions opt;
    if (build_sa(*c, d_Tint, 4, (u64)n, opt, &res) != 0) return -2;
    call.stop_timer();
    if (!download_indexes<IDX>(*c, res.SA, SA, (u64)n, res.scratch)) return -2;
    if (!call.finish()) return -2;
    store_freq(*c, freq);
    return 0;
}

// PLCP of a generalized SA (reference libsais_plcp_gsa, :8381-8397): matches stop at the separators.
template <typename IDX>
IDX plcp_gsa_body(Ctx *c, const uint8_t *T, const IDX *SA, IDX *PLCP, IDX n)
{
    if (T == nullptr || SA == nullptr || PLCP == nullptr || n < 0 || (n > 0 && T[n - 1] != 0)) return -1;
    if (n <= 1) { if (n == 1) PLCP[0] = 0; return 0; }
    if (!c || !c->ok || (u64)n > kMaxN - 512) return -2;
    Call call(*c);
    if (!c->reserve((size_t)n + kPad + gsa_workspace_bytes((u64)n) + (size_t)n * (4 + 4 + 8 + 8) + plcp_workspace_bytes((u64)n) + 8192)) return -2;
    const u8 *d_T = (const u8 *)upload_text(*c, T, (size_t)n);
    u32 *d_SA = upload_indexes<IDX>(*c, SA, (u64)n);
    u32 *d_P = c->alloc_n<u32>((size_t)n);
    void *wide = sizeof(IDX) == 8 ? c->alloc((size_t)n * 8) : nullptr;
    if (!d_T || !d_SA || !d_P || (sizeof(IDX) == 8 && !wide)) return -2;
    call.start_timer();
    const u32 *d_Tint = build_gsa_text(*c, d_T, (u64)n);
    if (!d_Tint) return -2;
    if (run_plcp(*c, d_Tint, 4, d_SA, d_P, (u64)n) != 0) return -2;
    call.stop_timer();
    if (!download_indexes<IDX>(*c, d_P, PLCP, (u64)n, wide)) return -2;
    if (!call.finish()) return -2;
    return 0;
}

// BWT with optional aux sampling (r == 0: none).  Returns primary index (r == 0) or 0.
template <typename IDX>
IDX bwt_body(Ctx *c, const uint8_t *T, uint8_t *U, IDX *A, IDX n, IDX fs, IDX *freq, IDX r, IDX *I, bool aux)
{
    if (T == nullptr || U == nullptr || A == nullptr || n < 0 || fs < 0) return -1;
    if (aux && (r < 2 || (r & (r - 1)) != 0 || I == nullptr)) return -1;
    if (n <= 1) {
        host_freq(T, n, freq);
        if (n == 1) U[0] = T[0];
        if (aux) { I[0] = n; return 0; }
        return n;
    }
    if (sizeof(IDX) == 8) {
        const int G = multi_gpu_count((u64)n);
        if (G < 0) return -2;
        if (G > 0) {
            int have = libsais_cuda_device_count();
            std::vector<int> devs;
            for (int i = 0; i < G; ++i) devs.push_back(i % have);
            const i64 rc = bwt64_multi(T, U, (i64 *)A, (u64)n, (i64 *)freq, aux ? (u64)r : 0, aux ? (i64 *)I : nullptr, devs.data(), G);
            return rc < 0 ? (IDX)rc : (aux ? (IDX)0 : (IDX)rc);
        }
    }
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    Call call(*c);
    const u64 n_aux = aux ? ((u64)n - 1) / (u64)r + 1 : 0;
    if (!c->reserve((size_t)n * 3 + kPad + n_aux * 4 + sa_workspace_bytes((u64)n, 1) + 8192)) return -2;
    const u8 *d_T = (const u8 *)upload_text(*c, T, (size_t)n);
    u8 *d_U = c->alloc_n<u8>((size_t)n);
    u8 *d_rows = c->alloc_n<u8>((size_t)n);
    u32 *d_I = n_aux ? c->alloc_n<u32>(n_aux) : nullptr;
    if (!d_T || !d_U || !d_rows || (n_aux && !d_I)) return -2;
    call.start_timer();
    SAResult res; SAOptions opt;
    opt.want_sa = false; opt.bwt_rows = d_rows; opt.aux_r = aux ? (u64)r : 0; opt.aux_I = d_I;
    // A pinned output buffer receives the rows of settled slots while round 0 still sorts (boundary
    // src/libsais.c:7097-7121): the device alias of U is needed for the few rows that are sent again at the end.
    u8 *U_dev_alias = nullptr;
    const uint8_t last_symbol = T[n - 1];                 // U may be T (include/libsais.h): read before any row lands in U
    if (host_is_pinned(U) && cudaHostGetDevicePointer((void **)&U_dev_alias, (void *)U, 0) == cudaSuccess && U_dev_alias) { opt.h_U = U; opt.h_T = T; }
    else cudaGetLastError();
    if (build_sa(*c, d_T, 1, (u64)n, opt, &res) != 0) return -2;
    if (res.primary < 1 || res.primary > (u64)n) return -2;
    const u64 p0 = res.primary - 1;
    if (res.u_streamed && p0 >= res.p0_lo && p0 < res.p0_hi) {
        // everything outside the bucket of suffix 0 is on its way; wait for those copies, then the bucket itself (two
        // pieces around the dropped row) and the rows that were settled after round 0
        c->check(cudaEventRecord(c->chunk_ev[Ctx::kChunkEvents - 1], c->copy_stream));
        c->check(cudaStreamWaitEvent(c->stream, c->chunk_ev[Ctx::kChunkEvents - 1], 0));
        if (p0 > res.p0_lo) c->check(cudaMemcpyAsync(U + res.p0_lo + 1, d_rows + res.p0_lo, p0 - res.p0_lo, cudaMemcpyDeviceToHost, c->stream));
        if (p0 + 1 < res.p0_hi) c->check(cudaMemcpyAsync(U + p0 + 1, d_rows + p0 + 1, res.p0_hi - p0 - 1, cudaMemcpyDeviceToHost, c->stream));
        if (run_bwt_patch(*c, res.patch_slots, res.n_patch, d_rows, U_dev_alias, p0) != 0) return -2;
        call.stop_timer();
        U[0] = last_symbol;
    } else {
        if (opt.h_U && c->copy_stream) c->check(cudaStreamSynchronize(c->copy_stream));       // chunks may be in flight into U: let them land first
        if (run_bwt_finish(*c, d_T, d_rows, d_U, (u64)n, res.primary) != 0) return -2;
        call.stop_timer();
        if (!copy_d2h(*c, U, d_U, (size_t)n)) return -2;
    }
    if (n_aux && !download_indexes<IDX>(*c, d_I, I, n_aux, res.scratch)) return -2;
    if (!call.finish()) return -2;
    store_freq(*c, freq);
    return aux ? 0 : (IDX)res.primary;
}

template <typename IDX>
IDX unbwt_body(Ctx *c, const uint8_t *T, uint8_t *U, IDX *A, IDX n, const IDX *freq, IDX r, const IDX *I)
{
    (void)freq;                    // recomputed on the device (the reference merely trusts it)
    if (T == nullptr || U == nullptr || A == nullptr || n < 0 || I == nullptr) return -1;
    if (r != n && (r < 2 || (r & (r - 1)) != 0)) return -1;
    if (n <= 1) {
        if (I[0] != n) return -1;
        if (n == 1) U[0] = T[0];
        return 0;
    }
    for (IDX t = 0; t <= (n - 1) / r; ++t) if (I[t] <= 0 || I[t] > n) return -1;
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    Call call(*c);
    const u64 n_aux = r < n ? ((u64)n - 1) / (u64)r + 1 : 0;        // r == n: only the primary index is known
    if (!c->reserve((size_t)n * 2 + kPad + unbwt_workspace_bytes((u64)n) + n_aux * 12 + 8192)) return -2;
    const u8 *d_B = (const u8 *)upload_text(*c, T, (size_t)n);
    u8 *d_U = c->alloc_n<u8>((size_t)n);
    if (!d_B || !d_U) return -2;
    const u32 *d_I = nullptr;
    if (n_aux >= 2) { d_I = upload_indexes<IDX>(*c, I, n_aux); if (!d_I) return -2; }
    call.start_timer();
    if (run_unbwt(*c, d_B, d_U, (u64)n, (u64)I[0], n_aux >= 2 ? (u64)r : 0, d_I, n_aux) != 0) return -2;
    call.stop_timer();
    if (!copy_d2h(*c, U, d_U, (size_t)n)) return -2;
    if (!call.finish()) return -2;
    return 0;
}

template <typename IDX, typename SYM>
IDX plcp_body(Ctx *c, const SYM *T, const IDX *SA, IDX *PLCP, IDX n)
{
    if (T == nullptr || SA == nullptr || PLCP == nullptr || n < 0) return -1;
    if (n <= 1) { if (n == 1) PLCP[0] = 0; return 0; }
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    Call call(*c);
    const size_t tb = (size_t)n * sizeof(SYM);
    if (!c->reserve(tb + kPad + (size_t)n * (4 + 4 + 8 + 8) + plcp_workspace_bytes((u64)n) + 8192)) return -2;
    const void *d_T = upload_text(*c, T, tb);
    u32 *d_SA = upload_indexes<IDX>(*c, SA, (u64)n);
    u32 *d_P = c->alloc_n<u32>((size_t)n);
    void *wide = sizeof(IDX) == 8 ? c->alloc((size_t)n * 8) : nullptr;
    if (!d_T || !d_SA || !d_P || (sizeof(IDX) == 8 && !wide)) return -2;
    call.start_timer();
    if (run_plcp(*c, d_T, (int)sizeof(SYM), d_SA, d_P, (u64)n) != 0) return -2;
    call.stop_timer();
    if (!download_indexes<IDX>(*c, d_P, PLCP, (u64)n, wide)) return -2;
    if (!call.finish()) return -2;
    return 0;
}

template <typename IDX>
IDX lcp_body(Ctx *c, const IDX *PLCP, const IDX *SA, IDX *LCP, IDX n)
{
    if (PLCP == nullptr || SA == nullptr || LCP == nullptr || n < 0) return -1;
    if (n <= 1) { if (n == 1) LCP[0] = PLCP[SA[0]]; return 0; }
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    Call call(*c);
    if (!c->reserve((size_t)n * (4 + 4 + 4 + 8 + 8 + 8) + 8192)) return -2;
    u32 *d_P = upload_indexes<IDX>(*c, PLCP, (u64)n);
    u32 *d_SA = upload_indexes<IDX>(*c, SA, (u64)n);
    u32 *d_L = c->alloc_n<u32>((size_t)n);
    void *wide = sizeof(IDX) == 8 ? c->alloc((size_t)n * 8) : nullptr;
    if (!d_P || !d_SA || !d_L || (sizeof(IDX) == 8 && !wide)) return -2;
    call.start_timer();
    if (run_lcp(*c, d_P, d_SA, d_L, (u64)n) != 0) return -2;
    call.stop_timer();
    if (!download_indexes<IDX>(*c, d_L, LCP, (u64)n, wide)) return -2;   // LCP may alias SA: SA was consumed above
    if (!call.finish()) return -2;
    return 0;
}

// ------------------------------------------------------------------------------------------
// 16-bit symbols (reference include/libsais16.h, include/libsais16x64.h; src/libsais16.c:6995 ff.): the SA and PLCP cores run on
// the text widened to 32-bit symbols on the device; BWT rows are gathered from the suffix array; the inverse BWT handles
// the 65536-symbol alphabet (post.cu).  Validation and fast paths mirror the 8-bit bodies above.
// ------------------------------------------------------------------------------------------
template <typename IDX> void host_freq16(const uint16_t *T, IDX n, IDX *freq)
{
    if (!freq) return;
    for (int s = 0; s < 65536; ++s) freq[s] = 0;
    for (IDX i = 0; i < n; ++i) freq[T[i]]++;
}
template <typename IDX> bool download_freq16(Ctx &c, const u64 *d_hist, IDX *freq)
{
    std::vector<u64> h(65536);
    if (!copy_d2h(c, h.data(), d_hist, 65536 * sizeof(u64)) || !c.sync()) return false;
    for (int s = 0; s < 65536; ++s) freq[s] = (IDX)h[s];
    return true;
}

// device text of a uint16_t host text: the raw symbols and the widened (or, for GSA, separator-ranked) 32-bit text
struct Text16 { const uint16_t *raw = nullptr; const u32 *wide = nullptr; };
bool upload_text16(Ctx &c, const uint16_t *T, u64 n, bool gsa, Text16 *out)
{
    const uint16_t *d16 = (const uint16_t *)upload_text(c, T, (size_t)n * 2);
    if (!d16) return false;
    out->raw = d16;
    if (gsa) {
        out->wide = build_gsa_text16(c, d16, n);
        return out->wide != nullptr;
    }
    u32 *w = (u32 *)c.alloc((size_t)n * 4 + kPad);
    if (!w) return false;
    c.check(cudaMemsetAsync((char *)w + (size_t)n * 4, 0, kPad, c.stream));
    run_widen16(c, d16, w, n);
    out->wide = w;
    return true;
}

template <typename IDX>
IDX sa16_body(Ctx *c, const uint16_t *T, IDX *SA, IDX n, IDX fs, IDX *freq, bool gsa)
{
    if (T == nullptr || SA == nullptr || n < 0 || fs < 0 || (gsa && n > 0 && T[n - 1] != 0)) return -1;
    if (n < 2) { host_freq16(T, n, freq); if (n == 1) SA[0] = 0; return 0; }
    if (!c || !c->ok || (u64)n > kMaxN - 512) return -2;
    Call call(*c);
    if (!c->reserve((size_t)n * 2 + kPad * 2 + (size_t)n * 4 + 65536 * 8 + gsa_workspace_bytes((u64)n) + sa_workspace_bytes((u64)n, 4) + 16384)) return -2;
    Text16 t;
    u64 *d_hist = freq ? c->alloc_n<u64>(65536) : nullptr;
    if (freq && !d_hist) return -2;
    if (!upload_text16(*c, T, (u64)n, gsa, &t)) return -2;
    call.start_timer();
    if (freq) run_hist_u16(*c, t.raw, (u64)n, d_hist);
    if (gsa) {      // empty members are rejected like the reference does
        c->check(cudaMemcpyAsync(c->h_scalars + S_GSA_INVALID, c->d_scalars + S_GSA_INVALID, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        if (!c->sync()) return -2;
        if (c->h_scalars[S_GSA_INVALID] != 0) return -1;
    }
    SAResult res; SAOptions opt;
    if (build_sa(*c, t.wide, 4, (u64)n, opt, &res) != 0) return -2;
    call.stop_timer();
    if (!download_indexes<IDX>(*c, res.SA, SA, (u64)n, res.scratch)) return -2;
    if (freq && !download_freq16(*c, d_hist, freq)) return -2;
    if (!call.finish()) return -2;
    return 0;
}

template <typename IDX>
IDX bwt16_body(Ctx *c, const uint16_t *T, uint16_t *U, IDX *A, IDX n, IDX fs, IDX *freq, IDX r, IDX *I, bool aux)
{
    if (T == nullptr || U == nullptr || A == nullptr || n < 0 || fs < 0) return -1;
    if (aux && (r < 2 || (r & (r - 1)) != 0 || I == nullptr)) return -1;
    if (n <= 1) {
        host_freq16(T, n, freq);
        if (n == 1) U[0] = T[0];
        if (aux) { I[0] = n; return 0; }
        return n;
    }
    if (!c || !c->ok || (u64)n > kMaxN - 512) return -2;
    Call call(*c);
    const u64 n_aux = aux ? ((u64)n - 1) / (u64)r + 1 : 0;
    if (!c->reserve((size_t)n * 4 + kPad * 2 + (size_t)n * 4 + 65536 * 8 + n_aux * 4 + sa_workspace_bytes((u64)n, 4) + 16384)) return -2;
    Text16 t;
    u64 *d_hist = freq ? c->alloc_n<u64>(65536) : nullptr;
    uint16_t *d_U = c->alloc_n<uint16_t>((size_t)n);
    u32 *d_I = n_aux ? c->alloc_n<u32>(n_aux) : nullptr;
    if ((freq && !d_hist) || !d_U || (n_aux && !d_I)) return -2;
    if (!upload_text16(*c, T, (u64)n, false, &t)) return -2;
    call.start_timer();
    if (freq) run_hist_u16(*c, t.raw, (u64)n, d_hist);
    SAResult res; SAOptions opt;
    if (build_sa(*c, t.wide, 4, (u64)n, opt, &res) != 0) return -2;
    u64 primary = 0;
    if (run_bwt16(*c, t.raw, res.SA, d_U, (u64)n, aux ? (u64)r : 0, d_I, &primary) != 0) return -2;
    call.stop_timer();
    if (!copy_d2h(*c, U, d_U, (size_t)n * 2)) return -2;
    if (n_aux && !download_indexes<IDX>(*c, d_I, I, n_aux, res.scratch)) return -2;
    if (freq && !download_freq16(*c, d_hist, freq)) return -2;
    if (!call.finish()) return -2;
    return aux ? 0 : (IDX)primary;
}

template <typename IDX>
IDX unbwt16_body(Ctx *c, const uint16_t *T, uint16_t *U, IDX *A, IDX n, const IDX *freq, IDX r, const IDX *I)
{
    (void)freq;
    if (T == nullptr || U == nullptr || A == nullptr || n < 0 || I == nullptr) return -1;
    if (r != n && (r < 2 || (r & (r - 1)) != 0)) return -1;
    if (n <= 1) {
        if (I[0] != n) return -1;
        if (n == 1) U[0] = T[0];
        return 0;
    }
    for (IDX t = 0; t <= (n - 1) / r; ++t) if (I[t] <= 0 || I[t] > n) return -1;
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    Call call(*c);
    const u64 n_aux = r < n ? ((u64)n - 1) / (u64)r + 1 : 0;
    if (!c->reserve((size_t)n * 4 + kPad + unbwt16_workspace_bytes((u64)n) + n_aux * 12 + 8192)) return -2;
    uint16_t *d_B = (uint16_t *)upload_text(*c, T, (size_t)n * 2);
    uint16_t *d_U = c->alloc_n<uint16_t>((size_t)n);
    if (!d_B || !d_U) return -2;
    const u32 *d_I = nullptr;
    if (n_aux >= 2) { d_I = upload_indexes<IDX>(*c, I, n_aux); if (!d_I) return -2; }
    call.start_timer();
    if (run_unbwt16(*c, d_B, d_U, (u64)n, (u64)I[0], n_aux >= 2 ? (u64)r : 0, d_I, n_aux) != 0) return -2;
    call.stop_timer();
    if (!copy_d2h(*c, U, d_U, (size_t)n * 2)) return -2;
    if (!call.finish()) return -2;
    return 0;
}

template <typename IDX>
IDX plcp16_body(Ctx *c, const uint16_t *T, const IDX *SA, IDX *PLCP, IDX n, bool gsa)
{
    if (T == nullptr || SA == nullptr || PLCP == nullptr || n < 0 || (gsa && n > 0 && T[n - 1] != 0)) return -1;
    if (n <= 1) { if (n == 1) PLCP[0] = 0; return 0; }
    if (!c || !c->ok || (u64)n > kMaxN - 512) return -2;
    Call call(*c);
    if (!c->reserve((size_t)n * 2 + kPad * 2 + gsa_workspace_bytes((u64)n) + (size_t)n * (4 + 4 + 4 + 8 + 8) + plcp_workspace_bytes((u64)n) + 16384)) return -2;
    Text16 t;
    if (!upload_text16(*c, T, (u64)n, gsa, &t)) return -2;
    u32 *d_SA = upload_indexes<IDX>(*c, SA, (u64)n);
    u32 *d_P = c->alloc_n<u32>((size_t)n);
    void *wide = sizeof(IDX) == 8 ? c->alloc((size_t)n * 8) : nullptr;
    if (!d_SA || !d_P || (sizeof(IDX) == 8 && !wide)) return -2;
    call.start_timer();
    if (run_plcp(*c, t.wide, 4, d_SA, d_P, (u64)n) != 0) return -2;
    call.stop_timer();
    if (!download_indexes<IDX>(*c, d_P, PLCP, (u64)n, wide)) return -2;
    if (!call.finish()) return -2;
    return 0;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ contexts
void *libsais_create_ctx(void) { return new_ctx(-1); }
void *libsais_create_ctx_omp(int32_t threads) { return threads < 0 ? nullptr : new_ctx(-1); }
void libsais_free_ctx(void *ctx) { if (ctx) { Ctx *c = as_ctx(ctx); c->destroy(); delete c; } }
void *libsais_unbwt_create_ctx(void) { return new_ctx(-1); }
void *libsais_unbwt_create_ctx_omp(int32_t threads) { return threads < 0 ? nullptr : new_ctx(-1); }
void libsais_unbwt_free_ctx(void *ctx) { libsais_free_ctx(ctx); }

// ------------------------------------------------------------------ libsais (int32)
int32_t libsais(const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq)
{ return sa_body<int32_t>(default_ctx(), T, SA, n, fs, freq); }
int32_t libsais_ctx(const void *ctx, const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq)
{ if (ctx == nullptr) return -1; return sa_body<int32_t>(as_ctx(ctx), T, SA, n, fs, freq); }
int32_t libsais_omp(const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq, int32_t threads)
{ if (threads < 0) return -1; return sa_body<int32_t>(default_ctx(), T, SA, n, fs, freq); }

int32_t libsais_int(int32_t *T, int32_t *SA, int32_t n, int32_t k, int32_t fs)
{ return sa_int_body<int32_t, int32_t>(default_ctx(), T, SA, n, k, fs); }
int32_t libsais_int_omp(int32_t *T, int32_t *SA, int32_t n, int32_t k, int32_t fs, int32_t threads)
{ if (threads < 0) return -1; return sa_int_body<int32_t, int32_t>(default_ctx(), T, SA, n, k, fs); }

// generalized suffix arrays
int32_t libsais_gsa(const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq)
{ return gsa_body<int32_t>(default_ctx(), T, SA, n, fs, freq); }
int32_t libsais_gsa_ctx(const void *ctx, const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq)
{ if (ctx == nullptr) return -1; return gsa_body<int32_t>(as_ctx(ctx), T, SA, n, fs, freq); }
int32_t libsais_gsa_omp(const uint8_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq, int32_t threads)
{ if (threads < 0) return -1; return gsa_body<int32_t>(default_ctx(), T, SA, n, fs, freq); }
int32_t libsais_plcp_gsa(const uint8_t *T, const int32_t *SA, int32_t *PLCP, int32_t n)
{ return plcp_gsa_body<int32_t>(default_ctx(), T, SA, PLCP, n); }
int32_t libsais_plcp_gsa_omp(const uint8_t *T, const int32_t *SA, int32_t *PLCP, int32_t n, int32_t threads)
{ if (threads < 0) return -1; return plcp_gsa_body<int32_t>(default_ctx(), T, SA, PLCP, n); }

int32_t libsais_bwt(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, int32_t fs, int32_t *freq)
{ return bwt_body<int32_t>(default_ctx(), T, U, A, n, fs, freq, 0, nullptr, false); }
int32_t libsais_bwt_aux(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, int32_t fs, int32_t *freq, int32_t r, int32_t *I)
{ return bwt_body<int32_t>(default_ctx(), T, U, A, n, fs, freq, r, I, true); }
int32_t libsais_bwt_ctx(const void *ctx, const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, int32_t fs, int32_t *freq)
{ if (ctx == nullptr) return -1; return bwt_body<int32_t>(as_ctx(ctx), T, U, A, n, fs, freq, 0, nullptr, false); }
int32_t libsais_bwt_aux_ctx(const void *ctx, const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, int32_t fs, int32_t *freq, int32_t r, int32_t *I)
{ if (ctx == nullptr) return -1; return bwt_body<int32_t>(as_ctx(ctx), T, U, A, n, fs, freq, r, I, true); }
int32_t libsais_bwt_omp(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, int32_t fs, int32_t *freq, int32_t threads)
{ if (threads < 0) return -1; return bwt_body<int32_t>(default_ctx(), T, U, A, n, fs, freq, 0, nullptr, false); }
int32_t libsais_bwt_aux_omp(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, int32_t fs, int32_t *freq, int32_t r, int32_t *I, int32_t threads)
{ if (threads < 0) return -1; return bwt_body<int32_t>(default_ctx(), T, U, A, n, fs, freq, r, I, true); }

int32_t libsais_unbwt(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, const int32_t *freq, int32_t i)
{ return unbwt_body<int32_t>(default_ctx(), T, U, A, n, freq, n, &i); }
int32_t libsais_unbwt_ctx(const void *ctx, const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, const int32_t *freq, int32_t i)
{ if (ctx == nullptr) return -1; return unbwt_body<int32_t>(as_ctx(ctx), T, U, A, n, freq, n, &i); }
int32_t libsais_unbwt_aux(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, const int32_t *freq, int32_t r, const int32_t *I)
{ return unbwt_body<int32_t>(default_ctx(), T, U, A, n, freq, r, I); }
int32_t libsais_unbwt_aux_ctx(const void *ctx, const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, const int32_t *freq, int32_t r, const int32_t *I)
{ if (ctx == nullptr) return -1; return unbwt_body<int32_t>(as_ctx(ctx), T, U, A, n, freq, r, I); }
int32_t libsais_unbwt_omp(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, const int32_t *freq, int32_t i, int32_t threads)
{ if (threads < 0) return -1; return unbwt_body<int32_t>(default_ctx(), T, U, A, n, freq, n, &i); }
int32_t libsais_unbwt_aux_omp(const uint8_t *T, uint8_t *U, int32_t *A, int32_t n, const int32_t *freq, int32_t r, const int32_t *I, int32_t threads)
{ if (threads < 0) return -1; return unbwt_body<int32_t>(default_ctx(), T, U, A, n, freq, r, I); }

int32_t libsais_plcp(const uint8_t *T, const int32_t *SA, int32_t *PLCP, int32_t n)
{ return plcp_body<int32_t, uint8_t>(default_ctx(), T, SA, PLCP, n); }
int32_t libsais_plcp_int(const int32_t *T, const int32_t *SA, int32_t *PLCP, int32_t n)
{ return plcp_body<int32_t, int32_t>(default_ctx(), T, SA, PLCP, n); }
int32_t libsais_lcp(const int32_t *PLCP, const int32_t *SA, int32_t *LCP, int32_t n)
{ return lcp_body<int32_t>(default_ctx(), PLCP, SA, LCP, n); }
int32_t libsais_plcp_omp(const uint8_t *T, const int32_t *SA, int32_t *PLCP, int32_t n, int32_t threads)
{ if (threads < 0) return -1; return plcp_body<int32_t, uint8_t>(default_ctx(), T, SA, PLCP, n); }
int32_t libsais_plcp_int_omp(const int32_t *T, const int32_t *SA, int32_t *PLCP, int32_t n, int32_t threads)
{ if (threads < 0) return -1; return plcp_body<int32_t, int32_t>(default_ctx(), T, SA, PLCP, n); }
int32_t libsais_lcp_omp(const int32_t *PLCP, const int32_t *SA, int32_t *LCP, int32_t n, int32_t threads)
{ if (threads < 0) return -1; return lcp_body<int32_t>(default_ctx(), PLCP, SA, LCP, n); }

// ------------------------------------------------------------------ libsais64 (int64)
int64_t libsais64(const uint8_t *T, int64_t *SA, int64_t n, int64_t fs, int64_t *freq)
{ return sa_body<int64_t>(default_ctx(), T, SA, n, fs, freq); }
int64_t libsais64_omp(const uint8_t *T, int64_t *SA, int64_t n, int64_t fs, int64_t *freq, int64_t threads)
{ if (threads < 0) return -1; return sa_body<int64_t>(default_ctx(), T, SA, n, fs, freq); }
int64_t libsais64_long(int64_t *T, int64_t *SA, int64_t n, int64_t k, int64_t fs)
{ return sa_int_body<int64_t, int64_t>(default_ctx(), T, SA, n, k, fs); }
int64_t libsais64_long_omp(int64_t *T, int64_t *SA, int64_t n, int64_t k, int64_t fs, int64_t threads)
{ if (threads < 0) return -1; return sa_int_body<int64_t, int64_t>(default_ctx(), T, SA, n, k, fs); }

int64_t libsais64_gsa(const uint8_t *T, int64_t *SA, int64_t n, int64_t fs, int64_t *freq)
{ return gsa_body<int64_t>(default_ctx(), T, SA, n, fs, freq); }
int64_t libsais64_gsa_omp(const uint8_t *T, int64_t *SA, int64_t n, int64_t fs, int64_t *freq, int64_t threads)
{ if (threads < 0) return -1; return gsa_body<int64_t>(default_ctx(), T, SA, n, fs, freq); }
int64_t libsais64_plcp_gsa(const uint8_t *T, const int64_t *SA, int64_t *PLCP, int64_t n)
{ return plcp_gsa_body<int64_t>(default_ctx(), T, SA, PLCP, n); }
int64_t libsais64_plcp_gsa_omp(const uint8_t *T, const int64_t *SA, int64_t *PLCP, int64_t n, int64_t threads)
{ if (threads < 0) return -1; return plcp_gsa_body<int64_t>(default_ctx(), T, SA, PLCP, n); }

int64_t libsais64_bwt(const uint8_t *T, uint8_t *U, int64_t *A, int64_t n, int64_t fs, int64_t *freq)
{ return bwt_body<int64_t>(default_ctx(), T, U, A, n, fs, freq, 0, nullptr, false); }
int64_t libsais64_bwt_aux(const uint8_t *T, uint8_t *U, int64_t *A, int64_t n, int64_t fs, int64_t *freq, int64_t r, int64_t *I)
{ return bwt_body<int64_t>(default_ctx(), T, U, A, n, fs, freq, r, I, true); }
int64_t libsais64_bwt_omp(const uint8_t *T, uint8_t *U, int64_t *A, int64_t n, int64_t fs, int64_t *freq, int64_t threads)
{ if (threads < 0) return -1; return bwt_body<int64_t>(default_ctx(), T, U, A, n, fs, freq, 0, nullptr, false); }
int64_t libsais64_bwt_aux_omp(const uint8_t *T, uint8_t *U, int64_t *A, int64_t n, int64_t fs, int64_t *freq, int64_t r, int64_t *I, int64_t threads)
{ if (threads < 0) return -1; return bwt_body<int64_t>(default_ctx(), T, U, A, n, fs, freq, r, I, true); }

int64_t libsais64_unbwt(const uint8_t *T, uint8_t *U, int64_t *A, int64_t n, const int64_t *freq, int64_t i)
{ return unbwt_body<int64_t>(default_ctx(), T, U, A, n, freq, n, &i); }
int64_t libsais64_unbwt_aux(const uint8_t *T, uint8_t *U, int64_t *A, int64_t n, const int64_t *freq, int64_t r, const int64_t *I)
{ return unbwt_body<int64_t>(default_ctx(), T, U, A, n, freq, r, I); }
int64_t libsais64_unbwt_omp(const uint8_t *T, uint8_t *U, int64_t *A, int64_t n, const int64_t *freq, int64_t i, int64_t threads)
{ if (threads < 0) return -1; return unbwt_body<int64_t>(default_ctx(), T, U, A, n, freq, n, &i); }
int64_t libsais64_unbwt_aux_omp(const uint8_t *T, uint8_t *U, int64_t *A, int64_t n, const int64_t *freq, int64_t r, const int64_t *I, int64_t threads)
{ if (threads < 0) return -1; return unbwt_body<int64_t>(default_ctx(), T, U, A, n, freq, r, I); }

int64_t libsais64_plcp(const uint8_t *T, const int64_t *SA, int64_t *PLCP, int64_t n)
{ return plcp_body<int64_t, uint8_t>(default_ctx(), T, SA, PLCP, n); }
int64_t libsais64_lcp(const int64_t *PLCP, const int64_t *SA, int64_t *LCP, int64_t n)
{ return lcp_body<int64_t>(default_ctx(), PLCP, SA, LCP, n); }
int64_t libsais64_plcp_omp(const uint8_t *T, const int64_t *SA, int64_t *PLCP, int64_t n, int64_t threads)
{ if (threads < 0) return -1; return plcp_body<int64_t, uint8_t>(default_ctx(), T, SA, PLCP, n); }
int64_t libsais64_lcp_omp(const int64_t *PLCP, const int64_t *SA, int64_t *LCP, int64_t n, int64_t threads)
{ if (threads < 0) return -1; return lcp_body<int64_t>(default_ctx(), PLCP, SA, LCP, n); }

// ------------------------------------------------------------------ extras (libsais_cuda.h)
int32_t libsais_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
void *libsais_cuda_create_ctx(int32_t device) { return new_ctx(device); }
void *libsais_cuda_stream(const void *ctx) { Ctx *c = ctx ? as_ctx(ctx) : default_ctx(); return c ? (void *)c->stream : nullptr; }
int32_t libsais_cuda_set_profiling(const void *ctx, int32_t on)
{ Ctx *c = ctx ? as_ctx(ctx) : default_ctx(); if (!c) return -2; c->profiling = on != 0; return 0; }

int32_t libsais_cuda_get_stats(const void *ctx, libsais_cuda_stats *out)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !out) return -1;
    std::memset(out, 0, sizeof(*out));
    out->n_classes = KC_COUNT;
    out->n_rounds = (int32_t)c->rounds.size();
    out->total_launches = c->total_launches;
    for (int i = 0; i < KC_COUNT; ++i) { out->launches[i] = c->launches[i]; out->ms[i] = c->ms[i]; out->bytes[i] = c->bytes[i]; }
    out->device_ms = c->last_device_ms;
    out->workspace_bytes = c->ws_cap;
    return 0;
}
int32_t libsais_cuda_get_round(const void *ctx, int32_t round, libsais_cuda_round *out)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !out || round < 0 || (size_t)round >= c->rounds.size()) return -1;
    const RoundStat &r = c->rounds[round];
    out->h = r.h; out->n_active = r.n_active; out->n_groups = r.n_groups; out->passes = r.passes; out->key_bits = r.key_bits;
    out->device_ms = r.ms; out->bytes = r.bytes;
    return 0;
}
// Give the context's device workspace and pinned staging back to the system (it is re-grown on the next call).
int32_t libsais_cuda_release_workspace(const void *ctx)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c) return -1;
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->ws) { cudaFree(c->ws); c->ws = nullptr; c->ws_cap = 0; c->ws_off = 0; }
    if (c->dist_words) { cudaFree(c->dist_words); c->dist_words = nullptr; }
    c->free_staging();
    cudaGetLastError();
    return 0;
}
const char *libsais_cuda_kernel_class_name(int32_t kc) { return kc >= 0 && kc < KC_COUNT ? kKernelClassName[kc] : ""; }
// debug: raw device scalars S_ERR.. (look-back statistics when built with -DLSC_LOOKBACK_STATS)
int32_t libsais_cuda_debug_scalars(const void *ctx, uint32_t *out, int32_t count)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !out || count < 0 || count > 16) return -1;
    DeviceGuard g(c->device);
    // counters live at u32 offset 112 from S_ERR (= u64 slot S_ERR + 56); read and clear
    u64 *p = c->d_scalars + S_ERR + 56;
    if (cudaMemcpy(out, p, count * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    return cudaMemset(p, 0, 16 * sizeof(uint32_t)) == cudaSuccess ? 0 : -2;
}
int32_t libsais_cuda_last_error(const void *ctx) { Ctx *c = ctx ? as_ctx(ctx) : default_ctx(); return c ? (int32_t)c->last_error : -1; }

int64_t libsais_cuda_sa_dev(const void *ctx, const uint8_t *d_T, uint32_t *d_SA, int64_t n)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (d_T == nullptr || d_SA == nullptr || n < 0) return -1;
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    if (n == 0) return 0;
    Call call(*c);
    if (!c->reserve(sa_workspace_bytes((u64)n, 1) + 4096)) return -2;
    call.start_timer();
    SAResult res; SAOptions opt;
    opt.sa_out = d_SA;
    if (build_sa(*c, d_T, 1, (u64)n, opt, &res) != 0) return -2;
    call.stop_timer();
    return call.finish() ? 0 : -2;
}

int64_t libsais_cuda_bwt_dev(const void *ctx, const uint8_t *d_T, uint8_t *d_U, int64_t n)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (d_T == nullptr || d_U == nullptr || n < 0) return -1;
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    if (n == 0) return 0;
    Call call(*c);
    if (!c->reserve((size_t)n + sa_workspace_bytes((u64)n, 1) + 4096)) return -2;
    u8 *d_rows = c->alloc_n<u8>((size_t)n);
    if (!d_rows) return -2;
    call.start_timer();
    SAResult res; SAOptions opt;
    opt.want_sa = false; opt.bwt_rows = d_rows;
    if (build_sa(*c, d_T, 1, (u64)n, opt, &res) != 0) return -2;
    if (res.primary < 1 || res.primary > (u64)n) return -2;
    if (run_bwt_finish(*c, d_T, d_rows, d_U, (u64)n, res.primary) != 0) return -2;
    call.stop_timer();
    if (!call.finish()) return -2;
    return (int64_t)res.primary;
}

int64_t libsais_cuda_plcp_dev(const void *ctx, const uint8_t *d_T, const uint32_t *d_SA, uint32_t *d_PLCP, int64_t n)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (d_T == nullptr || d_SA == nullptr || d_PLCP == nullptr || n < 0) return -1;
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    if (n == 0) return 0;
    Call call(*c);
    if (!c->reserve((size_t)n + kPad + plcp_workspace_bytes((u64)n) + 4096)) return -2;
    // own padded copy of the text: the compare kernel's 8-byte windows read past the end
    u8 *d_Tp = c->alloc_n<u8>((size_t)n + kPad);
    if (!d_Tp) return -2;
    call.start_timer();
    c->check(cudaMemsetAsync(d_Tp + n, 0, kPad, c->stream));
    c->check(cudaMemcpyAsync(d_Tp, d_T, (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    if (run_plcp(*c, d_Tp, 1, d_SA, d_PLCP, (u64)n) != 0) return -2;
    call.stop_timer();
    return call.finish() ? 0 : -2;
}

int64_t libsais_cuda_lcp_dev(const void *ctx, const uint32_t *d_PLCP, const uint32_t *d_SA, uint32_t *d_LCP, int64_t n)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (d_PLCP == nullptr || d_SA == nullptr || d_LCP == nullptr || n < 0) return -1;
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    if (n == 0) return 0;
    Call call(*c);
    call.start_timer();
    if (run_lcp(*c, d_PLCP, d_SA, d_LCP, (u64)n) != 0) return -2;
    call.stop_timer();
    return call.finish() ? 0 : -2;
}

int64_t libsais_cuda_unbwt_dev(const void *ctx, const uint8_t *d_B, uint8_t *d_U, int64_t n, int64_t primary)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (d_B == nullptr || d_U == nullptr || n < 0) return -1;
    if (n == 0) return primary == 0 ? 0 : -1;
    if (primary < 1 || primary > n) return -1;
    if (!c || !c->ok || (u64)n > kMaxN) return -2;
    Call call(*c);
    if (!c->reserve(unbwt_workspace_bytes((u64)n) + 8192)) return -2;
    call.start_timer();
    if (run_unbwt(*c, d_B, d_U, (u64)n, (u64)primary) != 0) return -2;
    call.stop_timer();
    return call.finish() ? 0 : -2;
}


// ------------------------------------------------------------------ batch of independent blocks (BASELINE config 4)
// The reference's model for many independent texts is "one context per host thread"
// (include/libsais.h:53-55).  This is the same model packaged for GPUs: block b runs on
// devices[b % ndevices]; every device gets `lanes` host threads with one context (= stream +
// workspace) each, so the H2D copy of one block overlaps the kernels of another and the D2H of
// a third -- both PCIe directions and the SMs stay busy.  Contexts live in a process-wide pool.
namespace {
std::mutex g_batch_mutex;
std::map<int, std::vector<Ctx *>> g_batch_pool;

Ctx *batch_ctx(int device, int lane)
{
    std::vector<Ctx *> &v = g_batch_pool[device];
    while ((int)v.size() <= lane) v.push_back(nullptr);
    if (!v[lane]) v[lane] = new_ctx(device);
    return v[lane];
}
}  // namespace

int32_t libsais_cuda_bwt_batch(const uint8_t *const *T, uint8_t *const *U, const int32_t *n, int32_t *primary, float *device_ms,
                               int32_t nblocks, const int32_t *devices, int32_t ndevices, int32_t lanes)
{
    if (T == nullptr || U == nullptr || n == nullptr || primary == nullptr || nblocks < 0 || ndevices < 0 || lanes < 0) return -1;
    if (nblocks == 0) return 0;
    int ndev_all = libsais_cuda_device_count();
    if (ndev_all <= 0) return -2;
    std::vector<int> devs;
    if (devices == nullptr || ndevices == 0) {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return -2; }
        devs.push_back(cur);
    } else {
        for (int i = 0; i < ndevices; ++i) { if (devices[i] < 0 || devices[i] >= ndev_all) return -1; devs.push_back(devices[i]); }
    }
    if (lanes == 0) lanes = 3;
    if (lanes > 8) lanes = 8;
    std::lock_guard<std::mutex> guard(g_batch_mutex);
    const int D = (int)devs.size();
    std::vector<Ctx *> ctxs((size_t)D * lanes, nullptr);
    for (int d = 0; d < D; ++d)
        for (int l = 0; l < lanes; ++l) {
            // two entries of `devices` may name the same GPU: they get distinct lanes of that GPU's pool
            int dup = 0;
            for (int e = 0; e < d; ++e) if (devs[e] == devs[d]) ++dup;
            Ctx *c = batch_ctx(devs[d], dup * lanes + l);
            if (!c) return -2;
            ctxs[(size_t)d * lanes + l] = c;
        }
    std::vector<std::atomic<int>> next(D);
    for (int d = 0; d < D; ++d) next[d].store(0);
    std::atomic<int> failed(0);
    static int32_t dummy_A[4];
    auto worker = [&](int d, int l) {
        Ctx *c = ctxs[(size_t)d * lanes + l];
        cudaSetDevice(c->device);
        for (;;) {
            const int j = next[d].fetch_add(1);
            const long long b = (long long)d + (long long)j * D;
            if (b >= nblocks) break;
            const int32_t rc = bwt_body<int32_t>(c, T[b], U[b], dummy_A, n[b], 0, nullptr, 0, nullptr, false);
            primary[b] = rc;
            if (device_ms) device_ms[b] = c->last_device_ms;
            if (rc < 0) failed.store(1);
        }
    };
    std::vector<std::thread> th;
    for (int d = 0; d < D; ++d)
        for (int l = 0; l < lanes; ++l) th.emplace_back(worker, d, l);
    for (auto &t : th) t.join();
    return failed.load() ? -2 : 0;
}

int64_t libsais_cuda_sa64_multi(const uint8_t *T, int64_t *SA, int64_t n, int64_t *freq, const int32_t *devices, int32_t ndevices,
                                libsais_cuda_dist_stats *stats)
{
    if (T == nullptr || n < 0 || ndevices < 1 || ndevices > 63) return -1;
    if (n < 2) { if (freq) host_freq(T, n, freq); if (n == 1 && SA) SA[0] = 0; return 0; }
    const int have = libsais_cuda_device_count();
    if (have <= 0) return -2;
    std::vector<int> devs;
    for (int i = 0; i < ndevices; ++i) {
        const int d = devices ? devices[i] : i % have;
        if (d < 0 || d >= have) return -1;
        devs.push_back(d);
    }
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = sa64_multi(T, (i64 *)SA, (u64)n, (i64 *)freq, devs.data(), ndevices, stats);
    if (stats) stats->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

// Free the pooled contexts of the batch entry point (device workspaces, streams).
void libsais_cuda_batch_release(void)
{
    std::lock_guard<std::mutex> guard(g_batch_mutex);
    for (auto &kv : g_batch_pool)
        for (Ctx *c : kv.second) if (c) { c->destroy(); delete c; }
    g_batch_pool.clear();
}

// ------------------------------------------------------------------ distributed building blocks
// (include/libsais_cuda.h; orchestrated by libsais_b200/dist.py with torch.distributed all-to-alls)
int64_t libsais_cuda_dist_prepare(const void *ctx, const uint8_t *d_T, int64_t n, int32_t *k_out, int32_t *key_bits_out)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !c->ok) return -2;
    if (d_T == nullptr || n <= 0 || k_out == nullptr || key_bits_out == nullptr) return -1;
    Call call(*c);
    int k = 0, K = 0;
    int rc = dist_prepare(*c, d_T, (u64)n, &k, &K);
    *k_out = k; *key_bits_out = K;
    return rc;
}
int64_t libsais_cuda_dist_keys(const void *ctx, int64_t lo, int64_t count, uint64_t *d_keys, uint32_t *d_pos)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !c->ok) return -2;
    if (lo < 0 || count < 0 || d_keys == nullptr || d_pos == nullptr) return -1;
    Call call(*c);
    int rc = dist_keys(*c, (u64)lo, (u64)count, d_keys, d_pos);
    return rc == 0 && call.finish() ? 0 : (rc ? rc : -2);
}
int64_t libsais_cuda_sort_pairs_dev(const void *ctx, uint64_t *d_keys, uint32_t *d_vals, uint64_t *d_keys_alt, uint32_t *d_vals_alt,
                                    int64_t count, int32_t lo_bit, int32_t hi_bit)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !c->ok) return -2;
    if (count < 0 || lo_bit < 0 || hi_bit > 64 || hi_bit < lo_bit) return -1;
    if (count == 0) return 0;
    Call call(*c);
    if (!c->reserve(sort_workspace_bytes((u64)count) + 4096)) return -2;
    int where = run_sort_pairs(*c, d_keys, d_vals, d_keys_alt, d_vals_alt, (u64)count, lo_bit, hi_bit);
    return where >= 0 && call.finish() ? where : -2;
}
int64_t libsais_cuda_sort_u32_pairs_dev(const void *ctx, uint32_t *d_keys, uint32_t *d_vals, uint32_t *d_keys_alt, uint32_t *d_vals_alt,
                                        int64_t count, int32_t lo_bit, int32_t hi_bit)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !c->ok) return -2;
    if (count < 0 || lo_bit < 0 || hi_bit > 32 || hi_bit < lo_bit) return -1;
    if (count == 0) return 0;
    Call call(*c);
    if (!c->reserve(sort_workspace_bytes((u64)count) + 4096)) return -2;
    int where = run_sort_u32_pairs(*c, d_keys, d_vals, d_keys_alt, d_vals_alt, (u64)count, lo_bit, hi_bit);
    return where >= 0 && call.finish() ? where : -2;
}
int64_t libsais_cuda_rank_stage_dev(const void *ctx, const uint64_t *d_keys, const uint32_t *d_pos, const uint32_t *d_slot_in,
                                    int64_t count, uint32_t slot_base, uint32_t *d_sa_local, uint32_t *d_pair_pos, uint32_t *d_pair_rank,
                                    uint32_t *d_act_pos, uint32_t *d_act_slot, uint32_t *d_act_grp, uint64_t *counts_out)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !c->ok) return -2;
    if (count < 0 || counts_out == nullptr) return -1;
    Call call(*c);
    if (!c->reserve(rank_stage_workspace_bytes((u64)count) + 4096)) return -2;
    u64 counts[2] = {0, 0};
    int rc = run_rank_stage(*c, d_keys, d_pos, d_slot_in, (u64)count, slot_base, d_sa_local, d_pair_pos, d_pair_rank,
                            d_act_pos, d_act_slot, d_act_grp, counts);
    counts_out[0] = counts[0]; counts_out[1] = counts[1];
    return rc == 0 && call.finish() ? 0 : -2;
}
int64_t libsais_cuda_gather_u32_dev(const void *ctx, const uint32_t *d_src, int64_t src_len, const uint32_t *d_idx, int64_t count,
                                    uint32_t idx_offset, uint32_t *d_out)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !c->ok) return -2;
    if (count < 0 || src_len < 0) return -1;
    Call call(*c);
    run_gather_u32(*c, d_src, (u64)src_len, d_idx, (u64)count, idx_offset, d_out);
    return call.finish() ? 0 : -2;
}
int64_t libsais_cuda_scatter_u32_dev(const void *ctx, uint32_t *d_dst, int64_t dst_len, const uint32_t *d_idx, const uint32_t *d_val,
                                     int64_t count, uint32_t idx_offset)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !c->ok) return -2;
    if (count < 0 || dst_len < 0) return -1;
    Call call(*c);
    c->reserve(scatter_workspace_bytes((u64)count) + 4096);      // best effort: without scratch the scatter is unpartitioned
    run_scatter_u32(*c, d_dst, (u64)dst_len, d_idx, d_val, (u64)count, idx_offset);
    return call.finish() ? 0 : -2;
}

int64_t libsais_cuda_dist_route_dev(const void *ctx, const uint32_t *d_a, const uint32_t *d_b, int64_t count, int64_t add, int64_t limit,
                                    int64_t block, int32_t world, uint32_t *d_a_out, uint32_t *d_b_out, uint64_t *counts_out)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !c->ok) return -2;
    if (count < 0 || add < 0 || limit < 0 || block <= 0 || world <= 0 || counts_out == nullptr) return -1;
    Call call(*c);
    if (!c->reserve(route_workspace_bytes((u64)count) + 4096)) return -2;
    u64 counts[64] = {0};
    int rc = run_route(*c, d_a, d_b, (u64)count, (u64)add, (u64)limit, (u64)block, (u32)world, d_a_out, d_b_out, counts);
    for (int r = 0; r < world && r < 64; ++r) counts_out[r] = counts[r];
    return rc == 0 && call.finish() ? 0 : (rc == -1 ? -1 : -2);
}
int64_t libsais_cuda_dist_partition_dev(const void *ctx, uint64_t *d_keys, uint32_t *d_pos, int64_t count, const uint64_t *d_splitters,
                                        int32_t nsplit, uint64_t *d_keys_out, uint32_t *d_pos_out, uint64_t *counts_out)
{
    Ctx *c = ctx ? as_ctx(ctx) : default_ctx();
    if (!c || !c->ok) return -2;
    if (count < 0 || nsplit < 0 || counts_out == nullptr) return -1;
    Call call(*c);
    if (!c->reserve(sort_workspace_bytes((u64)count) + 4096)) return -2;
    u64 counts[65] = {0};
    int rc = run_partition_by_splitters(*c, d_keys, d_pos, (u64)count, d_splitters, (u32)nsplit, d_keys_out, d_pos_out, counts);
    for (int r = 0; r <= nsplit && r < 65; ++r) counts_out[r] = counts[r];
    return rc == 0 && call.finish() ? 0 : (rc == -1 ? -1 : -2);
}

// ------------------------------------------------------------------ libsais16 / libsais16x64 (16-bit symbols)
void *libsais16_create_ctx(void) { return new_ctx(-1); }
void *libsais16_create_ctx_omp(int32_t threads) { return threads < 0 ? nullptr : new_ctx(-1); }
void libsais16_free_ctx(void *ctx) { libsais_free_ctx(ctx); }
void *libsais16_unbwt_create_ctx(void) { return new_ctx(-1); }
void *libsais16_unbwt_create_ctx_omp(int32_t threads) { return threads < 0 ? nullptr : new_ctx(-1); }
void libsais16_unbwt_free_ctx(void *ctx) { libsais_free_ctx(ctx); }

#define LSC16(IDX, P)                                                                                                                      \
    IDX P(const uint16_t *T, IDX *SA, IDX n, IDX fs, IDX *freq) { return sa16_body<IDX>(default_ctx(), T, SA, n, fs, freq, false); }       \
    IDX P##_gsa(const uint16_t *T, IDX *SA, IDX n, IDX fs, IDX *freq) { return sa16_body<IDX>(default_ctx(), T, SA, n, fs, freq, true); } \
    IDX P##_omp(const uint16_t *T, IDX *SA, IDX n, IDX fs, IDX *freq, IDX threads)                                                         \
    { if (threads < 0) return -1; return sa16_body<IDX>(default_ctx(), T, SA, n, fs, freq, false); }                                       \
    IDX P##_gsa_omp(const uint16_t *T, IDX *SA, IDX n, IDX fs, IDX *freq, IDX threads)                                                     \
    { if (threads < 0) return -1; return sa16_body<IDX>(default_ctx(), T, SA, n, fs, freq, true); }                                        \
    IDX P##_bwt(const uint16_t *T, uint16_t *U, IDX *A, IDX n, IDX fs, IDX *freq)                                                          \
    { return bwt16_body<IDX>(default_ctx(), T, U, A, n, fs, freq, 0, nullptr, false); }                                                    \
    IDX P##_bwt_aux(const uint16_t *T, uint16_t *U, IDX *A, IDX n, IDX fs, IDX *freq, IDX r, IDX *I)                                       \
    { return bwt16_body<IDX>(default_ctx(), T, U, A, n, fs, freq, r, I, true); }                                                           \
    IDX P##_bwt_omp(const uint16_t *T, uint16_t *U, IDX *A, IDX n, IDX fs, IDX *freq, IDX threads)                                         \
    { if (threads < 0) return -1; return bwt16_body<IDX>(default_ctx(), T, U, A, n, fs, freq, 0, nullptr, false); }                        \
    IDX P##_bwt_aux_omp(const uint16_t *T, uint16_t *U, IDX *A, IDX n, IDX fs, IDX *freq, IDX r, IDX *I, IDX threads)                      \
    { if (threads < 0) return -1; return bwt16_body<IDX>(default_ctx(), T, U, A, n, fs, freq, r, I, true); }                               \
    IDX P##_unbwt(const uint16_t *T, uint16_t *U, IDX *A, IDX n, const IDX *freq, IDX i)                                                   \
    { return unbwt16_body<IDX>(default_ctx(), T, U, A, n, freq, n, &i); }                                                                  \
    IDX P##_unbwt_aux(const uint16_t *T, uint16_t *U, IDX *A, IDX n, const IDX *freq, IDX r, const IDX *I)                                 \
    { return unbwt16_body<IDX>(default_ctx(), T, U, A, n, freq, r, I); }                                                                   \
    IDX P##_unbwt_omp(const uint16_t *T, uint16_t *U, IDX *A, IDX n, const IDX *freq, IDX i, IDX threads)                                  \
    { if (threads < 0) return -1; return unbwt16_body<IDX>(default_ctx(), T, U, A, n, freq, n, &i); }                                      \
    IDX P##_unbwt_aux_omp(const uint16_t *T, uint16_t *U, IDX *A, IDX n, const IDX *freq, IDX r, const IDX *I, IDX threads)                \
    { if (threads < 0) return -1; return unbwt16_body<IDX>(default_ctx(), T, U, A, n, freq, r, I); }                                       \
    IDX P##_plcp(const uint16_t *T, const IDX *SA, IDX *PLCP, IDX n) { return plcp16_body<IDX>(default_ctx(), T, SA, PLCP, n, false); }    \
    IDX P##_plcp_gsa(const uint16_t *T, const IDX *SA, IDX *PLCP, IDX n) { return plcp16_body<IDX>(default_ctx(), T, SA, PLCP, n, true); } \
    IDX P##_lcp(const IDX *PLCP, const IDX *SA, IDX *LCP, IDX n) { return lcp_body<IDX>(default_ctx(), PLCP, SA, LCP, n); }                \
    IDX P##_plcp_omp(const uint16_t *T, const IDX *SA, IDX *PLCP, IDX n, IDX threads)                                                      \
    { if (threads < 0) return -1; return plcp16_body<IDX>(default_ctx(), T, SA, PLCP, n, false); }                                         \
    IDX P##_plcp_gsa_omp(const uint16_t *T, const IDX *SA, IDX *PLCP, IDX n, IDX threads)                                                  \
    { if (threads < 0) return -1; return plcp16_body<IDX>(default_ctx(), T, SA, PLCP, n, true); }                                          \
    IDX P##_lcp_omp(const IDX *PLCP, const IDX *SA, IDX *LCP, IDX n, IDX threads)                                                          \
    { if (threads < 0) return -1; return lcp_body<IDX>(default_ctx(), PLCP, SA, LCP, n); }
LSC16(int32_t, libsais16)
LSC16(int64_t, libsais16x64)
#undef LSC16

int32_t libsais16_int(int32_t *T, int32_t *SA, int32_t n, int32_t k, int32_t fs) { return sa_int_body<int32_t, int32_t>(default_ctx(), T, SA, n, k, fs); }
int32_t libsais16_int_omp(int32_t *T, int32_t *SA, int32_t n, int32_t k, int32_t fs, int32_t threads)
{ if (threads < 0) return -1; return sa_int_body<int32_t, int32_t>(default_ctx(), T, SA, n, k, fs); }
int64_t libsais16x64_long(int64_t *T, int64_t *SA, int64_t n, int64_t k, int64_t fs) { return sa_int_body<int64_t, int64_t>(default_ctx(), T, SA, n, k, fs); }
int64_t libsais16x64_long_omp(int64_t *T, int64_t *SA, int64_t n, int64_t k, int64_t fs, int64_t threads)
{ if (threads < 0) return -1; return sa_int_body<int64_t, int64_t>(default_ctx(), T, SA, n, k, fs); }
int32_t libsais16_ctx(const void *ctx, const uint16_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq)
{ if (ctx == nullptr) return -1; return sa16_body<int32_t>(as_ctx(ctx), T, SA, n, fs, freq, false); }
int32_t libsais16_gsa_ctx(const void *ctx, const uint16_t *T, int32_t *SA, int32_t n, int32_t fs, int32_t *freq)
{ if (ctx == nullptr) return -1; return sa16_body<int32_t>(as_ctx(ctx), T, SA, n, fs, freq, true); }
int32_t libsais16_bwt_ctx(const void *ctx, const uint16_t *T, uint16_t *U, int32_t *A, int32_t n, int32_t fs, int32_t *freq)
{ if (ctx == nullptr) return -1; return bwt16_body<int32_t>(as_ctx(ctx), T, U, A, n, fs, freq, 0, nullptr, false); }
int32_t libsais16_bwt_aux_ctx(const void *ctx, const uint16_t *T, uint16_t *U, int32_t *A, int32_t n, int32_t fs, int32_t *freq, int32_t r, int32_t *I)
{ if (ctx == nullptr) return -1; return bwt16_body<int32_t>(as_ctx(ctx), T, U, A, n, fs, freq, r, I, true); }
int32_t libsais16_unbwt_ctx(const void *ctx, const uint16_t *T, uint16_t *U, int32_t *A, int32_t n, const int32_t *freq, int32_t i)
{ if (ctx == nullptr) return -1; return unbwt16_body<int32_t>(as_ctx(ctx), T, U, A, n, freq, n, &i); }
int32_t libsais16_unbwt_aux_ctx(const void *ctx, const uint16_t *T, uint16_t *U, int32_t *A, int32_t n, const int32_t *freq, int32_t r, const int32_t *I)
{ if (ctx == nullptr) return -1; return unbwt16_body<int32_t>(as_ctx(ctx), T, U, A, n, freq, r, I); }

}  // extern "C"
