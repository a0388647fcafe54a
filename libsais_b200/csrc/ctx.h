// ctx.h -- the libsais_cuda context: one GPU, one stream, one grow-only device workspace.
// Plays the role of the reference's LIBSAIS_CONTEXT / LIBSAIS_UNBWT_CONTEXT
// (reference src/libsais.c:86-99, :237-256, :7326-7349): reusable scratch owned by the caller.
#pragma once
#include <vector>
#include <string>
#include "common.cuh"

namespace lsc {

struct RoundStat {
    u64 h;          // prefix length already sorted when the round started (0 for the initial sort)
    u64 n_active;   // suffixes sorted in this round
    u64 n_groups;   // non-singleton groups they formed
    int passes;     // radix digit passes executed
    int key_bits;   // significant key bits sorted
    double ms = 0;      // with profiling: device time of the kernels launched in the round (CUDA events)
    double bytes = 0;   //                 and their algorithmic bytes
};

struct Ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool ok = false;

    // grow-only workspace arena, bump-allocated per call
    char *ws = nullptr;
    size_t ws_cap = 0, ws_off = 0;

    // small persistent scalars: device array + pinned host mirror
    u64 *d_scalars = nullptr;     // kNumScalars
    u64 *h_scalars = nullptr;
    static const int kNumScalars = 1024;

    // pinned staging for small host<->device transfers (freq, aux indexes up to this size)
    cudaError_t last_error = cudaSuccess;

    // launch accounting
    bool profiling = false;
    int pass_class_override = -1;         // account onesweep launches to another kernel class (partitioned scatter)
    struct Pending { int kc; cudaEvent_t a, b; double bytes; int round; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> event_pool;
    u64 launches[KC_COUNT] = {0};
    double ms[KC_COUNT] = {0};
    double bytes[KC_COUNT] = {0};
    u64 total_launches = 0;
    std::vector<RoundStat> rounds;
    float last_device_ms = 0.f;   // first-kernel -> last-kernel device time of the last *_dev call

    // distributed building blocks: packed text kept across calls (sa_core.cu dist_prepare)
    u64 *dist_words = nullptr; u64 dist_n = 0; int dist_b = 0, dist_k = 0;

    // second stream + events for results that leave while the call still computes (streamed BWT rows: sa_core.cu, api.cu)
    cudaStream_t copy_stream = nullptr;
    static const int kChunkEvents = 16;
    cudaEvent_t chunk_ev[kChunkEvents] = {nullptr};
    bool ensure_copy_stream();

    // pinned staging lanes for large pageable host transfers (hostcopy.cu)
    struct StageLane { cudaStream_t stream = nullptr; void *buf[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr}; };
    std::vector<StageLane> stage;
    bool ensure_staging(int workers);
    void free_staging();

    bool init(int dev);
    void destroy();

    bool reserve(size_t bytes);           // make the arena at least this large (before a call)
    void reset_arena() { ws_off = 0; }
    void *alloc(size_t bytes);            // bump allocate (256-B aligned); nullptr when exhausted
    template <typename T> T *alloc_n(size_t n) { return (T *)alloc(n * sizeof(T)); }

    void reset_stats();
    void begin(int kc, double algo_bytes);
    void end();
    void resolve_profile();               // after a stream sync: fold pending event pairs into ms[]
    bool check(cudaError_t e) { if (e != cudaSuccess && last_error == cudaSuccess) last_error = e; return e == cudaSuccess; }
    bool failed() const { return last_error != cudaSuccess; }
    bool sync() { return check(cudaStreamSynchronize(stream)); }
};

// RAII: make the ctx's device current for the duration of a call, restore afterwards.
struct DeviceGuard {
    int prev = -1; bool changed = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) { changed = cudaSetDevice(dev) == cudaSuccess; }
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};

#define LSC_LAUNCH(ctx, kc, algo_bytes, kernel, grid, block, smem, ...)                      \
    do {                                                                                      \
        (ctx).begin((kc), (double)(algo_bytes));                                              \
        kernel<<<(grid), (block), (smem), (ctx).stream>>>(__VA_ARGS__);                       \
        (ctx).end();                                                                          \
    } while (0)

}  // namespace lsc
