// hostcopy.cu -- see hostcopy.h
#include "hostcopy.h"
#include <cstring>
#include <cstdlib>
#include <thread>
#include <vector>

namespace lsc {

static const size_t kChunk = 8u << 20;          // bytes per staged chunk
static const size_t kStageMin = 16u << 20;      // below this a direct copy is as good
static const int kMaxWorkers = 6;

static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static int stage_workers()
{
    static int w = -1;
    if (w < 0) {
        const char *e = getenv("LIBSAIS_CUDA_COPY_THREADS");
        int v = (e && *e) ? atoi(e) : 4;
        unsigned hc = std::thread::hardware_concurrency();
        if (hc && (unsigned)v > hc) v = (int)hc;
        w = v < 0 ? 0 : (v > kMaxWorkers ? kMaxWorkers : v);
    }
    return w;
}

bool Ctx::ensure_staging(int workers)
{
    if ((int)stage.size() >= workers) return true;
    while ((int)stage.size() < workers) {
        StageLane lane;
        if (cudaStreamCreateWithFlags(&lane.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
        for (int i = 0; i < 2; ++i) {
            if (cudaMallocHost(&lane.buf[i], kChunk) != cudaSuccess) { cudaGetLastError(); return false; }
            if (cudaEventCreateWithFlags(&lane.ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
        }
        stage.push_back(lane);
    }
    return true;
}

void Ctx::free_staging()
{
    for (auto &l : stage) {
        for (int i = 0; i < 2; ++i) { if (l.buf[i]) cudaFreeHost(l.buf[i]); if (l.ev[i]) cudaEventDestroy(l.ev[i]); }
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    stage.clear();
}

// worker w moves chunks w, w+W, w+2W, ... through its two pinned slots
static void lane_h2d(Ctx::StageLane *l, int device, char *d, const char *h, size_t bytes, int w, int W, int *ok)
{
    cudaSetDevice(device);
    size_t nchunks = (bytes + kChunk - 1) / kChunk;
    int slot = 0;
    bool used[2] = {false, false};
    for (size_t ci = (size_t)w; ci < nchunks; ci += (size_t)W, slot ^= 1) {
        size_t off = ci * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
        if (used[slot] && cudaEventSynchronize(l->ev[slot]) != cudaSuccess) { *ok = 0; return; }
        std::memcpy(l->buf[slot], h + off, len);
        if (cudaMemcpyAsync(d + off, l->buf[slot], len, cudaMemcpyHostToDevice, l->stream) != cudaSuccess) { *ok = 0; return; }
        cudaEventRecord(l->ev[slot], l->stream);
        used[slot] = true;
    }
    if (cudaStreamSynchronize(l->stream) != cudaSuccess) *ok = 0;
}

static void lane_d2h(Ctx::StageLane *l, int device, char *h, const char *d, size_t bytes, int w, int W, int *ok)
{
    cudaSetDevice(device);
    size_t nchunks = (bytes + kChunk - 1) / kChunk;
    // software pipeline: DMA of chunk k+1 runs while chunk k is memcpy'd out of its pinned slot
    size_t pend_off[2] = {0, 0}, pend_len[2] = {0, 0};
    bool pend[2] = {false, false};
    int slot = 0;
    for (size_t ci = (size_t)w; ci < nchunks; ci += (size_t)W, slot ^= 1) {
        size_t off = ci * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
        if (pend[slot]) {                                   // this slot still holds an undelivered chunk
            if (cudaEventSynchronize(l->ev[slot]) != cudaSuccess) { *ok = 0; return; }
            std::memcpy(h + pend_off[slot], l->buf[slot], pend_len[slot]);
            pend[slot] = false;
        }
        if (cudaMemcpyAsync(l->buf[slot], d + off, len, cudaMemcpyDeviceToHost, l->stream) != cudaSuccess) { *ok = 0; return; }
        cudaEventRecord(l->ev[slot], l->stream);
        pend[slot] = true; pend_off[slot] = off; pend_len[slot] = len;
        int other = slot ^ 1;
        if (pend[other]) {
            if (cudaEventSynchronize(l->ev[other]) != cudaSuccess) { *ok = 0; return; }
            std::memcpy(h + pend_off[other], l->buf[other], pend_len[other]);
            pend[other] = false;
        }
    }
    for (int s = 0; s < 2; ++s) if (pend[s]) {
        if (cudaEventSynchronize(l->ev[s]) != cudaSuccess) { *ok = 0; return; }
        std::memcpy(h + pend_off[s], l->buf[s], pend_len[s]);
    }
}

bool host_is_pinned(const void *p) { return is_pinned(p); }

bool copy_h2d(Ctx &c, void *d_dst, const void *h_src, size_t bytes)
{
    if (bytes == 0) return true;
    const int W = stage_workers();
    if (bytes < kStageMin || W == 0 || is_pinned(h_src) || !c.ensure_staging(W))
        return c.check(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c.stream));
    // earlier work on the context's stream may still read/write the destination arena: finish it first
    if (!c.sync()) return false;
    int ok[kMaxWorkers];
    std::vector<std::thread> th;
    for (int w = 0; w < W; ++w) { ok[w] = 1; th.emplace_back(lane_h2d, &c.stage[w], c.device, (char *)d_dst, (const char *)h_src, bytes, w, W, &ok[w]); }
    for (auto &t : th) t.join();
    for (int w = 0; w < W; ++w) if (!ok[w]) return c.check(cudaErrorUnknown);
    return true;                                            // lanes synchronised their streams: data is resident
}

bool copy_d2h(Ctx &c, void *h_dst, const void *d_src, size_t bytes)
{
    if (bytes == 0) return true;
    const int W = stage_workers();
    if (bytes < kStageMin || W == 0 || is_pinned(h_dst) || !c.ensure_staging(W))
        return c.check(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c.stream));
    if (!c.sync()) return false;                            // producers of d_src run on the context's stream
    int ok[kMaxWorkers];
    std::vector<std::thread> th;
    for (int w = 0; w < W; ++w) { ok[w] = 1; th.emplace_back(lane_d2h, &c.stage[w], c.device, (char *)h_dst, (const char *)d_src, bytes, w, W, &ok[w]); }
    for (auto &t : th) t.join();
    for (int w = 0; w < W; ++w) if (!ok[w]) return c.check(cudaErrorUnknown);
    return true;
}

}  // namespace lsc
