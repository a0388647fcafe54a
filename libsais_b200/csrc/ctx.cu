// ctx.cu -- context lifetime, workspace arena and launch accounting.
#include "ctx.h"
#include <cstring>

namespace lsc {

bool Ctx::init(int dev)
{
    device = dev;
    if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return false; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return false; }
    sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaMalloc(&d_scalars, kNumScalars * sizeof(u64)) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaMallocHost(&h_scalars, kNumScalars * sizeof(u64)) != cudaSuccess) { cudaGetLastError(); return false; }
    std::memset(h_scalars, 0, kNumScalars * sizeof(u64));
    ok = true;
    return true;
}

void Ctx::destroy()
{
    DeviceGuard g(device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto &p : pending) { event_pool.push_back(p.a); event_pool.push_back(p.b); }
    pending.clear();
    for (auto e : event_pool) cudaEventDestroy(e);
    event_pool.clear();
    free_staging();
    for (int i = 0; i < kChunkEvents; ++i) if (chunk_ev[i]) { cudaEventDestroy(chunk_ev[i]); chunk_ev[i] = nullptr; }
    if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); copy_stream = nullptr; }
    if (dist_words) { cudaFree(dist_words); dist_words = nullptr; }
    if (ws) cudaFree(ws);
    if (d_scalars) cudaFree(d_scalars);
    if (h_scalars) cudaFreeHost(h_scalars);
    if (stream) cudaStreamDestroy(stream);
    ws = nullptr; d_scalars = nullptr; h_scalars = nullptr; stream = nullptr; ok = false;
    cudaGetLastError();
}

bool Ctx::ensure_copy_stream()
{
    if (copy_stream) return true;
    if (cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); copy_stream = nullptr; return false; }
    for (int i = 0; i < kChunkEvents; ++i)
        if (cudaEventCreateWithFlags(&chunk_ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
    return true;
}

bool Ctx::reserve(size_t bytes)
{
    if (bytes <= ws_cap) return true;
    if (ws) { cudaStreamSynchronize(stream); cudaFree(ws); ws = nullptr; ws_cap = 0; }
    size_t want = bytes + (bytes >> 4);          // a little slack so near-equal sizes do not re-allocate
    if (cudaMalloc(&ws, want) != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        if (cudaMalloc(&ws, want) != cudaSuccess) { cudaGetLastError(); ws = nullptr; return false; }
    }
    ws_cap = want;
    return true;
}

void *Ctx::alloc(size_t bytes)
{
    size_t a = (ws_off + 255) & ~(size_t)255;
    if (a + bytes > ws_cap) { if (last_error == cudaSuccess) last_error = cudaErrorMemoryAllocation; return nullptr; }
    ws_off = a + bytes;
    return ws + a;
}

void Ctx::reset_stats()
{
    for (auto &p : pending) { event_pool.push_back(p.a); event_pool.push_back(p.b); }
    pending.clear();
    for (int i = 0; i < KC_COUNT; ++i) { launches[i] = 0; ms[i] = 0; bytes[i] = 0; }
    total_launches = 0;
    rounds.clear();
    last_error = cudaSuccess;
}

void Ctx::begin(int kc, double algo_bytes)
{
    launches[kc]++; total_launches++;
    bytes[kc] += algo_bytes;
    if (!profiling) return;
    Pending p; p.kc = kc; p.bytes = algo_bytes; p.round = (int)rounds.size();      // the round being built (its stat is pushed when it ends)
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; ++i) {
        if (!event_pool.empty()) { ev[i] = event_pool.back(); event_pool.pop_back(); }
        else cudaEventCreate(&ev[i]);
    }
    p.a = ev[0]; p.b = ev[1];
    cudaEventRecord(p.a, stream);
    pending.push_back(p);
}

void Ctx::end()
{
    check(cudaGetLastError());
    if (profiling && !pending.empty()) cudaEventRecord(pending.back().b, stream);
}

void Ctx::resolve_profile()
{
    for (auto &p : pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) {
            ms[p.kc] += t;
            if (p.round >= 0 && (size_t)p.round < rounds.size()) { rounds[p.round].ms += t; rounds[p.round].bytes += p.bytes; }
        } else cudaGetLastError();
        event_pool.push_back(p.a); event_pool.push_back(p.b);
    }
    pending.clear();
}

}  // namespace lsc
