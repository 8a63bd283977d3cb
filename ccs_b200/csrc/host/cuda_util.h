// Small CUDA runtime helpers for the host engine: grow-only device / pinned-host buffers and
// error plumbing.  (Host C++ only; compiled by g++ against the CUDA runtime headers.)
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace ccs {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};
struct OomError : std::runtime_error {
    explicit OomError(const std::string& s) : std::runtime_error(s) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        std::snprintf(buf, sizeof(buf), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
        if (e == cudaErrorMemoryAllocation) throw OomError(buf);
        throw CudaError(buf);
    }
}
#define CCS_CUDA(x) ::ccs::cuda_check((x), #x, __FILE__, __LINE__)

// Wait for a stream WITHOUT spinning: cudaStreamSynchronize busy-waits by default, and a stage context keeps one host
// thread per lane waiting on its stream most of the time -- ten spinning threads per GPU starve the host logic of the
// other ranks on a shared box (measured: 4 GPUs on one box dropped to 63 % per-GPU throughput).  A blocking-sync event
// puts the thread to sleep until the GPU interrupt arrives.
// The event is owned by the caller (an engine creates one with make_blocking_event() and destroys it with the engine):
// a thread_local event would leak one event per short-lived lane thread.
inline cudaError_t make_blocking_event(cudaEvent_t* ev) {
    return cudaEventCreateWithFlags(ev, cudaEventBlockingSync | cudaEventDisableTiming);
}
inline cudaError_t stream_sync_blocking(cudaStream_t s, cudaEvent_t ev) {
    cudaError_t e = cudaEventRecord(ev, s);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev);
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    size_t* budget_used = nullptr;   // optional accounting against the ctx budget
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    void release() {
        if (p) { cudaFree(p); if (budget_used) *budget_used -= cap * sizeof(T); }
        p = nullptr; cap = 0;
    }
    // grow-only; contents are NOT preserved
    void ensure(size_t n, size_t budget = 0) {
        if (n <= cap) return;
        size_t want = n + n / 4 + 64;   // generous slack: a regrow frees + reallocates and stalls every lane
        if (budget && budget_used && *budget_used - cap * sizeof(T) + want * sizeof(T) > budget) want = n;
        if (budget && budget_used && *budget_used - cap * sizeof(T) + want * sizeof(T) > budget)
            throw OomError("device budget exceeded");
        release();
        CCS_CUDA(cudaMalloc((void**)&p, want * sizeof(T)));
        cap = want;
        if (budget_used) *budget_used += cap * sizeof(T);
    }
};

template <class T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    PinBuf() = default;
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    void ensure(size_t n) {
        if (n <= cap) return;
        const size_t want = n + n / 4 + 64;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        CCS_CUDA(cudaHostAlloc((void**)&p, want * sizeof(T), cudaHostAllocDefault));
        cap = want;
    }
};

}  // namespace ccs
