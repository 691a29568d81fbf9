// rt3_rt.h — thin runtime layer under the C ABI: device memory, streams, events, launch macros
// and error propagation (status code + thread-local message, no exceptions across the ABI;
// reference behaviour being replaced: RENDERTOY3O_CUDA_CHECK throwing, src/util/exception.h:11-96).
//
// Two flavours of the same sources:
//   * default (nvcc, sm_100a): the product, librt3.so.  Kernels run on the GPU, period.
//   * RT3_EMULATE (g++): tests/emul's kernel-logic simulator — each "launch" is a host loop over
//     thread ids, device memory is host memory.  Test infrastructure for the GPU-less CI box only;
//     it is never linked into librt3.so and is not reachable from the product API.
#pragma once
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>

#ifndef RT3_EMULATE
#include <cuda_runtime.h>
#endif

namespace rt3 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

extern thread_local std::string g_last_error;
extern std::atomic<unsigned long long> g_launch_count;  // kernels launched by this library (stats.kernel_launches)
inline void count_launch() { ++g_launch_count; }

#define RT3_REQUIRE(cond, code, msg)                                         \
    do {                                                                     \
        if (!(cond)) throw ::rt3::Error((code), std::string(msg));           \
    } while (0)

#ifdef RT3_EMULATE
// ------------------------------------------------------------------------------------ emulation
typedef int Stream;
struct Event { double t = 0; };
#define RT3_CUDA(call) (void)0
inline void* dev_alloc(size_t bytes) {  // 256-byte aligned like cudaMalloc
    const size_t n = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
    void* p = aligned_alloc(256, n);
    if (!p) throw Error(-2, "emul: out of memory");
    memset(p, 0, n);
    return p;
}
inline void dev_free(void* p) { free(p); }
inline void h2d(void* d, const void* h, size_t n, Stream) { if (n) memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n, Stream) { if (n) memcpy(h, d, n); }
inline void d2d(void* d, const void* s, size_t n, Stream) { if (n) memcpy(d, s, n); }
inline void dev_memset(void* d, int v, size_t n, Stream) { if (n) memset(d, v, n); }
inline void stream_sync(Stream) {}
inline void event_record(Event&, Stream) {}
inline bool event_recorded(const Event&) { return false; }
inline void stream_wait(Stream, Event&) {}
inline float event_ms(Event&, Event&) { return 0.0f; }

#define RT3_GLOBAL(name, ...) static void name(uint32_t rt3_tid_, uint32_t rt3_n_, __VA_ARGS__)
#define RT3_THREAD_ID() rt3_tid_
#define RT3_LAUNCH_1D(name, n, st, ...)                                                                      \
    do {                                                                                                     \
        for (uint32_t i_ = 0; i_ < (uint32_t)(n); ++i_) name(i_, (uint32_t)(n), __VA_ARGS__);                \
        ::rt3::count_launch();                                                                               \
    } while (0)
#else
// ------------------------------------------------------------------------------------ CUDA (product)
typedef cudaStream_t Stream;
#define RT3_CUDA(call)                                                                                           \
    do {                                                                                                         \
        cudaError_t e_ = (call);                                                                                 \
        if (e_ != cudaSuccess) {                                                                                 \
            char buf_[512];                                                                                      \
            snprintf(buf_, sizeof(buf_), "CUDA call '%s' failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw ::rt3::Error(-2, buf_);                                                                        \
        }                                                                                                        \
    } while (0)
struct Event {
    cudaEvent_t e = nullptr;
    Event() {}
    ~Event() { if (e) cudaEventDestroy(e); }
    Event(const Event&) = delete;
    Event& operator=(const Event&) = delete;
};
inline void* dev_alloc(size_t bytes) { void* p = nullptr; RT3_CUDA(cudaMalloc(&p, bytes ? bytes : 16)); return p; }
inline void dev_free(void* p) { if (p) cudaFree(p); }
inline void h2d(void* d, const void* h, size_t n, Stream s) { if (n) RT3_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
inline void d2h(void* h, const void* d, size_t n, Stream s) { if (n) RT3_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline void d2d(void* d, const void* s_, size_t n, Stream s) { if (n) RT3_CUDA(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); }
inline void dev_memset(void* d, int v, size_t n, Stream s) { if (n) RT3_CUDA(cudaMemsetAsync(d, v, n, s)); }
inline void stream_sync(Stream s) { RT3_CUDA(cudaStreamSynchronize(s)); }
inline void event_record(Event& e, Stream s) { if (!e.e) RT3_CUDA(cudaEventCreate(&e.e)); RT3_CUDA(cudaEventRecord(e.e, s)); }
inline bool event_recorded(const Event& e) { return e.e != nullptr; }
inline void stream_wait(Stream s, Event& e) { RT3_CUDA(cudaStreamWaitEvent(s, e.e, 0)); }  // e must have been recorded
inline float event_ms(Event& a, Event& b) { float ms = 0; RT3_CUDA(cudaEventElapsedTime(&ms, a.e, b.e)); return ms; }

#define RT3_GLOBAL(name, ...) __global__ void name(uint32_t rt3_n_, __VA_ARGS__)
#define RT3_THREAD_ID() (blockIdx.x * blockDim.x + threadIdx.x)
#define RT3_LAUNCH_1D(name, n, st, ...)                                                        \
    do {                                                                                       \
        if ((n) > 0) {                                                                         \
            name<<<(unsigned)(((size_t)(n) + 255) / 256), 256, 0, st>>>((uint32_t)(n), __VA_ARGS__); \
            RT3_CUDA(cudaGetLastError());                                                      \
            ::rt3::count_launch();                                                             \
        }                                                                                      \
    } while (0)
#endif

// RAII device buffer (reference: CUDABuffer<T>, src/cuda/cuda_buffer.h:14-49)
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void alloc(size_t count) { release(); p = (T*)dev_alloc(count * sizeof(T)); n = count; }
    void ensure(size_t count) { if (count > n) alloc(count); }
    void release() { if (p) dev_free(p); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(T); }
};

}  // namespace rt3
