// DEVELOPMENT TOOL, NOT A PRODUCT PATH.  A stand-in for <cuda_runtime.h> that lets g++ compile the engine's CUDA
// sources (after scripts/emu/prep.py rewrote the <<< >>> launches) into a functional model that runs on the host:
// every CUDA thread is a fibre, a block's fibres are scheduled round-robin (so __syncthreads and warp shuffles have
// their CUDA meaning), blocks run one after another, streams execute in issue order, "device memory" is host memory.
// It exists so that the gpu-marked tests can be dry-run in the GPU-less build container (scripts/emu/run.sh copies
// the repository to /tmp and builds there); nothing in the package, bench.py or the tests loads it.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <tuple>
#include <utility>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
    dim3(unsigned long long x_) : x((unsigned)x_), y(1), z(1) {}
    dim3(long long x_) : x((unsigned)x_), y(1), z(1) {}
    dim3(int x_) : x((unsigned)x_), y(1), z(1) {}
    dim3(long x_) : x((unsigned)x_), y(1), z(1) {}
    dim3(unsigned long x_) : x((unsigned)x_), y(1), z(1) {}
};
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef struct EmuStream *cudaStream_t;
typedef struct EmuEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocMapped = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaFuncAttributes { int numRegs; };

namespace emu {
struct ThreadCtx { uint3 tid, bid; dim3 bdim, gdim; };
ThreadCtx &ctx();
void syncthreads();
void warpExchange(const void *mine, void *theirs, size_t size, unsigned srcLane, unsigned mask);
unsigned long long nowNs();
void yield();
struct Cfg {
    dim3 g, b;
    cudaStream_t stream;
    Cfg(dim3 g_, dim3 b_, size_t = 0, cudaStream_t s = nullptr) : g(g_), b(b_), stream(s) {}
};
void runGrid(const Cfg &cfg, void (*tramp)(void *), void *closure, const void *kernelId);
// stream order: the operation runs on the stream's worker after everything enqueued before it (nullptr: at once)
void enqueue(cudaStream_t stream, std::function<void()> op);
template <class... P>
struct Bound {
    void (*k)(P...);
    Cfg cfg;
    template <class... A>
    void operator()(A &&...args) const
    {
        // like a real launch, the arguments are copied when the launch is issued
        std::tuple<std::decay_t<P>...> copy(std::forward<A>(args)...);
        void (*kernel)(P...) = k;
        const Cfg c = cfg;
        enqueue(c.stream, [copy, kernel, c]() {
            auto call = [&]() { std::apply(kernel, copy); };
            runGrid(c, [](void *f) { (*static_cast<decltype(call) *>(f))(); }, &call, (const void *)kernel);
        });
    }
};
template <class... P>
Bound<P...> bind(void (*k)(P...), const Cfg &cfg) { return Bound<P...>{k, cfg}; }
} // namespace emu

#define threadIdx (emu::ctx().tid)
#define blockIdx (emu::ctx().bid)
#define blockDim (emu::ctx().bdim)
#define gridDim (emu::ctx().gdim)

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { emu::yield(); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta)
{
    const unsigned lane = emu::ctx().tid.x & 31u;
    T out = v;
    emu::warpExchange(&v, &out, sizeof(T), lane + delta < 32u ? lane + delta : lane, mask);
    return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, unsigned laneMask)
{
    const unsigned lane = emu::ctx().tid.x & 31u;
    T out = v;
    emu::warpExchange(&v, &out, sizeof(T), lane ^ laneMask, mask);
    return out;
}
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicMax(T *p, T v)
{
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
using std::max;
using std::min;

extern "C" {
cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned flags);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaHostGetDevicePointer(void **dev, void *host, unsigned flags);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, int kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, int kind, cudaStream_t s);
cudaError_t cudaMemset(void *dst, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *dst, int v, size_t bytes, cudaStream_t s);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaGetLastError(void);
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
}
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned flags) { return cudaHostAlloc((void **)p, bytes, flags); }
template <class K> static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, K) { a->numRegs = 0; return cudaSuccess; }
