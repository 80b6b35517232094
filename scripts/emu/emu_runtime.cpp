// DEVELOPMENT TOOL (see include/cuda_runtime.h): host-side model of the CUDA execution the engine relies on.
//   * a block's threads are ucontext fibres on the calling OS thread, scheduled round-robin; a fibre gives up the
//     processor only in __syncthreads, a warp exchange or __nanosleep, so kernels without those run straight through.
//     A kernel whose first block never yielded is run as plain calls from then on (fast); if it yields later after
//     all, the model stops with a message rather than guessing.
//   * blocks run in ascending order (the hardware's dispatch order, which the fused peer step relies on for progress)
//   * streams and events execute at issue time; cudaMalloc is an aligned host allocation with red zones that
//     cudaFree and the library destructor check
#include "cuda_runtime.h"
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>
#include <atomic>
#include <cstdio>
#include <ctime>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <set>
#include <thread>
#include <vector>

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define EMU_ASAN 1
#endif

namespace emu {
namespace {
constexpr size_t kStack = 256 * 1024;
struct Fibre {
    ucontext_t uc;
    char *stack = nullptr;
    ThreadCtx tc;
    bool done = true;
};
struct BlockState {
    std::vector<Fibre> fibres;
    unsigned nThreads = 0;
    ucontext_t sched;
    int current = -1;
    int alive = 0;
    // block barrier
    int arrived = 0;
    unsigned long long generation = 0;
    // warp exchange: per warp staging + two-phase barrier
    struct Warp {
        alignas(16) unsigned char slot[32][16];
        int arrived = 0;
        unsigned long long generation = 0;
    };
    std::vector<Warp> warps;
    void (*tramp)(void *) = nullptr;
    void *closure = nullptr;
    const void *schedStack = nullptr; // AddressSanitizer: the scheduler's stack, learned on the first switch into a fibre
    size_t schedStackSize = 0;
    bool yielded = false;   // some thread of the current launch used a barrier / exchange / sleep
    bool direct = false;    // threads are plain calls: no fibre to switch away from
    ThreadCtx directCtx;
};
std::mutex g_kernelMu;
std::map<const void *, bool> g_kernelYields;
thread_local BlockState *g_bs = nullptr;
thread_local ThreadCtx g_hostCtx;

// fibre -> scheduler (last == true: the fibre is finished and its stack will be reused)
void toScheduler(BlockState *bs, Fibre &f, bool last)
{
#ifdef EMU_ASAN
    void *fake = nullptr;
    __sanitizer_start_switch_fiber(last ? nullptr : &fake, bs->schedStack, bs->schedStackSize);
    swapcontext(&f.uc, &bs->sched);
    __sanitizer_finish_switch_fiber(fake, &bs->schedStack, &bs->schedStackSize);
#else
    (void)last;
    swapcontext(&f.uc, &bs->sched);
#endif
}
// scheduler -> fibre
void toFibre(BlockState *bs, Fibre &f)
{
#ifdef EMU_ASAN
    void *fake = nullptr;
    __sanitizer_start_switch_fiber(&fake, f.stack, kStack);
    swapcontext(&bs->sched, &f.uc);
    __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#else
    swapcontext(&bs->sched, &f.uc);
#endif
}

void fibreMain()
{
    BlockState *bs = g_bs;
#ifdef EMU_ASAN
    __sanitizer_finish_switch_fiber(nullptr, &bs->schedStack, &bs->schedStackSize);
#endif
    Fibre &f = bs->fibres[bs->current];
    bs->tramp(bs->closure);
    f.done = true;
    bs->alive--;
    // a thread that exits no longer takes part in barriers
    if (bs->arrived > 0 && bs->arrived >= bs->alive) { bs->arrived = 0; bs->generation++; }
    toScheduler(bs, f, true);
}
} // namespace

ThreadCtx &ctx()
{
    BlockState *bs = g_bs;
    if (!bs) return g_hostCtx;
    if (bs->direct) return bs->directCtx;
    return bs->current >= 0 ? bs->fibres[bs->current].tc : g_hostCtx;
}

static void noteYield(BlockState *bs)
{
    if (bs->direct) { fprintf(stderr, "emu: a kernel classified as barrier-free reached a barrier / exchange / sleep (block %u)\n", bs->directCtx.bid.x); abort(); }
    bs->yielded = true;
}

void yield()
{
    BlockState *bs = g_bs;
    // a sleeping thread waits for another stream (another OS thread here), never for a thread of its own block
    if (!bs || bs->direct || bs->current < 0) { sched_yield(); return; }
    toScheduler(bs, bs->fibres[bs->current], false);
}

void syncthreads()
{
    BlockState *bs = g_bs;
    noteYield(bs);
    const unsigned long long gen = bs->generation;
    if (++bs->arrived >= bs->alive) { bs->arrived = 0; bs->generation++; return; }
    while (bs->generation == gen) yield();
}

static void warpBarrier(BlockState *bs, BlockState::Warp &w, int participants)
{
    const unsigned long long gen = w.generation;
    if (++w.arrived >= participants) { w.arrived = 0; w.generation++; return; }
    while (w.generation == gen) yield();
}

void warpExchange(const void *mine, void *theirs, size_t size, unsigned srcLane, unsigned mask)
{
    BlockState *bs = g_bs;
    noteYield(bs);
    const unsigned t = bs->fibres[bs->current].tc.tid.x;
    BlockState::Warp &w = bs->warps[t >> 5];
    const unsigned lanesHere = std::min<unsigned>(32u, bs->nThreads - (t & ~31u));
    int participants = 0;
    for (unsigned l = 0; l < lanesHere; ++l) participants += (mask >> l) & 1u;
    if (size > 16) { fprintf(stderr, "emu: warp exchange of %zu bytes\n", size); abort(); }
    memcpy(w.slot[t & 31u], mine, size);
    warpBarrier(bs, w, participants);
    if (srcLane < lanesHere) memcpy(theirs, w.slot[srcLane], size);
    warpBarrier(bs, w, participants);
}

unsigned long long nowNs()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

void runGrid(const Cfg &cfg, void (*tramp)(void *), void *closure, const void *kernelId)
{
    if (g_bs) { fprintf(stderr, "emu: nested launch\n"); abort(); }
    const unsigned nThreads = cfg.b.x * cfg.b.y * cfg.b.z;
    if (nThreads == 0 || nThreads > 1024 || cfg.g.x == 0) { fprintf(stderr, "emu: invalid launch configuration (%u blocks of %u)\n", cfg.g.x, nThreads); abort(); }
    static thread_local BlockState bs;
    if (bs.fibres.size() < nThreads) {
        const size_t old = bs.fibres.size();
        bs.fibres.resize(nThreads);
        for (size_t k = old; k < nThreads; ++k) bs.fibres[k].stack = (char *)malloc(kStack);
    }
    std::vector<Fibre> &fs = bs.fibres;
    bs.tramp = tramp;
    bs.closure = closure;
    g_bs = &bs;
    bs.nThreads = nThreads;
    bs.warps.assign((nThreads + 31) / 32, BlockState::Warp());
    bs.yielded = false;
    bs.direct = false;
    bool known = false, yields = true;
    {
        std::lock_guard<std::mutex> lock(g_kernelMu);
        auto it = g_kernelYields.find(kernelId);
        if (it != g_kernelYields.end()) { known = true; yields = it->second; }
    }
    bool firstBlock = true;
    for (unsigned bz = 0; bz < cfg.g.z; ++bz)
        for (unsigned by = 0; by < cfg.g.y; ++by)
            for (unsigned bx = 0; bx < cfg.g.x; ++bx) {
                if (!firstBlock && !known) {
                    std::lock_guard<std::mutex> lock(g_kernelMu);
                    g_kernelYields[kernelId] = bs.yielded;
                    known = true;
                    yields = bs.yielded;
                }
                firstBlock = false;
                if (known && !yields) {
                    bs.direct = true;
                    bs.directCtx.bid = uint3{bx, by, bz};
                    bs.directCtx.bdim = cfg.b;
                    bs.directCtx.gdim = cfg.g;
                    for (unsigned t = 0; t < nThreads; ++t) {
                        bs.directCtx.tid = uint3{t % cfg.b.x, (t / cfg.b.x) % cfg.b.y, t / (cfg.b.x * cfg.b.y)};
                        tramp(closure);
                    }
                    bs.direct = false;
                    continue;
                }
                bs.alive = (int)nThreads;
                bs.arrived = 0;
                for (auto &w : bs.warps) w.arrived = 0;
                for (unsigned t = 0; t < nThreads; ++t) {
                    Fibre &f = fs[t];
                    f.done = false;
                    f.tc.tid = uint3{t % cfg.b.x, (t / cfg.b.x) % cfg.b.y, t / (cfg.b.x * cfg.b.y)};
                    f.tc.bid = uint3{bx, by, bz};
                    f.tc.bdim = cfg.b;
                    f.tc.gdim = cfg.g;
                    getcontext(&f.uc);
                    f.uc.uc_stack.ss_sp = f.stack;
                    f.uc.uc_stack.ss_size = kStack;
                    f.uc.uc_link = nullptr;
                    makecontext(&f.uc, fibreMain, 0);
                }
                int remaining = (int)nThreads;
                while (remaining > 0) {
                    int progressed = 0;
                    for (unsigned t = 0; t < nThreads; ++t) {
                        if (fs[t].done) continue;
                        bs.current = (int)t;
                        toFibre(&bs, fs[t]);
                        bs.current = -1;
                        if (fs[t].done) { --remaining; ++progressed; }
                    }
                    (void)progressed;
                }
            }
    if (!known) {
        std::lock_guard<std::mutex> lock(g_kernelMu);
        g_kernelYields[kernelId] = bs.yielded;
    }
    g_bs = nullptr;
}
} // namespace emu

// ------------------------------------------------------------------------------------------------------------------
namespace {
void drainAll();
constexpr size_t kZone = 256;
constexpr unsigned char kFill = 0xA5;
std::mutex g_mu;
std::map<void *, size_t> g_allocs;
int g_lastError = 0;

struct Checker {
    ~Checker()
    {
        for (auto &a : g_allocs) {
            const unsigned char *lo = (const unsigned char *)a.first - kZone, *hi = (const unsigned char *)a.first + a.second;
            for (size_t k = 0; k < kZone; ++k)
                if (lo[k] != kFill || hi[k] != kFill) { fprintf(stderr, "emu: red zone of allocation %p (%zu bytes) overwritten\n", a.first, a.second); break; }
        }
    }
} g_checker;

bool zonesIntact(void *p, size_t bytes)
{
    const unsigned char *lo = (const unsigned char *)p - kZone, *hi = (const unsigned char *)p + bytes;
    for (size_t k = 0; k < kZone; ++k)
        if (lo[k] != kFill || hi[k] != kFill) return false;
    return true;
}
} // namespace

struct EmuStream {
    std::mutex m;
    std::condition_variable work, idle;
    std::deque<std::function<void()>> q;
    bool busy = false, stop = false;
    std::thread worker;
    void loop()
    {
        std::unique_lock<std::mutex> lock(m);
        for (;;) {
            work.wait(lock, [&] { return stop || !q.empty(); });
            if (q.empty()) return;
            std::function<void()> op = std::move(q.front());
            q.pop_front();
            busy = true;
            lock.unlock();
            op();
            lock.lock();
            busy = false;
            if (q.empty()) idle.notify_all();
        }
    }
    void drain()
    {
        std::unique_lock<std::mutex> lock(m);
        idle.wait(lock, [&] { return q.empty() && !busy; });
    }
};
struct EmuEvent {
    std::mutex m;
    std::condition_variable cv;
    unsigned long long issued = 0, completed = 0, t = 0;
};
namespace {
std::mutex g_streamsMu;
std::set<EmuStream *> g_streams;
void drainAll()
{
    std::vector<EmuStream *> all;
    {
        std::lock_guard<std::mutex> lock(g_streamsMu);
        all.assign(g_streams.begin(), g_streams.end());
    }
    for (EmuStream *s : all) s->drain();
}
cudaError_t newStream(cudaStream_t *out)
{
    EmuStream *s = new EmuStream;
    s->worker = std::thread([s] { s->loop(); });
    std::lock_guard<std::mutex> lock(g_streamsMu);
    g_streams.insert(s);
    *out = s;
    return cudaSuccess;
}
} // namespace
namespace emu {
void enqueue(cudaStream_t stream, std::function<void()> op)
{
    if (!stream) { op(); return; }
    std::lock_guard<std::mutex> lock(stream->m);
    stream->q.push_back(std::move(op));
    stream->work.notify_one();
}
} // namespace emu

// N processes as N ranks (torchrun + gloo): with CHIMP_EMU_ARENA=<file in /dev/shm> every process maps the same file at
// the same address and cudaMalloc hands out pieces of it, so "device" pointers -- and with them the CUDA IPC handles,
// which are the pointers themselves here -- mean the same memory in every rank, like peer-mapped GPU memory does.
struct Arena {
    std::atomic<size_t> offset;
};
Arena *g_arena = nullptr;
constexpr size_t kArenaBytes = 48ull << 30;
Arena *arena()
{
    static bool tried = false;
    if (tried) return g_arena;
    tried = true;
    const char *path = getenv("CHIMP_EMU_ARENA");
    if (!path || !*path) return nullptr;
    const int fd = open(path, O_RDWR | O_CREAT, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)kArenaBytes) != 0) { perror("emu arena"); abort(); }
    void *base = mmap((void *)0x6f0000000000ull, kArenaBytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED_NOREPLACE | MAP_NORESERVE, fd, 0);
    close(fd);
    if (base != (void *)0x6f0000000000ull) { perror("emu arena mmap"); abort(); }
    g_arena = (Arena *)base;
    size_t zero = 0;
    g_arena->offset.compare_exchange_strong(zero, 4096); // the first process leaves room for this header
    return g_arena;
}

extern "C" {
cudaError_t cudaMalloc(void **p, size_t bytes)
{
    const size_t padded = (bytes + 255) / 256 * 256;
    unsigned char *raw;
    if (Arena *a = arena()) {
        const size_t off = a->offset.fetch_add(padded + 2 * kZone);
        if (off + padded + 2 * kZone > kArenaBytes) return cudaErrorMemoryAllocation;
        raw = (unsigned char *)a + off;
    } else {
        raw = (unsigned char *)aligned_alloc(256, padded + 2 * kZone);
    }
    if (!raw) return cudaErrorMemoryAllocation;
    memset(raw, kFill, padded + 2 * kZone); // fresh device memory is not zero: poison it (0xA5A5... is a huge negative double / int)
    *p = raw + kZone;
    std::lock_guard<std::mutex> lock(g_mu);
    g_allocs[*p] = bytes;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    drainAll(); // cudaFree synchronises the device
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) { fprintf(stderr, "emu: cudaFree of unknown pointer %p\n", p); abort(); }
    if (!zonesIntact(p, it->second)) { fprintf(stderr, "emu: OUT-OF-BOUNDS WRITE next to allocation %p (%zu bytes)\n", p, it->second); abort(); }
    g_allocs.erase(it);
    if (!g_arena) free((unsigned char *)p - kZone); // arena memory is not reused
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) { *p = calloc(1, bytes); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void **dev, void *host, unsigned) { *dev = host; return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, int) { if (bytes) memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, int kind, cudaStream_t s)
{
    if (!bytes) return cudaSuccess;
    if (kind == cudaMemcpyHostToDevice) {
        // a pageable source is staged before the call returns: take the snapshot now, deliver it in stream order
        std::vector<unsigned char> staged((const unsigned char *)src, (const unsigned char *)src + bytes);
        emu::enqueue(s, [dst, staged]() { memcpy(dst, staged.data(), staged.size()); });
    } else
        emu::enqueue(s, [dst, src, bytes]() { memmove(dst, src, bytes); });
    return cudaSuccess;
}
cudaError_t cudaMemset(void *dst, int v, size_t bytes) { if (bytes) memset(dst, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *dst, int v, size_t bytes, cudaStream_t s) { if (bytes) emu::enqueue(s, [dst, v, bytes]() { memset(dst, v, bytes); }); return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d >= 0 ? cudaSuccess : cudaErrorInvalidValue; } // every rank's "device" is the host
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { const char *w = getenv("LOCAL_WORLD_SIZE"); *n = w ? atoi(w) : 1; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { drainAll(); return cudaSuccess; }
cudaError_t cudaGetLastError(void) { const int e = g_lastError; g_lastError = 0; return e; }
const char *cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return newStream(s); }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { return newStream(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s)
{
    s->drain();
    {
        std::lock_guard<std::mutex> lock(s->m);
        s->stop = true;
        s->work.notify_all();
    }
    s->worker.join();
    {
        std::lock_guard<std::mutex> lock(g_streamsMu);
        g_streams.erase(s);
    }
    delete s;
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) { if (s) s->drain(); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned)
{
    unsigned long long gen;
    {
        std::lock_guard<std::mutex> lock(e->m);
        gen = e->issued;
    }
    if (gen == 0) return cudaSuccess; // never recorded: no dependency
    emu::enqueue(s, [e, gen]() {
        std::unique_lock<std::mutex> lock(e->m);
        e->cv.wait(lock, [&] { return e->completed >= gen; });
    });
    return cudaSuccess;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new EmuEvent; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new EmuEvent; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; } // kept: operations that name it may still be queued
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s)
{
    unsigned long long gen;
    {
        std::lock_guard<std::mutex> lock(e->m);
        gen = ++e->issued;
    }
    emu::enqueue(s, [e, gen]() {
        std::lock_guard<std::mutex> lock(e->m);
        e->t = emu::nowNs();
        e->completed = gen;
        e->cv.notify_all();
    });
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e)
{
    std::unique_lock<std::mutex> lock(e->m);
    e->cv.wait(lock, [&] { return e->completed >= e->issued; });
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)((double)(b->t - a->t) * 1e-6); return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
}
