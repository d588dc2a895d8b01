// TEST INFRASTRUCTURE — never part of libodis_b200.so, never on the product path.
//
// A small host emulation of the CUDA execution model, enough to run this repository's kernel SOURCE on the CPU so that kernel
// logic (indexing, operation order, barriers, warp shuffles, last-block reductions, launch sequencing in the engine) can be
// checked in the `-m "not gpu"` suite, where no GPU exists. tests/simt/build_emu.py rewrites the launch syntax
// (`k<<<g, b, s, st>>>(args)` -> simt::launch), `extern __shared__` and the inline PTX of csrc/*.cu and compiles the result
// against this header (found as <cuda_runtime.h> through tests/simt/stub/) into tests/_build/libodis_b200_emu.so.
//
// Model: a stream is a worker thread with a FIFO of operations (launches, copies, memsets, event records / waits), so calls return
// before the work is done, streams run concurrently and a missing synchronisation shows. Within a launch the CTAs run one after
// another; the threads of a CTA are fibers (a 10-instruction x86-64 context switch) that run until they reach __syncthreads, a
// warp shuffle or an mbarrier wait, where they wait for their CTA / warp / barrier. Shared memory is `static thread_local`
// storage (one CTA is alive per stream at a time), global memory is the host heap, a captured graph is the list of its launches.
// Several "devices" are just several streams: kernels of different solvers run at the same time and talk through peer flags
// (system-scope acquire / release accesses become __atomic operations). What this does NOT model: memory consistency between
// the CTAs of one launch, timing, caches. It is a logic check, not a substitute for the B200 runs.
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <utility>
#include <vector>

#define ODIS_SIMT_EMULATION 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __align__(n) alignas(n)

using std::max;
using std::min;

struct double2 { double x, y; };
struct int2 { int x, y; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline double2 make_double2(double a, double b) { return double2{a, b}; }
inline int2 make_int2(int a, int b) { return int2{a, b}; }

// threadIdx / blockIdx / blockDim / gridDim live in the per-stream-thread state (simt::ThreadState, below) behind one thread-local
// pointer (built with -mtls-dialect=gnu2: TLS descriptors keep that access cheap in a dlopen'ed library)

// ---- CUDA runtime subset: streams are worker threads, memory is the host heap ----
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorUnknown = 999, cudaErrorPeerAccessAlreadyEnabled = 704;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };
enum cudaStreamCaptureMode { cudaStreamCaptureModeRelaxed };
struct cudaIpcMemHandle_t { char reserved[64]; };

namespace simt {
using Closure = std::function<void()>;
struct Graph { std::vector<Closure> launches; };
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace simt

struct simt_stream {
    std::thread worker;
    std::mutex m;
    std::condition_variable wake, idle;
    std::deque<simt::Closure> q;
    bool stop = false, busy = false;
    int device = 0;                          // the device that was current when the stream was created
    bool blocking = true;                    // false: created with cudaStreamNonBlocking, the legacy stream does not synchronise with it
    simt::Graph* capture = nullptr;          // non-null between cudaStreamBeginCapture and cudaStreamEndCapture (host side)
    void run() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            wake.wait(lk, [&] { return stop || !q.empty(); });
            if (q.empty()) return;
            simt::Closure op = std::move(q.front());
            q.pop_front();
            busy = true;
            lk.unlock();
            op();
            lk.lock();
            busy = false;
            if (q.empty()) idle.notify_all();
        }
    }
    void push(simt::Closure op) {
        { std::lock_guard<std::mutex> lk(m); q.push_back(std::move(op)); }
        wake.notify_one();
    }
    void drain() {
        std::unique_lock<std::mutex> lk(m);
        idle.wait(lk, [&] { return q.empty() && !busy; });
    }
};
struct simt_event {
    std::mutex m;
    std::condition_variable cv;
    unsigned long long recorded = 0, completed = 0;
    double t_ms = 0.0;
};
typedef simt_stream* cudaStream_t;
typedef simt_event* cudaEvent_t;
typedef simt::Graph* cudaGraph_t;
typedef simt::Graph* cudaGraphExec_t;

namespace simt {
inline std::mutex registry_mutex;
inline std::vector<simt_stream*> streams;
inline long long kernels_run = 0;
inline thread_local int current_device = 0;
// device-wide synchronisation (cudaFree, cudaDeviceSynchronize): every stream of the calling thread's current device
inline void drain_device() {
    std::vector<simt_stream*> copy;
    { std::lock_guard<std::mutex> lk(registry_mutex); copy = streams; }
    for (simt_stream* st : copy) if (st->device == current_device) st->drain();
}
// what the legacy default stream synchronises with: the BLOCKING streams of the current device only. Streams created with
// cudaStreamNonBlocking (all of the engine's) are not waited for, as on the hardware -- a cudaMemcpy that relied on it would show.
inline void drain_legacy() {
    std::vector<simt_stream*> copy;
    { std::lock_guard<std::mutex> lk(registry_mutex); copy = streams; }
    for (simt_stream* st : copy) if (st->device == current_device && st->blocking) st->drain();
}
// an operation issued to `st`: in order behind what was issued before
inline void issue(simt_stream* st, Closure op) {
    if (st == nullptr) { drain_legacy(); op(); return; }
    st->push(std::move(op));
}
}  // namespace simt

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime: unsupported call"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 8; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { simt::current_device = d; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = simt::current_device; return cudaSuccess; }
// a small "GPU": persistent kernels size their grids from this, so each emulated CTA works through many tiles (stage reuse and
// mbarrier parity wrap-around get exercised on small grids) and fewer idle fibers are created
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 6; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorUnknown; }
inline cudaError_t cudaFree(void* p) { simt::drain_device(); std::free(p); return cudaSuccess; }
// page-locked host blocks are remembered: an asynchronous copy into one really is asynchronous (the caller must wait for it)
namespace simt {
inline std::vector<std::pair<const unsigned char*, size_t>> pinned_blocks;
inline bool is_pinned(const void* p) {
    std::lock_guard<std::mutex> lk(registry_mutex);
    for (auto& b : pinned_blocks) if ((const unsigned char*)p >= b.first && (const unsigned char*)p < b.first + b.second) return true;
    return false;
}
}  // namespace simt
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
    if (cudaMalloc(p, n) != cudaSuccess) return cudaErrorUnknown;
    std::lock_guard<std::mutex> lk(simt::registry_mutex);
    simt::pinned_blocks.emplace_back((const unsigned char*)*p, n ? n : 1);
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void* p) {
    {
        std::lock_guard<std::mutex> lk(simt::registry_mutex);
        for (size_t i = 0; i < simt::pinned_blocks.size(); i++)
            if (simt::pinned_blocks[i].first == (const unsigned char*)p) { simt::pinned_blocks.erase(simt::pinned_blocks.begin() + (long)i); break; }
    }
    return cudaFree(p);
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { simt::drain_legacy(); std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st = nullptr) {
    if ((k == cudaMemcpyHostToDevice || k == cudaMemcpyHostToHost) && !simt::is_pinned(s)) {   // pageable source: staged at call time, as the driver does
        auto staged = std::make_shared<std::vector<unsigned char>>((const unsigned char*)s, (const unsigned char*)s + n);
        simt::issue(st, [d, staged]() { std::memcpy(d, staged->data(), staged->size()); });
    } else {
        simt::issue(st, [d, s, n]() { std::memmove(d, s, n); });
        if (k == cudaMemcpyDeviceToHost && st && !simt::is_pinned(d)) st->drain();   // pageable destination: the call returns when the data is there
    }
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t st = nullptr) {
    simt::issue(st, [=]() { for (size_t r = 0; r < h; r++) std::memmove((char*)d + r * dp, (const char*)s + r * sp, w); });
    if (st) st->drain();
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = nullptr) { simt::issue(st, [d, v, n]() { std::memset(d, v, n); }); return cudaSuccess; }
template <class T>
inline cudaError_t cudaMemcpyToSymbol(T& symbol, const void* src, size_t n) {
    simt::drain_legacy();
    // a __constant__ symbol has one instance per device in CUDA, one for all "devices" here: an identical upload by the next device is not
    // repeated (it would look like a write under a kernel of the first device that is reading the table)
    if (std::memcmp((const void*)&symbol, src, n) != 0) std::memcpy((void*)&symbol, src, n);
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* out) {
    simt_stream* st = new simt_stream();
    st->device = simt::current_device;
    st->worker = std::thread([st] { st->run(); });
    { std::lock_guard<std::mutex> lk(simt::registry_mutex); simt::streams.push_back(st); }
    *out = st;
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags) { cudaStreamCreate(s); (*s)->blocking = !(flags & cudaStreamNonBlocking); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t st) { if (st) st->drain(); else simt::drain_legacy(); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t st) {
    st->drain();
    { std::lock_guard<std::mutex> lk(simt::registry_mutex); simt::streams.erase(std::remove(simt::streams.begin(), simt::streams.end(), st), simt::streams.end()); }
    { std::lock_guard<std::mutex> lk(st->m); st->stop = true; }
    st->wake.notify_one();
    st->worker.join();
    delete st;
    return cudaSuccess;
}
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new simt_event(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { simt::drain_device(); delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st = nullptr) {
    unsigned long long ticket;
    { std::lock_guard<std::mutex> lk(e->m); ticket = ++e->recorded; }
    simt::issue(st, [e, ticket]() {
        { std::lock_guard<std::mutex> lk(e->m); e->completed = ticket; e->t_ms = simt::now_ms(); }
        e->cv.notify_all();
    });
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t e) {
    std::unique_lock<std::mutex> lk(e->m);
    const unsigned long long ticket = e->recorded;
    e->cv.wait(lk, [&] { return e->completed >= ticket; });
    return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t e, unsigned) {
    unsigned long long ticket;
    { std::lock_guard<std::mutex> lk(e->m); ticket = e->recorded; }
    simt::issue(st, [e, ticket]() { std::unique_lock<std::mutex> lk(e->m); e->cv.wait(lk, [&] { return e->completed >= ticket; }); });
    return cudaSuccess;
}
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    cudaEventSynchronize(a); cudaEventSynchronize(b);
    *ms = (float)std::max(b->t_ms - a->t_ms, 1e-6);
    return cudaSuccess;
}
// one process holds every "device": peers use each other's pointers directly (odis_halo_connect's same-process branch)
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof *h); std::memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode) { st->capture = new simt::Graph(); return cudaSuccess; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t* g) { *g = st->capture; st->capture = nullptr; return cudaSuccess; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) { *e = new simt::Graph(*g); return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { simt::drain_device(); delete g; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t g, cudaStream_t st) {
    simt::issue(st, [g]() { for (auto& c : g->launches) c(); });
    return cudaSuccess;
}

// ---- device intrinsics ----
inline int __double2hiint(double x) { long long b; std::memcpy(&b, &x, sizeof b); return (int)(b >> 32); }
inline int __double2loint(double x) { long long b; std::memcpy(&b, &x, sizeof b); return (int)(b & 0xffffffffll); }
inline double __hiloint2double(int hi, int lo) {
    const unsigned long long b = ((unsigned long long)(unsigned int)hi << 32) | (unsigned long long)(unsigned int)lo;
    double x; std::memcpy(&x, &b, sizeof x); return x;
}
inline double __drcp_rn(double x) { return 1.0 / x; }                   // rcp.rn.f64 is correctly rounded, as IEEE division
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
namespace simt { inline void slow_fence(); }
// ODIS_EMU_SLOW_FENCE=R: a system-scope fence takes R scheduler rounds, during which the other threads of the CTA run on (on the
// hardware it is an NVLink round trip, several microseconds: long enough for the rest of a small kernel to finish)
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); simt::slow_fence(); }
inline long long clock64() { return (long long)(simt::now_ms() * 1.0e6); }          // "cycles" = nanoseconds: the kernels' spin bounds stay seconds

namespace simt {

constexpr int kStackBytes = 256 * 1024;
enum State { RUNNABLE, AT_BARRIER, AT_SHUFFLE, AT_NAMED_BARRIER, DONE };
// x86-64 System V context switch (callee-saved registers + stack pointer), defined once in simt_switch.cpp (build_emu.py):
// saves the caller's context on its own stack, stores that stack pointer in *from_sp and resumes the context at to_sp.
extern "C" void simt_switch(void** from_sp, void* to_sp);
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    State state = DONE;
    int named_id = 0, named_count = 0;      // bar.sync id, count
    long long spins = 0;                    // consecutive unsuccessful polls of an mbarrier
};
struct WarpExchange { uint64_t pending[32], result[32]; };
struct Cta {
    std::vector<Fiber> fibers;
    std::vector<WarpExchange> warps;
    void* scheduler_sp = nullptr;
    int current = -1;
    const Closure* body = nullptr;
};
// everything a stream's worker thread needs while it runs kernels
struct ThreadState {
    uint3 tid{0, 0, 0}, bid{0, 0, 0};
    dim3 bdim, gdim;
    Cta cta_;
    unsigned char* dyn = nullptr;              // dynamic shared memory of the CTA in flight
};
inline thread_local ThreadState* self_ = nullptr;
inline void ensure_thread_state() {
    if (self_) return;
    self_ = new ThreadState();
    self_->dyn = (unsigned char*)std::aligned_alloc(128, 228 * 1024);
}
#define threadIdx (simt::self_->tid)
#define blockIdx (simt::self_->bid)
#define blockDim (simt::self_->bdim)
#define gridDim (simt::self_->gdim)
#define cta self_->cta_
inline unsigned char* dynamic_shared() { return self_->dyn; }

inline void fiber_entry() {
    (*cta.body)();
    Fiber& f = cta.fibers[(size_t)cta.current];
    f.state = DONE;
    simt_switch(&f.sp, cta.scheduler_sp);          // never resumed
    std::abort();
}
inline void yield(State why) {
    Cta& c = self_->cta_;
    Fiber& f = c.fibers[(size_t)c.current];
    f.state = why;
    simt_switch(&f.sp, c.scheduler_sp);
}
inline void slow_fence() {
    static const int rounds = [] { const char* e = std::getenv("ODIS_EMU_SLOW_FENCE"); return e ? std::atoi(e) : 0; }();
    if (rounds <= 0 || self_ == nullptr || self_->cta_.current < 0 || self_->cta_.body == nullptr) return;     // host code, or not inside a kernel
    for (int r = 0; r < rounds; r++) {
        yield(RUNNABLE);
        std::this_thread::yield();                  // ... and the CTAs on the other threads (ODIS_EMU_CTA_THREADS)
    }
}
// a fresh context that simt_switch can resume: six callee-saved register slots, the entry address its `ret` jumps to, and a
// null return address so that the entry function starts with the ABI's stack alignment
inline void prepare(Fiber& f) {
    uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
    void** slot = (void**)top;
    *--slot = nullptr;                              // fake return address of fiber_entry
    *--slot = (void*)&fiber_entry;
    for (int r = 0; r < 6; r++) *--slot = nullptr;  // rbp rbx r12 r13 r14 r15
    f.sp = (void*)slot;
}
inline void set_thread(int linear, const dim3& block) {
    threadIdx.x = (unsigned)linear % block.x;
    threadIdx.y = ((unsigned)linear / block.x) % block.y;
    threadIdx.z = (unsigned)linear / (block.x * block.y);
}

// one CTA: fibers run round-robin to their next barrier / shuffle / end
inline void run_cta(const dim3& block, const Closure& body) {
    ThreadState* const me = self_;             // one thread-local lookup; the fibers of this CTA share it
    Cta& c = me->cta_;
    const int n = (int)(block.x * block.y * block.z);
    if ((int)c.fibers.size() < n) {
        const size_t old = c.fibers.size();
        c.fibers.resize((size_t)n);
        for (size_t i = old; i < (size_t)n; i++) c.fibers[i].stack = (char*)std::malloc(kStackBytes);
    }
    c.warps.assign((size_t)(n + 31) / 32, WarpExchange{});
    c.body = &body;
    for (int i = 0; i < n; i++) {
        Fiber& f = c.fibers[(size_t)i];
        prepare(f);
        f.state = RUNNABLE;
        f.spins = 0;
    }
    int alive = n;
    while (alive > 0) {
        bool progressed = false;
        for (int i = 0; i < n; i++) {
            Fiber& f = c.fibers[(size_t)i];
            if (f.state != RUNNABLE) continue;
            c.current = i;
            me->tid.x = (unsigned)i % block.x; me->tid.y = ((unsigned)i / block.x) % block.y; me->tid.z = (unsigned)i / (block.x * block.y);
            simt_switch(&c.scheduler_sp, f.sp);
            progressed = true;
            if (f.state == DONE) alive--;
        }
        // warps whose live lanes all wait at a shuffle exchange their values
        for (int w = 0; w < (n + 31) / 32; w++) {
            int waiting = 0, live = 0;
            for (int l = 0; l < 32 && w * 32 + l < n; l++) {
                const State st = c.fibers[(size_t)(w * 32 + l)].state;
                live += st != DONE;
                waiting += st == AT_SHUFFLE;
            }
            if (live > 0 && waiting == live) {
                std::memcpy(c.warps[(size_t)w].result, c.warps[(size_t)w].pending, sizeof c.warps[(size_t)w].result);
                for (int l = 0; l < 32 && w * 32 + l < n; l++)
                    if (c.fibers[(size_t)(w * 32 + l)].state == AT_SHUFFLE) c.fibers[(size_t)(w * 32 + l)].state = RUNNABLE;
                progressed = true;
            }
        }
        // named barriers (bar.sync id, count): open when `count` threads wait on the id
        for (int id = 1; id < 16; id++) {
            int waiting = 0, need = 0;
            for (int i = 0; i < n; i++)
                if (c.fibers[(size_t)i].state == AT_NAMED_BARRIER && c.fibers[(size_t)i].named_id == id) { waiting++; need = c.fibers[(size_t)i].named_count; }
            if (waiting > 0 && waiting >= need) {
                for (int i = 0; i < n; i++)
                    if (c.fibers[(size_t)i].state == AT_NAMED_BARRIER && c.fibers[(size_t)i].named_id == id) c.fibers[(size_t)i].state = RUNNABLE;
                progressed = true;
            }
        }
        // the barrier opens when every live thread of the CTA has arrived
        int at_barrier = 0, live = 0;
        for (int i = 0; i < n; i++) {
            live += c.fibers[(size_t)i].state != DONE;
            at_barrier += c.fibers[(size_t)i].state == AT_BARRIER;
        }
        if (live > 0 && at_barrier == live) {
            for (int i = 0; i < n; i++)
                if (c.fibers[(size_t)i].state == AT_BARRIER) c.fibers[(size_t)i].state = RUNNABLE;
            progressed = true;
        }
        if (!progressed && alive > 0) {
            std::fprintf(stderr, "simt_emu: deadlock in a CTA (divergent barrier / shuffle)\n");
            std::abort();
        }
    }
    c.current = -1;
    c.body = nullptr;
}

inline void run_grid(dim3 grid, dim3 block, const Closure& body) {
    ensure_thread_state();
    ThreadState* const me = self_;
    me->gdim = grid;
    me->bdim = block;
    // ODIS_EMU_CTA_THREADS=K (K > 1): the CTAs of a launch run on K OS threads at once instead of one after another, so that the
    // thread sanitizer also sees accesses of DIFFERENT CTAs that no ticket / fence / atomic orders (the epoch the last CTA counted
    // while another CTA's halo warp still read it, round 2, would have been reported). CTAs are claimed in index order from an atomic
    // counter; a kernel that needs all its CTAs resident at once (grid barrier) can run here when K >= its grid.
    static const int cta_threads = [] { const char* e = std::getenv("ODIS_EMU_CTA_THREADS"); return e ? std::atoi(e) : 0; }();
    const unsigned total = grid.x * grid.y * grid.z;
    if (cta_threads > 1 && total > 1) {
        std::atomic<unsigned> next{0};
        const int dev = current_device;
        auto worker = [&]() {
            current_device = dev;
            ensure_thread_state();
            ThreadState* const w = self_;
            w->gdim = grid;
            w->bdim = block;
            // (with a thread per CTA every CTA is live at once, claimed one each: what a grid-wide barrier needs)
            for (unsigned i = next.fetch_add(1); i < total; i = (unsigned)cta_threads >= total ? total : next.fetch_add(1)) {
                w->bid = uint3{i % grid.x, (i / grid.x) % grid.y, i / (grid.x * grid.y)};
                run_cta(block, body);
            }
            if (w != me) {                                   // a helper thread of this launch: its fiber stacks go with it
                for (Fiber& f : w->cta_.fibers) std::free(f.stack);
                std::free(w->dyn);
                delete w;
                self_ = nullptr;
            }
        };
        std::vector<std::thread> pool;
        const int n = std::min<int>(cta_threads, (int)total);
        for (int k = 1; k < n; k++) pool.emplace_back(worker);
        worker();
        for (auto& t : pool) t.join();
        __atomic_fetch_add(&kernels_run, 1, __ATOMIC_RELAXED);
        return;
    }
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++)
            for (unsigned x = 0; x < grid.x; x++) {
                me->bid = uint3{x, y, z};
                run_cta(block, body);
            }
    __atomic_fetch_add(&kernels_run, 1, __ATOMIC_RELAXED);
}

// kernel<<<grid, block, smem, stream>>>(args...): arguments are copied at launch time, as CUDA does
template <class K, class... A>
inline void launch(dim3 grid, dim3 block, size_t /*smem*/, cudaStream_t st, K kernel, A&&... args) {
    auto bound = std::make_tuple(std::decay_t<A>(std::forward<A>(args))...);
    Closure run = [grid, block, kernel, bound]() mutable {
        Closure body = [&]() { std::apply([&](auto&... a) { kernel(a...); }, bound); };
        run_grid(grid, block, body);
    };
    if (st && st->capture) st->capture->launches.push_back(run);
    else issue(st, run);
}

template <class T>
inline T shuffle(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Cta& c = self_->cta_;
    const int linear = c.current, lane = linear & 31;
    WarpExchange& w = c.warps[(size_t)(linear >> 5)];
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.pending[lane] = bits;
    Fiber& f = c.fibers[(size_t)linear];
    f.state = AT_SHUFFLE;
    simt_switch(&f.sp, c.scheduler_sp);
    if (src_lane < 0 || src_lane > 31) src_lane = lane;
    T out;
    std::memcpy(&out, &w.result[src_lane], sizeof(T));
    return out;
}

// mma.sync.aligned.m8n8k4.row.col.f64: D[8x8] = A[8x4] B[4x8] + C. Fragment layout: lane holds A[lane/4][lane%4], B[lane%4][lane/4]
// and C/D[lane/4][2*(lane%4) + {0,1}]. Emulated with two warp-wide exchanges (every lane publishes its element, then reads the four
// it needs); the products are accumulated with fused multiply-adds in k order.
inline void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    const int lane = cta.current & 31, row = lane >> 2, col0 = (lane & 3) * 2;
    WarpExchange& w = cta.warps[(size_t)(cta.current >> 5)];
    double A[4], B0[4], B1[4];
    auto publish = [&](double v) { uint64_t bits; std::memcpy(&bits, &v, 8); w.pending[lane] = bits; yield(AT_SHUFFLE); };
    auto read = [&](int src) { double v; std::memcpy(&v, &w.result[src], 8); return v; };
    publish(a);
    for (int k = 0; k < 4; k++) A[k] = read(row * 4 + k);
    publish(b);
    for (int k = 0; k < 4; k++) { B0[k] = read(col0 * 4 + k); B1[k] = read((col0 + 1) * 4 + k); }
    for (int k = 0; k < 4; k++) { c0 = std::fma(A[k], B0[k], c0); c1 = std::fma(A[k], B1[k], c1); }
}

// ---- shared-window addresses, mbarrier, 1-D bulk copies (the staged kernels) ----
// Dynamic shared memory is one static buffer; a "shared address" is the offset into it.
inline uint32_t shared_address(const void* p) {
    const std::ptrdiff_t off = (const unsigned char*)p - dynamic_shared();
    if (off < 0 || off >= (std::ptrdiff_t)(228 * 1024)) { std::fprintf(stderr, "simt_emu: address is not in dynamic shared memory\n"); std::abort(); }
    return (uint32_t)off;
}
inline void* shared_pointer(uint32_t a) { return dynamic_shared() + a; }
// mbarrier object in its 64-bit shared-memory word: phase parity, arrivals still pending in this phase, the count they are reset
// to, and the transaction bytes still expected (may go negative when a copy completes before its expect_tx, as in hardware)
struct Mbarrier { uint32_t phase : 1; uint32_t pending : 15; uint32_t count : 15; int32_t tx; };
static_assert(sizeof(Mbarrier) == 8, "mbarrier fits its shared-memory word");
inline void mbar_check(Mbarrier* b) {
    if (b->pending == 0 && b->tx == 0) { b->phase ^= 1u; b->pending = b->count; }
}
inline void mbar_init(uint32_t a, uint32_t count) { Mbarrier* b = (Mbarrier*)shared_pointer(a); b->phase = 0; b->pending = count; b->count = count; b->tx = 0; }
inline void mbar_arrive(uint32_t a, uint32_t expect_bytes) {
    Mbarrier* b = (Mbarrier*)shared_pointer(a);
    if (b->pending == 0) { std::fprintf(stderr, "simt_emu: more arrivals than the mbarrier expects\n"); std::abort(); }
    b->tx += (int32_t)expect_bytes;
    b->pending -= 1;
    mbar_check(b);
}
inline void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    if ((bytes & 15u) || (dst & 15u) || ((uintptr_t)src & 15u)) { std::fprintf(stderr, "simt_emu: cp.async.bulk needs 16-byte alignment and size\n"); std::abort(); }
    std::memcpy(shared_pointer(dst), src, bytes);                  // completes at once
    Mbarrier* b = (Mbarrier*)shared_pointer(bar);
    b->tx -= (int32_t)bytes;
    mbar_check(b);
}
// mbarrier.try_wait.parity loop: the phase with the given parity has completed when the barrier has moved on to the other parity
inline void mbar_wait(uint32_t a, uint32_t parity) {
    const Mbarrier* b = (const Mbarrier*)shared_pointer(a);
    Fiber& f = cta.fibers[(size_t)cta.current];
    while (b->phase == (parity & 1u)) {
        if (++f.spins > 50000000) { std::fprintf(stderr, "simt_emu: mbarrier wait never satisfied (deadlock)\n"); std::abort(); }
        yield(RUNNABLE);
    }
    f.spins = 0;
}
inline void named_barrier(int id, int count) {
    Fiber& f = cta.fibers[(size_t)cta.current];
    f.named_id = id;
    f.named_count = count;
    yield(AT_NAMED_BARRIER);
}

}  // namespace simt

inline void __syncthreads() { simt::yield(simt::AT_BARRIER); }
inline void __syncwarp(unsigned = 0xffffffffu) { (void)simt::shuffle(0, simt::cta.current & 31); }       // the warp's live lanes meet here
inline size_t __cvta_generic_to_shared(const void* p) { return simt::shared_address(p); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int delta) { const int lane = simt::cta.current & 31; return simt::shuffle(v, lane + delta < 32 ? lane + delta : lane); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int mask) { return simt::shuffle(v, (simt::cta.current & 31) ^ mask); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return simt::shuffle(v, src & 31); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
