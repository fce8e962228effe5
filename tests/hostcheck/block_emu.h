// TEST HARNESS (CPU): a minimal block-level emulation of CUDA for the product's kernels and launch code.
// Every simulated CUDA thread of a block is a FIBER (ucontext) on the OS thread that runs the block; a small
// pool of OS threads runs the blocks of a launch side by side.  __syncthreads and the warp shuffles are
// cooperative barriers (a fiber that reaches one yields; the barrier opens when every live thread of the block /
// warp waits at it; exited threads drop out), __shared__ variables are thread_local (= one copy per running
// block), atomics on global memory are real atomics because blocks run concurrently.  A barrier that can never
// open (divergent __syncthreads) aborts with a message instead of hanging.  Enough for kernels whose warps call
// shuffles uniformly (every live lane of the warp together) -- which is what the product's kernels do;
// __match_any_sync degrades to "no peers" (the warp-aggregated histogram update becomes one atomic per lane:
// same result).  Not a product path.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <sys/mman.h>
#include <ucontext.h>
#include <cuda_runtime.h>          // dim3, cudaStream_t, error codes (host-side declarations only)

namespace emu {
struct Idx { unsigned x, y, z; };
enum { RUN = 0, AT_BLOCK = 1, AT_WARP = 2, DONE = 3 };
struct Fiber { ucontext_t ctx; Idx tid; int state; };
struct BlockCtx {
    ucontext_t sched;
    std::vector<Fiber> fib;
    std::vector<int> warp_live, warp_wait;
    std::vector<unsigned long long> wbuf;       // [warp][32] shuffle exchange
    int n = 0, cur = 0, live = 0, block_wait = 0;
    Idx block{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
    void* dyn = nullptr;
    void (*call)(void*) = nullptr;
    void* fobj = nullptr;
    char* stacks = nullptr;
    size_t stacks_bytes = 0;
};
constexpr size_t kStack = 256 * 1024;
inline thread_local BlockCtx* t_ctx = nullptr;

inline void release_if_complete(BlockCtx& c, int w) {
    if (c.live > 0 && c.block_wait == c.live) {
        for (auto& f : c.fib) if (f.state == AT_BLOCK) f.state = RUN;
        c.block_wait = 0;
    }
    if (w >= 0 && c.warp_live[w] > 0 && c.warp_wait[w] == c.warp_live[w]) {
        for (int l = 32 * w; l < 32 * w + 32 && l < c.n; l++) if (c.fib[l].state == AT_WARP) c.fib[l].state = RUN;
        c.warp_wait[w] = 0;
    }
}
inline void yield_at(int what) {
    BlockCtx& c = *t_ctx;
    Fiber& f = c.fib[c.cur];
    f.state = what;
    const int w = c.cur >> 5;
    if (what == AT_BLOCK) c.block_wait++; else c.warp_wait[w]++;
    release_if_complete(c, w);
    if (f.state == RUN) return;                 // the barrier opened with this arrival
    swapcontext(&f.ctx, &c.sched);
}
inline void fiber_main() {
    BlockCtx& c = *t_ctx;
    c.call(c.fobj);
    Fiber& f = c.fib[c.cur];
    f.state = DONE;
    c.live--;
    c.warp_live[c.cur >> 5]--;
    release_if_complete(c, c.cur >> 5);         // exited threads no longer take part
    swapcontext(&f.ctx, &c.sched);
}
inline void run_block(BlockCtx& c, dim3 block, unsigned bx, unsigned by, unsigned bz, size_t smem) {
    const int n = (int)(block.x * block.y * block.z), nw = (n + 31) / 32;
    c.n = n; c.live = n; c.block_wait = 0; c.block = {bx, by, bz};
    if ((int)c.fib.size() < n) c.fib.resize(n);
    c.warp_live.assign(nw, 0); c.warp_wait.assign(nw, 0); c.wbuf.assign((size_t)nw * 32, 0);
    if (c.stacks_bytes < (size_t)n * kStack) {
        if (c.stacks) munmap(c.stacks, c.stacks_bytes);
        c.stacks_bytes = (size_t)n * kStack;
        c.stacks = (char*)mmap(nullptr, c.stacks_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (c.stacks == MAP_FAILED) { std::fprintf(stderr, "block_emu: cannot map fiber stacks\n"); std::abort(); }
    }
    std::vector<unsigned long long> dyn(smem / 8 + 2);
    c.dyn = dyn.data();
    for (int t = 0; t < n; t++) {
        Fiber& f = c.fib[t];
        f.tid = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
        f.state = RUN;
        c.warp_live[t >> 5]++;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = c.stacks + (size_t)t * kStack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())fiber_main, 0);
    }
    // Fibers only switch at barriers, so the ORDER in which runnable fibers are resumed decides what a missing
    // __syncthreads gets to see.  MJB_EMU_ORDER=reverse | shuffle (default: ascending thread index) lets the
    // tests run the same kernels under adversarial orders: code that is correctly synchronised cannot tell.
    static const int order_mode = [] { const char* e = std::getenv("MJB_EMU_ORDER"); return !e ? 0 : (e[0] == 'r' ? 1 : (e[0] == 's' ? 2 : 0)); }();
    std::vector<int> order(n);
    for (int t = 0; t < n; t++) order[t] = order_mode == 1 ? n - 1 - t : t;
    unsigned long long lcg = 0x9E3779B97F4A7C15ull ^ ((unsigned long long)bx * 0x100000001B3ull + by * 7919ull + bz);
    unsigned long long pass = 0;
    while (c.live > 0) {
        bool progressed = false;
        if (order_mode == 2) {                       // a fresh permutation for every scheduling pass
            for (int i = n - 1; i > 0; i--) {
                lcg = lcg * 6364136223846793005ull + 1442695040888963407ull + pass;
                std::swap(order[i], order[(int)((lcg >> 33) % (unsigned long long)(i + 1))]);
            }
            pass++;
        }
        for (int oi = 0; oi < n; oi++) {
            const int t = order[oi];
            if (c.fib[t].state != RUN) continue;
            c.cur = t;
            progressed = true;
            swapcontext(&c.sched, &c.fib[t].ctx);
        }
        if (!progressed) {
            std::fprintf(stderr, "block_emu: deadlock in block (%u,%u,%u): %d live threads, %d at __syncthreads -- "
                                 "a barrier or warp shuffle is not reached by every live thread\n", bx, by, bz, c.live, c.block_wait);
            std::abort();
        }
    }
}

inline std::atomic<unsigned long long> g_launches{0};      // kernels launched so far (tests count launches per step)
template <class F> void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, F f) {
    g_launches.fetch_add(1);
    const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
    std::atomic<unsigned long long> next{0};
    auto worker = [&] {
        BlockCtx c;
        c.call = [](void* p) { (*(F*)p)(); };
        c.fobj = (void*)&f;
        c.bdim = {block.x, block.y, block.z}; c.gdim = {grid.x, grid.y, grid.z};    // per worker: launches from several host threads may overlap
        t_ctx = &c;
        for (;;) {
            const unsigned long long b = next.fetch_add(1);
            if (b >= nblocks) break;
            run_block(c, block, (unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y)), smem);
        }
        if (c.stacks) munmap(c.stacks, c.stacks_bytes);
        t_ctx = nullptr;
    };
    static const unsigned hw = [] { const char* e = std::getenv("MJB_EMU_THREADS"); unsigned n = e ? (unsigned)std::atoi(e) : std::thread::hardware_concurrency(); return n ? n : 1u; }();
    const unsigned nthreads = (unsigned)std::min<unsigned long long>(nblocks, hw);
    if (nthreads <= 1) {
        BlockCtx* outer = t_ctx;
        worker();
        t_ctx = outer;
        return;
    }
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (unsigned i = 0; i < nthreads; i++) th.emplace_back(worker);
    for (auto& x : th) x.join();
}
inline void* dyn_smem() { return t_ctx->dyn; }
inline const Idx& thread_idx() { return t_ctx->fib[t_ctx->cur].tid; }
inline const Idx& block_idx() { return t_ctx->block; }
inline unsigned linear_tid() { return (unsigned)t_ctx->cur; }
template <class T, class S> T shfl(T v, S src_of_lane) {
    BlockCtx& c = *t_ctx;
    const int w = c.cur >> 5, lane = c.cur & 31;
    const int lanes = std::min(32, c.n - 32 * w);
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof v);
    c.wbuf[(size_t)w * 32 + lane] = raw;
    yield_at(AT_WARP);
    const int s = src_of_lane(lane);
    T r = v;
    if (s >= 0 && s < lanes) { raw = t_ctx->wbuf[(size_t)w * 32 + s]; std::memcpy(&r, &raw, sizeof r); }
    yield_at(AT_WARP);
    return r;
}
}  // namespace emu

#define g_dyn_smem dyn_smem()          // emu::g_dyn_smem in generated code
#define threadIdx emu::thread_idx()
#define blockIdx emu::block_idx()
#define blockDim emu::t_ctx->bdim
#define gridDim emu::t_ctx->gdim
#undef __shared__
#define __shared__ static thread_local
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
static inline void __syncthreads() { emu::yield_at(emu::AT_BLOCK); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int o) { return emu::shfl(v, [o](int l) { return l + o; }); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) { return emu::shfl(v, [o](int l) { return l ^ o; }); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int o) { return emu::shfl(v, [o](int l) { return l - o < 0 ? l : l - o; }); }
static inline unsigned __activemask() { return 1u << (emu::linear_tid() & 31); }
template <class T> static inline unsigned __match_any_sync(unsigned, T) { return 1u << (emu::linear_tid() & 31); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline long long __double_as_longlong(double x) { long long r; std::memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; std::memcpy(&r, &x, 8); return r; }
static inline double __dadd_rn(double a, double b) { return a + b; }     // compile with -ffp-contract=off
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) {}
template <class T> static inline T __ldg(const T* p) { return *p; }
#define cudaGetLastError() cudaSuccess
#define cudaFuncSetAttribute(...) cudaSuccess
