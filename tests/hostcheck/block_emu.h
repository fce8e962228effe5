// TEST HARNESS (CPU): a minimal block-level emulation of CUDA for the product's reduction kernels (update.cu).
// Every simulated CUDA thread of a block is an OS thread; blocks run one after another.  __syncthreads is a
// std::barrier that exiting threads drop out of, warp shuffles exchange through a per-warp buffer behind a
// per-warp barrier, atomics are real atomics.  Enough for kernels whose warps call shuffles uniformly (every
// live lane of the warp together) -- which is what update.cu does; __match_any_sync degrades to "no peers"
// (the warp-aggregated histogram update becomes one atomic per lane: same result).  Not a product path.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
#include <cuda_runtime.h>          // dim3, cudaStream_t, error codes (host-side declarations only)

namespace emu {
struct Idx { unsigned x, y, z; };
struct Warp { std::unique_ptr<std::barrier<>> bar; unsigned long long buf[32]; int lanes; };
struct Block { std::unique_ptr<std::barrier<>> bar; std::vector<Warp> warps; };
inline thread_local Idx t_thread;
inline thread_local Block* t_block;
inline Idx g_block, g_bdim, g_gdim;
inline void* g_dyn_smem;

template <class F> void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, F f) {
    const unsigned n = block.x * block.y * block.z;
    std::vector<unsigned long long> dyn(smem / 8 + 2);
    g_bdim = {block.x, block.y, block.z}; g_gdim = {grid.x, grid.y, grid.z};
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        g_block = {bx, by, bz}; g_dyn_smem = dyn.data();
        Block ctx;
        ctx.bar = std::make_unique<std::barrier<>>(n);
        ctx.warps.resize((n + 31) / 32);
        for (unsigned w = 0; w < ctx.warps.size(); w++) {
            ctx.warps[w].lanes = (int)std::min(32u, n - 32 * w);
            ctx.warps[w].bar = std::make_unique<std::barrier<>>(ctx.warps[w].lanes);
        }
        std::vector<std::thread> th;
        th.reserve(n);
        for (unsigned t = 0; t < n; t++)
            th.emplace_back([&ctx, &f, t, block] {
                t_thread = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                t_block = &ctx;
                f();
                ctx.warps[t >> 5].bar->arrive_and_drop();       // exited lanes no longer take part
                ctx.bar->arrive_and_drop();
            });
        for (auto& x : th) x.join();
    }
}
inline unsigned linear_tid() { return t_thread.x + g_bdim.x * (t_thread.y + g_bdim.y * t_thread.z); }
template <class T, class S> T shfl(T v, S src_of_lane) {
    Warp& w = t_block->warps[linear_tid() >> 5];
    const int lane = (int)(linear_tid() & 31);
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof v);
    w.buf[lane] = raw;
    w.bar->arrive_and_wait();
    const int s = src_of_lane(lane);
    T r = v;
    if (s >= 0 && s < w.lanes) { raw = w.buf[s]; std::memcpy(&r, &raw, sizeof r); }
    w.bar->arrive_and_wait();
    return r;
}
}  // namespace emu

#define threadIdx emu::t_thread
#define blockIdx emu::g_block
#define blockDim emu::g_bdim
#define gridDim emu::g_gdim
#undef __shared__
#define __shared__ static
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
static inline void __syncthreads() { emu::t_block->bar->arrive_and_wait(); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int o) { return emu::shfl(v, [o](int l) { return l + o; }); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) { return emu::shfl(v, [o](int l) { return l ^ o; }); }
static inline unsigned __activemask() { return 1u << (emu::linear_tid() & 31); }
template <class T> static inline unsigned __match_any_sync(unsigned, T) { return 1u << (emu::linear_tid() & 31); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
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
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
template <class T> static inline T __ldg(const T* p) { return *p; }
#define cudaGetLastError() cudaSuccess
#define cudaFuncSetAttribute(...) cudaSuccess
