// TEST HARNESS (CPU): the handful of CUDA runtime calls and device intrinsics the product's launch code and
// one-thread-per-particle kernels use, restated for the host so that gen_lib_emu.py can build the WHOLE C ABI
// (capi_core.cu, noise.cu, pendulum.cu, rollout_reacher.cu, update.cu) as a host library on top of the block
// emulator.  "Device" memory is host memory, streams are synchronous, the special-function-unit approximations
// are libm.  Not a product path: nothing under mjmpc_b200/ knows about it.
#pragma once
#include "block_emu.h"
#include <cstdlib>

namespace emu {
inline cudaError_t malloc_(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t props_(cudaDeviceProp* p) {
    std::memset(p, 0, sizeof *p);
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    return cudaSuccess;
}
}  // namespace emu

#define cudaMalloc(p, n) emu::malloc_((void**)(p), (n))
#define cudaFree(p) (std::free(p), cudaSuccess)
#define cudaMemcpyAsync(dst, src, n, kind, s) (std::memcpy((dst), (src), (n)), (void)(s), cudaSuccess)
#define cudaMemcpyToSymbolAsync(sym, src, n, off, kind, s) (std::memcpy((char*)&(sym) + (off), (src), (n)), (void)(s), cudaSuccess)
#define cudaStreamSynchronize(s) ((void)(s), cudaSuccess)
#define cudaDeviceSynchronize() cudaSuccess
#define cudaFuncSetAttribute(...) cudaSuccess
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
#define cudaSetDevice(d) ((void)(d), cudaSuccess)
#define cudaGetDeviceProperties(p, d) emu::props_(p)
#define cudaDeviceGetAttribute(v, a, d) (*(v) = ((a) == cudaDevAttrMultiProcessorCount ? 148 : 1965000), cudaSuccess)
#define cudaGetErrorString(e) "emulated runtime"
// streams are synchronous and in order here: a side stream is just more work done at once
#define cudaStreamCreateWithFlags(ps, f) (*(ps) = (cudaStream_t)1, cudaSuccess)
#define cudaStreamCreateWithPriority(ps, f, p) (*(ps) = (cudaStream_t)1, cudaSuccess)
#define cudaDeviceGetStreamPriorityRange(lo, hi) (*(lo) = 0, *(hi) = -5, cudaSuccess)
#define cudaEventCreateWithFlags(pe, f) (*(pe) = (cudaEvent_t)1, cudaSuccess)
#define cudaEventRecord(e, s) ((void)(e), (void)(s), cudaSuccess)
#define cudaStreamWaitEvent(s, e, f) ((void)(s), (void)(e), cudaSuccess)

#undef __constant__
#define __constant__ static
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
#define __log2f(x) log2f(x)          // glibc declares these names for its own internals
#define __sinf(x) sinf(x)
#define __cosf(x) cosf(x)
