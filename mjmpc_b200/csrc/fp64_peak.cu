// FP64 FMA peak microbenchmark: the roofline denominator for the rollout kernel (K1).
// MEASURED_PEAKS.json only carries HBM and bf16 numbers, so the FP64 CUDA-core peak is
// measured here: every thread runs 8 independent dependent-FMA chains, enough warps per SM to
// cover the DFMA latency, timed with CUDA events.
#include "common.h"

namespace mjb {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace mjb

extern "C" int mjb_fp64_peak(int device, int blocks_per_sm, int iters, double* tflops_out, double* ms_out) {
    MJB_REQUIRE(tflops_out, "mjb_fp64_peak: null output");
    MJB_CUDA(cudaSetDevice(device));
    int sms = 0;
    MJB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (blocks_per_sm < 1) blocks_per_sm = 4;
    if (iters < 1) iters = 4096;
    const int grid = sms * blocks_per_sm, block = 256;
    double* d = nullptr;
    MJB_CUDA(cudaMalloc(&d, sizeof(double) * grid * block));
    cudaEvent_t e0, e1;
    MJB_CUDA(cudaEventCreate(&e0));
    MJB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        MJB_CUDA(cudaEventRecord(e0));
        mjb::fp64_peak_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-6);
        MJB_CUDA(cudaEventRecord(e1));
        MJB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MJB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    MJB_CUDA(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    const double flops = 2.0 * 8 * 16 * (double)iters * grid * block;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return MJB_OK;
}
