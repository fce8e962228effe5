// Development aid: cost of instruction supply on sm_100a for LONE warps running long straight-line loop bodies --
// the regime of the rollout kernel at <= 16 k particles per GPU.  Each warp runs a loop whose body is NI independent-
// chain FFMA instructions (8 chains: issue-bound at 1 instruction per cycle when instructions are there);
// W warps per block run either the SAME body or W DIFFERENT bodies (role-split style).  Prints cycles per
// instruction for body sizes from 4 KB to 96 KB.        nvcc -arch=sm_100a -O3 -o icache_probe icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, int SALT>
__device__ __forceinline__ void body(float (&x)[8], float a, float b) {
#pragma unroll
    for (int i = 0; i < NI / 8; i++) {
#pragma unroll
        for (int c = 0; c < 8; c++) x[c] = fmaf(x[c], a, b + (float)(SALT * 1000 + i));   // distinct immediates: no code folding
    }
}

template <int NI, int W, bool DIFF>
__global__ void __launch_bounds__(W * 32, 1) probe(float* out, long long* cyc, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int c = 0; c < 8; c++) x[c] = (float)(threadIdx.x + c);
    const int w = threadIdx.x >> 5;
    long long t0 = 0;
    for (int it = -2; it < iters; it++) {
        if (it == 0) { __syncthreads(); t0 = clock64(); }
        if (DIFF) {
            if (w == 0) body<NI, 0>(x, a, b);
            else if (w == 1) body<NI, 1>(x, a, b);
            else if (w == 2) body<NI, 2>(x, a, b);
            else body<NI, 3>(x, a, b);
        } else {
            body<NI, 0>(x, a, b);
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * W + w] = t1 - t0;
}

template <int NI, int W, bool DIFF>
void run(int blocks, float* out, long long* cyc) {
    const int iters = 200;
    probe<NI, W, DIFF><<<blocks, W * 32>>>(out, cyc, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    probe<NI, W, DIFF><<<blocks, W * 32>>>(out, cyc, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks * W, cudaMemcpyDeviceToHost);
    double mx = 0, mean = 0;
    for (int i = 0; i < blocks * W; i++) { mean += h[i]; if (h[i] > mx) mx = h[i]; }
    mean /= blocks * W;
    printf("body %5d instr (%3d KB per stream)  warps/SM %d  %s  blocks %3d : %.2f cycles/instr mean, %.2f max\n", NI, NI * 16 / 1024, W,
           DIFF ? "DIFFERENT bodies" : "same body       ", blocks, mean / (200.0 * NI), mx / (200.0 * NI));
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 128);
    cudaMalloc(&cyc, sizeof(long long) * 148 * 4);
    for (int blocks : {1, 148}) {
#define ROW(NI) run<NI, 1, false>(blocks, out, cyc); run<NI, 2, false>(blocks, out, cyc); run<NI, 4, false>(blocks, out, cyc); \
                run<NI, 2, true>(blocks, out, cyc); run<NI, 4, true>(blocks, out, cyc);
        ROW(256) ROW(512) ROW(1024) ROW(1536) ROW(2048) ROW(2560) ROW(3072) ROW(4096) ROW(6144)
    }
    return 0;
}
