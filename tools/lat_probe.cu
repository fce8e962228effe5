// Development aid: dependent-issue latencies (cycles) of the instructions on K1's critical path, one warp.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_variants/lat_probe tools/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP> __global__ void probe(double* out, long long* cyc, double a, double b) {
    double x = a + threadIdx.x * 1e-9, y = b, x2 = a * 0.5, y2 = b * 0.5, x3 = a * 0.25, y3 = b * 0.25, x4 = a * 0.125, y4 = b * 0.125;
    __shared__ double sh[64];
    sh[threadIdx.x] = 0.0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = fma(x, y, a);
        else if (OP == 1) x = x * y;
        else if (OP == 2) x = x + y;
        else if (OP == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
        else if (OP == 4) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); double e = fma(-x, r, 1.0); double t = fma(e, e, e); x = fma(r, t, r); }
        else if (OP == 5) { sh[threadIdx.x] = x; x = sh[threadIdx.x] + y; }       // STS -> LDS -> DADD round trip
        else if (OP == 6) x = x < y ? a : x;                                       // DSETP + FSEL pair
        else if (OP == 7) { x = fma(x, y, a); y = fma(y, b, a); }                  // two independent chains
        else if (OP == 9) { x = fma(x, y, a); y = fma(y, b, a); x2 = fma(x2, b, a); y2 = fma(y2, b, a); }
        else if (OP == 10) { x = fma(x, y, a); y = fma(y, b, a); x2 = fma(x2, b, a); y2 = fma(y2, b, a);
                             x3 = fma(x3, b, a); y3 = fma(y3, b, a); x4 = fma(x4, b, a); y4 = fma(y4, b, a); }
        else if (OP == 8) { int hi = __double2hiint(x) ^ 0x80000000; x = __hiloint2double(hi, __double2loint(x)); x = x + y; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y + x2 + y2 + x3 + y3 + x4 + y4;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 1024); cudaMalloc(&c, 8);
    const char* names[] = {"DFMA", "DMUL", "DADD", "MUFU.RCP64H", "rcp_pos (MUFU+3 DFMA)", "STS+LDS+DADD", "DSETP+FSEL", "2 indep DFMA chains (per pair)", "LOP3 sign flip + DADD", "4 indep DFMA chains (per 4)", "8 indep DFMA chains (per 8)"};
#define RUN(OP) { probe<OP><<<1, 32>>>(d, c, 1.0000001, 0.9999999); probe<OP><<<1, 32>>>(d, c, 1.0000001, 0.9999999); long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-32s %.2f cycles per dependent op\n", names[OP], (double)h / N); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
    // issue rate: 8 independent DFMA chains in one warp, and with 2 / 4 / 8 warps on one SM sub-partition set
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
