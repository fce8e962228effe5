// Error slot, device query and the model handle of the C ABI (include/mjmpc_b200.h).
#include <stdlib.h>
#include <string.h>
#include "common.h"

namespace mjb {
static thread_local char g_err[512] = "";
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace mjb

extern "C" const char* mjb_last_error(void) { return mjb::g_err; }
extern "C" int mjb_version(void) { return 100; }

extern "C" int mjb_device_info(int device, int* sm_count, int* clock_khz, int* cc_major, int* cc_minor) {
    cudaDeviceProp p;
    MJB_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    int khz = 0;
    MJB_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    if (clock_khz) *clock_khz = khz;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    MJB_REQUIRE(p.major == 10, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
    return MJB_OK;
}

static unsigned long long g_serial = 1;

extern "C" int mjb_model_create(const double* host_params, int n_instances, int device, mjb_model** out) {
    MJB_REQUIRE(host_params && out, "mjb_model_create: null pointer");
    MJB_REQUIRE(n_instances >= 1, "mjb_model_create: n_instances must be >= 1");
    MJB_CUDA(cudaSetDevice(device));
    mjb_model* m = (mjb_model*)calloc(1, sizeof(mjb_model));
    m->device = device;
    m->n_instances = n_instances;
    m->h_params = (double*)malloc(sizeof(double) * CH_NDEV * n_instances);
    cudaError_t e = cudaMalloc(&m->d_params, sizeof(double) * CH_NDEV * n_instances);
    if (e != cudaSuccess) {
        free(m->h_params); free(m);
        return mjb::set_error(MJB_ECUDA, "cudaMalloc(model) failed: %s", cudaGetErrorString(e));
    }
    *out = m;
    int rc = mjb_model_update(m, 0, n_instances, host_params, nullptr);
    if (rc != MJB_OK) { mjb_model_destroy(m); *out = nullptr; }
    return rc;
}

extern "C" int mjb_model_update(mjb_model* m, int first, int n, const double* host_params, void* stream) {
    MJB_REQUIRE(m && host_params, "mjb_model_update: null pointer");
    MJB_REQUIRE(first >= 0 && n >= 1 && first + n <= m->n_instances, "mjb_model_update: instance range out of bounds");
    for (int i = 0; i < n; i++) {
        double* P = m->h_params + (size_t)(first + i) * CH_NDEV;
        memcpy(P, host_params + (size_t)i * CH_NPARAM, sizeof(double) * CH_NPARAM);
        MJB_REQUIRE(P[CS_TIMESTEP] > 0.0, "model instance %d: timestep must be positive", first + i);
        MJB_REQUIRE(P[CS_FRAME_SKIP] >= 1.0, "model instance %d: frame_skip must be >= 1", first + i);
        for (int l = 0; l < MJB_NJ; l++) MJB_REQUIRE(P[CH_MASS + l] > 0.0, "model instance %d: link %d has no mass", first + i, l);
        mjb_derive_params(P);
    }
    m->fits_sawyer = 1;
    m->uniform_frame_skip = 1;
    for (int i = 0; i < m->n_instances; i++) {
        m->fits_sawyer &= mjb_params_fit_sawyer(m->h_params + (size_t)i * CH_NDEV);
        m->uniform_frame_skip &= m->h_params[(size_t)i * CH_NDEV + CS_FRAME_SKIP] == m->h_params[CS_FRAME_SKIP];
    }
    MJB_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = (cudaStream_t)stream;
    MJB_CUDA(cudaMemcpyAsync(m->d_params + (size_t)first * CH_NDEV, m->h_params + (size_t)first * CH_NDEV,
                             sizeof(double) * CH_NDEV * n, cudaMemcpyHostToDevice, s));
    MJB_CUDA(cudaStreamSynchronize(s));
    m->serial = ++g_serial;
    return mjb::const_bank_on_update(m);
}

extern "C" int mjb_model_n_instances(const mjb_model* m) { return m ? m->n_instances : 0; }

extern "C" int mjb_model_destroy(mjb_model* m) {
    if (!m) return MJB_OK;
    mjb::const_bank_release(m);
    cudaFree(m->d_params);
    free(m->h_params);
    free(m);
    return MJB_OK;
}
