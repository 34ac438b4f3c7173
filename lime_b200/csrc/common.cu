// Error state and device queries of liblime_b200.
#include "../../include/lime_b200.h"
#include "common.cuh"

namespace limeb200 {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }
}  // namespace limeb200

extern "C" {

int limeb200_version(void) { return LIMEB200_VERSION; }
const char* limeb200_last_error(void) { return limeb200::get_error(); }

int limeb200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                         long long* smem_optin, long long* l2_bytes) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        limeb200::set_error("no CUDA device available (%s): liblime_b200 has no CPU fallback",
                            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return LB_ERR_CUDA;
    }
    LB_REQUIRE(device >= 0 && device < ndev, "device %d out of range", device);
    cudaDeviceProp prop;
    LB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (smem_optin) *smem_optin = (long long)prop.sharedMemPerBlockOptin;
    if (l2_bytes) *l2_bytes = (long long)prop.l2CacheSize;
    return LB_OK;
}

int limeb200_memcpy_d2d(void* d_dst, const void* d_src, long long bytes, void* stream) {
    LB_REQUIRE(d_dst && d_src && bytes >= 0, "bad arguments");
    LB_CUDA(cudaMemcpyAsync(d_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return LB_OK;
}

}  // extern "C"
