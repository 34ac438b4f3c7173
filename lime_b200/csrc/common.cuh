// Shared device/host helpers for liblime_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>
#include <complex>

typedef double2 cplx;                       // interleaved complex128, same bytes as numpy complex128
typedef std::complex<double> hcplx;

// ---------------------------------------------------------------- error state
namespace limeb200 {
void set_error(const char* fmt, ...);
const char* get_error();
}
#define LB_OK 0
#define LB_ERR_ARG (-1)
#define LB_ERR_CUDA (-2)
#define LB_ERR_UNSUPPORTED (-3)
#define LB_ERR_STATE (-4)

#define LB_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            limeb200::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,            \
                                cudaGetErrorName(_e), cudaGetErrorString(_e));            \
            return LB_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define LB_REQUIRE(cond, ...)                                                             \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            limeb200::set_error(__VA_ARGS__);                                             \
            return LB_ERR_ARG;                                                            \
        }                                                                                 \
    } while (0)

// ---------------------------------------------------------------- complex math
__host__ __device__ __forceinline__ cplx cmake(double r, double i) { return make_double2(r, i); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cscale(double s, cplx a) { return make_double2(s * a.x, s * a.y); }
// acc += a*b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
// acc += a*conj(b)
__device__ __forceinline__ void cfma_conjb(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.y, b.x, acc.y);
    acc.y = fma(-a.x, b.y, acc.y);
}
// acc += s*a  (real s)
__device__ __forceinline__ void rfma(cplx& acc, double s, cplx a) {
    acc.x = fma(s, a.x, acc.x);
    acc.y = fma(s, a.y, acc.y);
}
// 1/(x + i y)
__device__ __forceinline__ cplx crecip(double x, double y) {
    double d = 1.0 / (x * x + y * y);
    return make_double2(x * d, -y * d);
}

// ---------------------------------------------------------------- misc
template <typename T>
static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

struct DevBuf {     // RAII device buffer owned by a plan
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    cudaError_t upload(const void* h, size_t n) {
        cudaError_t e = alloc(n);
        if (e != cudaSuccess || n == 0) return e;
        return cudaMemcpy(p, h, n, cudaMemcpyHostToDevice);
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// RK4 stage constants: y_stage = rho + a_s * k_{s-1};  acc += w_s * k_s
//   lime/phys.py:636-649: k2,k3 at dt/2, k4 at dt; rho += (k1 + 2k2 + 2k3 + k4)/6 * dt
