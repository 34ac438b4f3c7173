// Fused Runge-Kutta-Fehlberg 4(5) stage updates on device-resident state vectors.
//
// lime's examples/rkf45_test.py:7,115 drives an `r8_rkf45(f, neqn, y, yp, t, tout, relerr, abserr, flag)`
// integrator (Fehlberg's 4(5) pair with the Shampine-Watts step-size control) from a module that is not in
// the lime tree; BASELINE's north_star asks for "fused RK4/RKF45 stage updates that keep rho resident in HBM".
// The step control is host logic (lime_b200/rkf45.py); what runs per stage is here:
//
//   limeb200_rkf45_stage  one kernel per Fehlberg stage: the stage argument (or the 4th-order solution) as a
//                         fused linear combination of y, yp and the slopes computed so far -- one pass over the
//                         state instead of one pass per slope;
//   limeb200_rkf45_error  the solution s, the error estimate and its reduction in ONE pass: per element
//                         ee/et with et = |y| + |s| + ae, reduced to max over the whole batch (warp shuffle,
//                         one atomicMax per CTA on the bit pattern of the non-negative double), plus min(et).
//
// Vectors are plain double arrays (a complex128 density matrix is 2 N^2 doubles, as in Burkardt's real interface
// applied to the interleaved data).  Grid-stride kernels sized to the machine (148 SMs x 8 CTAs).
#include "common.cuh"

namespace {

// Fehlberg stage arguments, written as in the classical formulation (integer numerators over a common
// denominator, evaluated innermost-first) so that results do not depend on the order of fused operations here
template <int STAGE>
__global__ void __launch_bounds__(256)
rkf45_stage_kernel(long long n, const double* __restrict__ y, const double* __restrict__ yp,
                   const double* __restrict__ f1, const double* __restrict__ f2, const double* __restrict__ f3,
                   const double* __restrict__ f4, const double* __restrict__ f5, double h, double* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double yi = y[i], p = yp[i];
        double v;
        if (STAGE == 1) {
            v = yi + (h / 4.0) * p;
        } else if (STAGE == 2) {
            v = yi + (3.0 * h / 32.0) * (p + 3.0 * f1[i]);
        } else if (STAGE == 3) {
            v = yi + (h / 2197.0) * (1932.0 * p + (7296.0 * f2[i] - 7200.0 * f1[i]));
        } else if (STAGE == 4) {
            v = yi + (h / 4104.0) * ((8341.0 * p - 845.0 * f3[i]) + (29440.0 * f2[i] - 32832.0 * f1[i]));
        } else {   // STAGE == 5
            v = yi + (h / 20520.0) * ((-6080.0 * p + (9295.0 * f3[i] - 5643.0 * f4[i])) + (41040.0 * f1[i] - 28352.0 * f2[i]));
        }
        out[i] = v;
    }
}

// s = y + h/7618050 (902880 yp + 3855735 f3 - 1371249 f4 + 3953664 f2 + 277020 f5);
// ee = |-2090 yp + 21970 f3 - 15048 f4 + 22528 f2 - 27360 f5|,  et = |y| + |s| + ae
// result[0] = max ee/et (as the bit pattern of a non-negative double), result[1] = min et
__global__ void __launch_bounds__(256)
rkf45_error_kernel(long long n, const double* __restrict__ y, const double* __restrict__ yp,
                   const double* __restrict__ f2, const double* __restrict__ f3, const double* __restrict__ f4,
                   const double* __restrict__ f5, double h, double ae, double* __restrict__ s,
                   unsigned long long* __restrict__ result) {
    double mx = 0.0, mn = 1e300;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double yi = y[i], p = yp[i], a2 = f2[i], a3 = f3[i], a4 = f4[i], a5 = f5[i];
        const double si = yi + (h / 7618050.0) * ((902880.0 * p + (3855735.0 * a3 - 1371249.0 * a4)) + (3953664.0 * a2 + 277020.0 * a5));
        s[i] = si;
        const double et = fabs(yi) + fabs(si) + ae;
        const double ee = fabs((-2090.0 * p + (21970.0 * a3 - 15048.0 * a4)) + (22528.0 * a2 - 27360.0 * a5));
        mn = fmin(mn, et);
        if (et > 0.0) mx = fmax(mx, ee / et);
    }
    for (int off = 16; off > 0; off >>= 1) {
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, off));
    }
    __shared__ double smx[8], smn[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { smx[warp] = mx; smn[warp] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mx = fmax(mx, smx[w]); mn = fmin(mn, smn[w]); }
        // non-negative IEEE doubles order like their bit patterns
        atomicMax(result, (unsigned long long)__double_as_longlong(mx));
        atomicMin(result + 1, (unsigned long long)__double_as_longlong(fmax(mn, 0.0)));
    }
}

// y += a * x  (the Euler extrapolation of the last sliver of an interval)
__global__ void __launch_bounds__(256)
rkf45_axpy_kernel(long long n, double a, const double* __restrict__ x, double* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = y[i] + a * x[i];
}

// start-up step size: result[0] = max over k of the bit pattern of tol_k (any tol > 0), result[1] = min candidate h
// with tol_k = relerr |y_k| + abserr and candidate (tol_k / |yp_k|)^(1/5) where tol_k < |yp_k| h^5
__global__ void __launch_bounds__(256)
rkf45_hinit_kernel(long long n, const double* __restrict__ y, const double* __restrict__ yp, double relerr,
                   double abserr, double h0, unsigned long long* __restrict__ result) {
    double tmax = 0.0, hmin = h0;
    const double h5 = h0 * h0 * h0 * h0 * h0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double tol = relerr * fabs(y[i]) + abserr;
        if (tol > 0.0) {
            tmax = fmax(tmax, tol);
            const double ypk = fabs(yp[i]);
            if (tol < ypk * h5) hmin = fmin(hmin, pow(tol / ypk, 0.2));
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        tmax = fmax(tmax, __shfl_down_sync(0xffffffffu, tmax, off));
        hmin = fmin(hmin, __shfl_down_sync(0xffffffffu, hmin, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(result, (unsigned long long)__double_as_longlong(tmax));
        atomicMin(result + 1, (unsigned long long)__double_as_longlong(hmin));
    }
}

inline int rk_grid(long long n) {
    long long b = (n + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int limeb200_rkf45_stage(int stage, long long n, const double* d_y, const double* d_yp, const double* d_f1,
                         const double* d_f2, const double* d_f3, const double* d_f4, const double* d_f5, double h,
                         double* d_out, void* stream) {
    LB_REQUIRE(stage >= 1 && stage <= 5, "stage must be 1..5");
    LB_REQUIRE(n >= 1 && d_y && d_yp && d_out, "null argument");
    LB_REQUIRE(stage < 2 || d_f1, "stage %d needs f1", stage);
    LB_REQUIRE(stage < 3 || d_f2, "stage %d needs f2", stage);
    LB_REQUIRE(stage < 4 || d_f3, "stage %d needs f3", stage);
    LB_REQUIRE(stage < 5 || d_f4, "stage %d needs f4", stage);
    cudaStream_t st = (cudaStream_t)stream;
    const int g = rk_grid(n);
    switch (stage) {
        case 1: rkf45_stage_kernel<1><<<g, 256, 0, st>>>(n, d_y, d_yp, d_f1, d_f2, d_f3, d_f4, d_f5, h, d_out); break;
        case 2: rkf45_stage_kernel<2><<<g, 256, 0, st>>>(n, d_y, d_yp, d_f1, d_f2, d_f3, d_f4, d_f5, h, d_out); break;
        case 3: rkf45_stage_kernel<3><<<g, 256, 0, st>>>(n, d_y, d_yp, d_f1, d_f2, d_f3, d_f4, d_f5, h, d_out); break;
        case 4: rkf45_stage_kernel<4><<<g, 256, 0, st>>>(n, d_y, d_yp, d_f1, d_f2, d_f3, d_f4, d_f5, h, d_out); break;
        default: rkf45_stage_kernel<5><<<g, 256, 0, st>>>(n, d_y, d_yp, d_f1, d_f2, d_f3, d_f4, d_f5, h, d_out); break;
    }
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_rkf45_error(long long n, const double* d_y, const double* d_yp, const double* d_f2, const double* d_f3,
                         const double* d_f4, const double* d_f5, double h, double ae, double* d_s,
                         double* d_result2, void* stream) {
    LB_REQUIRE(n >= 1 && d_y && d_yp && d_f2 && d_f3 && d_f4 && d_f5 && d_s && d_result2, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long init[2] = {0ull, 0x7fefffffffffffffull};     // 0.0, DBL_MAX
    LB_CUDA(cudaMemcpyAsync(d_result2, init, 16, cudaMemcpyHostToDevice, st));
    rkf45_error_kernel<<<rk_grid(n), 256, 0, st>>>(n, d_y, d_yp, d_f2, d_f3, d_f4, d_f5, h, ae, d_s,
                                                    reinterpret_cast<unsigned long long*>(d_result2));
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_rkf45_hinit(long long n, const double* d_y, const double* d_yp, double relerr, double abserr, double h0,
                         double* d_result2, void* stream) {
    LB_REQUIRE(n >= 1 && d_y && d_yp && d_result2, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long init[2];
    init[0] = 0ull;
    memcpy(&init[1], &h0, 8);
    LB_REQUIRE(h0 >= 0.0, "h0 must be non-negative");
    LB_CUDA(cudaMemcpyAsync(d_result2, init, 16, cudaMemcpyHostToDevice, st));
    rkf45_hinit_kernel<<<rk_grid(n), 256, 0, st>>>(n, d_y, d_yp, relerr, abserr, h0,
                                                    reinterpret_cast<unsigned long long*>(d_result2));
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_rkf45_axpy(long long n, double a, const double* d_x, double* d_y, void* stream) {
    LB_REQUIRE(n >= 1 && d_x && d_y, "null argument");
    rkf45_axpy_kernel<<<rk_grid(n), 256, 0, (cudaStream_t)stream>>>(n, a, d_x, d_y);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

}  // extern "C"
