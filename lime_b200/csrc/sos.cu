// Sum-over-states response functions in factorised form (lime/signal/sos.py).
//
// Every third-order pathway in sos.py is a sum over state triples (b,c,d) of
//   mu mu mu mu * G(axis-1; pole) * G(axis-2; pole) * U(fixed delay)
// with G(w) = 1/(w - dE + i Gamma) (sos.py:388-403, 512-527, 618-633, 940-962, 1048-1069).
// Grouping the triple by the index the axis-1 factor depends on gives
//   S_t[r][c] = sum_q A_t[q][r] * B_t[q][c],
//   F_t[q][n] = sum_d W_t[q][d] / (z_n - e1[q][d] + i g1[q][d]) [ / (z_n - e2[q][d] + i g2[q][d]) ]
// The O(states^3) weights W (dipole products x delay propagators) are host-side setup; the
// O(grid) work -- factor tables and the rank-R outer-product sum, one complex store per
// grid point -- runs here.
#include "../../include/lime_b200.h"
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
sos_factor_kernel(const double* __restrict__ z, int n, const cplx* __restrict__ W,
                  const double2* __restrict__ p1, const double2* __restrict__ p2,
                  int R, int D, cplx* __restrict__ F) {
    const int t = blockIdx.z, q = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double zi = z[i];
    const cplx* w = W + ((size_t)t * R + q) * D;
    const double2* a = p1 + (size_t)q * D;
    const double2* b = p2 ? p2 + (size_t)q * D : nullptr;
    cplx s = cmake(0, 0);
    for (int d = 0; d < D; ++d) {
        double2 pa = a[d];
        cplx g = crecip(zi - pa.x, pa.y);
        if (b) { double2 pb = b[d]; g = cmul(g, crecip(zi - pb.x, pb.y)); }
        cfma(s, w[d], g);
    }
    F[((size_t)t * R + q) * n + i] = s;
}

// time-domain factors (lime/signal/2DES.py:37-60): F[t][q][n] = sum_d W[t][q][d] * G(t_n; e, g),
// G(t) = -i theta(t) exp(-i e t - g t), theta(0) = 1
__global__ void __launch_bounds__(256)
sos_factor_time_kernel(const double* __restrict__ tt, int n, const cplx* __restrict__ W,
                       const double2* __restrict__ p1, int R, int D, cplx* __restrict__ F) {
    const int t = blockIdx.z, q = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ti = tt[i];
    const cplx* w = W + ((size_t)t * R + q) * D;
    const double2* a = p1 + (size_t)q * D;
    cplx s = cmake(0, 0);
    if (ti >= 0.0) {
        for (int d = 0; d < D; ++d) {
            const double2 pa = a[d];
            double sn, cs;
            sincos(-pa.x * ti, &sn, &cs);
            const double m = exp(-pa.y * ti);
            // -i * m (cs + i sn) = m (sn - i cs)
            cfma(s, w[d], cmake(m * sn, -m * cs));
        }
    }
    F[((size_t)t * R + q) * n + i] = s;
}

// out[t][r][c] (+)= scale * sum_q A[ta][q][r] B[tb][q][c];  32x32 tile per CTA of 256 threads,
// each thread 4 columns 8 apart: the 8 lanes of a row read / write consecutive 16-byte elements, so
// the shared-memory loads are conflict-free (4 consecutive columns per thread gave 4-way conflicts and
// a shared-memory-bound kernel, profiles/r01_sos_outer_v1.txt) and every store instruction writes
// full 128-byte row segments
__global__ void __launch_bounds__(256)
sos_outer_kernel(const cplx* __restrict__ A, int TA, const cplx* __restrict__ Bf, int TB, int R,
                 int nrow, int ncol, double scale, int accumulate, cplx* __restrict__ out) {
    extern __shared__ double2 smem[];
    cplx* As = smem;                 // [R][32]
    cplx* Bs = smem + (size_t)R * 32;  // [R][32]
    const int t = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const cplx* At = A + (size_t)(TA > 1 ? t : 0) * R * nrow;
    const cplx* Bt = Bf + (size_t)(TB > 1 ? t : 0) * R * ncol;
    for (int l = threadIdx.x; l < R * 32; l += 256) {
        int q = l >> 5, x = l & 31;
        As[l] = (r0 + x < nrow) ? At[(size_t)q * nrow + r0 + x] : cmake(0, 0);
        Bs[l] = (c0 + x < ncol) ? Bt[(size_t)q * ncol + c0 + x] : cmake(0, 0);
    }
    __syncthreads();
    const int tr = threadIdx.x >> 3;            // 0..31 row in tile
    const int tc = threadIdx.x & 7;             // columns tc + 8 u
    cplx acc[4] = {cmake(0, 0), cmake(0, 0), cmake(0, 0), cmake(0, 0)};
    for (int q = 0; q < R; ++q) {
        const cplx av = As[q * 32 + tr];
#pragma unroll
        for (int u = 0; u < 4; ++u) cfma(acc[u], av, Bs[q * 32 + tc + 8 * u]);
    }
    const int r = r0 + tr;
    if (r >= nrow) return;
    cplx* o = out + ((size_t)t * nrow + r) * ncol;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int c = c0 + tc + 8 * u;
        if (c < ncol) {
            cplx v = cscale(scale, acc[u]);
            if (accumulate) v = cadd(v, o[c]);
            o[c] = v;
        }
    }
}

__global__ void __launch_bounds__(256)
sos_tpa2d_kernel(const double* __restrict__ E, const double* __restrict__ dip, const double* __restrict__ gamma,
                 int N, const int* __restrict__ eidx, int ne, const int* __restrict__ fidx, int nf,
                 const double* __restrict__ omegap, int np, const double* __restrict__ omega1, int n1,
                 int time_order, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n1 || i >= np) return;
    const double wp = omegap[i], w1 = omega1[j], w2 = wp - w1;
    const int g = 0;
    double sig = 0.0;
    for (int ff = 0; ff < nf; ++ff) {
        const int f = fidx[ff];
        cplx tmp = cmake(0, 0);
        for (int mm = 0; mm < ne; ++mm) {
            const int m = eidx[mm];
            const double d = dip[f * N + m] * dip[m * N + g];
            const double de = E[m] - E[g];
            cplx t1 = crecip(w1 - de, gamma[m]);
            if (!time_order) t1 = cadd(t1, crecip(w2 - de, gamma[m]));
            tmp.x = fma(d, t1.x, tmp.x);
            tmp.y = fma(d, t1.y, tmp.y);
        }
        const double x = wp - E[f] + E[g], wd = gamma[f];
        const double lor = 1.0 / 3.14159265358979323846 * wd / (wd * wd + x * x);
        sig = fma(tmp.x * tmp.x + tmp.y * tmp.y, lor, sig);
    }
    out[(size_t)i * n1 + j] = sig;
}


// ETPA double time integral, second contraction (lime/signal/sos.py:1171-1223):
//   out[c][f] = sum_a exp(i (Ef[f] - alpha[c]) t2[a]) V[a][c]
// V = (theta o (J + J^T)) . U1 comes from limeb200_zgemm; one CTA per column c = (pump frequency, intermediate state)
__global__ void __launch_bounds__(256)
etpa_reduce_kernel(const cplx* __restrict__ V, const double* __restrict__ t2, int n2, const double* __restrict__ alpha,
                   int C, const double* __restrict__ Ef, int nf, cplx* __restrict__ out) {
    __shared__ cplx red[8];
    const int c = blockIdx.x;
    const double al = alpha[c];
    for (int f = 0; f < nf; ++f) {
        const double w = Ef[f] - al;
        cplx acc = cmake(0, 0);
        for (int a = threadIdx.x; a < n2; a += blockDim.x) {
            double sn, cs;
            sincos(w * t2[a], &sn, &cs);
            cfma(acc, cmake(cs, sn), V[(size_t)a * C + c]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            acc.x += __shfl_down_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_down_sync(0xffffffffu, acc.y, o);
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            cplx t = red[0];
            for (int w8 = 1; w8 < (int)(blockDim.x >> 5); ++w8) t = cadd(t, red[w8]);
            out[(size_t)c * nf + f] = t;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" {

int limeb200_sos_factor(const double* d_z, int n, const double* d_W, const double* d_p1,
                        const double* d_p2, int T, int R, int D, double* d_F, void* stream) {
    LB_REQUIRE(d_z && d_W && d_p1 && d_F, "null argument");
    LB_REQUIRE(n >= 1 && T >= 1 && R >= 1 && D >= 1, "bad sizes");
    LB_REQUIRE(R <= 65535 && T <= 65535, "R/T too large");
    dim3 grid(ceil_div(n, 256), R, T);
    sos_factor_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_z, n, (const cplx*)d_W, (const double2*)d_p1,
                                                             (const double2*)d_p2, R, D, (cplx*)d_F);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_sos_factor_time(const double* d_t, int n, const double* d_W, const double* d_p1,
                             int T, int R, int D, double* d_F, void* stream) {
    LB_REQUIRE(d_t && d_W && d_p1 && d_F, "null argument");
    LB_REQUIRE(n >= 1 && T >= 1 && R >= 1 && D >= 1, "bad sizes");
    LB_REQUIRE(R <= 65535 && T <= 65535, "R/T too large");
    dim3 grid(ceil_div(n, 256), R, T);
    sos_factor_time_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_t, n, (const cplx*)d_W, (const double2*)d_p1,
                                                                  R, D, (cplx*)d_F);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_sos_outer(const double* d_A, int TA, const double* d_B, int TB, int T, int R,
                       int nrow, int ncol, double scale, int accumulate, double* d_out, void* stream) {
    LB_REQUIRE(d_A && d_B && d_out, "null argument");
    LB_REQUIRE(T >= 1 && R >= 1 && nrow >= 1 && ncol >= 1, "bad sizes");
    LB_REQUIRE((TA == 1 || TA == T) && (TB == 1 || TB == T), "TA/TB must be 1 or T");
    size_t smem = (size_t)2 * R * 32 * 16;
    LB_REQUIRE(smem <= 200 * 1024, "rank R=%d too large for one pass (max 200)", R);
    if (smem > 48 * 1024)
        LB_CUDA(cudaFuncSetAttribute(sos_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(ncol, 32), ceil_div(nrow, 32), T);
    sos_outer_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const cplx*)d_A, TA, (const cplx*)d_B, TB, R,
                                                               nrow, ncol, scale, accumulate, (cplx*)d_out);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_sos_tpa2d(const double* d_E, const double* d_dip, const double* d_gamma, int N,
                       const int* d_eidx, int ne, const int* d_fidx, int nf,
                       const double* d_omegap, int np, const double* d_omega1, int n1,
                       int time_order, double* d_out, void* stream) {
    LB_REQUIRE(d_E && d_dip && d_gamma && d_eidx && d_fidx && d_omegap && d_omega1 && d_out, "null argument");
    LB_REQUIRE(N >= 1 && ne >= 0 && nf >= 0 && np >= 1 && n1 >= 1, "bad sizes");
    dim3 grid(ceil_div(n1, 256), np);
    sos_tpa2d_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_E, d_dip, d_gamma, N, d_eidx, ne, d_fidx, nf,
                                                            d_omegap, np, d_omega1, n1, time_order, d_out);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_etpa_reduce(const double* d_V, const double* d_t2, int n2, const double* d_alpha, int C,
                         const double* d_Ef, int nf, double* d_out, void* stream) {
    LB_REQUIRE(d_V && d_t2 && d_alpha && d_Ef && d_out, "null argument");
    LB_REQUIRE(n2 >= 1 && C >= 1 && nf >= 1, "bad sizes");
    etpa_reduce_kernel<<<C, 256, 0, (cudaStream_t)stream>>>((const cplx*)d_V, d_t2, n2, d_alpha, C, d_Ef, nf, (cplx*)d_out);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

}  // extern "C"
