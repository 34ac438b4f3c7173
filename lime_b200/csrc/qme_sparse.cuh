// Sparse-operator (ELL) kernels for   d rho/dt = G rho + rho G^H + sum_s X_s rho Z_s^H
// (structured Hamiltonians: Jaynes-Cummings / Rabi ladders, tight-binding chains ...;
//  lime reaches this case by handing scipy.sparse CSR operands to _lindblad,
//  lime/oqs.py:1590-1688 through lime/phys.py:741-748).
//
//   k[i,j] = sum_p G[i,p] y[cG(i,p), j] + sum_p conj(G[j,p]) y[i, cG(j,p)]
//          + sum_s sum_p sum_q X_s[i,p] conj(Z_s[j,q]) y[cX(i,p), cZ(j,q)]
//
// Kernels:
//   qme_ell_global  : any N; one CTA per density matrix, all nsteps fused in one launch,
//                     stage vectors in a global scratch (L2 resident).
//   qme_ell_cluster : N*N*16 B * 2 <= cluster shared memory; one thread-block cluster per
//                     density matrix, rows split across the CTAs of the cluster, stage
//                     vectors in (distributed) shared memory, rho and the RK4 accumulator
//                     in registers, halo rows pushed to the neighbour CTAs through DSMEM.
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

struct EllOp {
    int w;                  // entries per row (padded: val = 0, col = row)
    const int* col;         // [N][w]
    const cplx* val;        // [nb][N][w]
};

#define QME_MAXS 4
struct QmeEllArgs {
    int N, S, E, B, nsteps, traj_every;
    int nb;
    EllOp G;
    EllOp X[QME_MAXS];
    EllOp Z[QME_MAXS];
    // observables as COO over rho's linear index: obs_e = sum_n eval[n] * rho[eidx[n]]
    const int* eptr;        // [E+1]
    const int* eidx;
    const cplx* eval;
    cplx* rho;              // [B][N*N]
    cplx* ybuf;             // global kernel: [2][B][N*N] scratch
    cplx* accbuf;           // global kernel: [B][N*N] scratch
    cplx* obs;              // [nsteps][B][E]
    cplx* traj;
    double dt;
    // cluster kernel geometry
    int step_vals;          // operator values are indexed by the RK4 STEP (time-dependent generator), not by the batch point
    int rows_per_cta;       // owned rows per CTA
    int halo;               // rows needed above/below the owned block
};

template <bool LDG> __device__ __forceinline__ cplx ell_ld(const cplx* p) {
    if (LDG) return __ldg(p);
    return *p;
}
template <bool LDG> __device__ __forceinline__ int ell_ld(const int* p) {
    if (LDG) return __ldg(p);
    return *p;
}

// LDG = true: operator arrays are read-only global memory (ld.global.nc);
// LDG = false: they were staged in shared memory.  y is never read through the nc path
// (it is rewritten inside the same kernel).
template <bool LDG>
__device__ __forceinline__ cplx ell_rhs_elem(const QmeEllArgs& a, size_t vb, int i, int j,
                                             const cplx* y, int ld, int row0) {
    // y points at a buffer whose row r lives at y[(r - row0) * ld + col]
    cplx s = cmake(0, 0);
    const int N = a.N;
    {
        const int w = a.G.w;
        const int* ci = a.G.col + (size_t)i * w;
        const cplx* vi = a.G.val + (vb * N + i) * w;
        for (int p = 0; p < w; ++p)
            cfma(s, ell_ld<LDG>(vi + p), y[(size_t)(ell_ld<LDG>(ci + p) - row0) * ld + j]);
        const int* cj = a.G.col + (size_t)j * w;
        const cplx* vj = a.G.val + (vb * N + j) * w;
        const cplx* yr = y + (size_t)(i - row0) * ld;
        for (int p = 0; p < w; ++p) cfma_conjb(s, yr[ell_ld<LDG>(cj + p)], ell_ld<LDG>(vj + p));
    }
    for (int sw = 0; sw < a.S; ++sw) {
        const int wx = a.X[sw].w, wz = a.Z[sw].w;
        const int* cx = a.X[sw].col + (size_t)i * wx;
        const cplx* vx = a.X[sw].val + (vb * N + i) * wx;
        const int* cz = a.Z[sw].col + (size_t)j * wz;
        const cplx* vz = a.Z[sw].val + (vb * N + j) * wz;
        for (int p = 0; p < wx; ++p) {
            const cplx xa = ell_ld<LDG>(vx + p);
            const cplx* yr = y + (size_t)(ell_ld<LDG>(cx + p) - row0) * ld;
            cplx inner = cmake(0, 0);
            for (int q = 0; q < wz; ++q) cfma_conjb(inner, yr[ell_ld<LDG>(cz + q)], ell_ld<LDG>(vz + q));
            cfma(s, xa, inner);
        }
    }
    return s;
}

__device__ __forceinline__ cplx block_reduce_cplx(cplx v, cplx* red /*[32]*/) {
    for (int off = 16; off > 0; off >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, off);
        v.y += __shfl_down_sync(0xffffffffu, v.y, off);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    cplx s = cmake(0, 0);
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) s = cadd(s, red[w]);
    return s;   // valid in thread 0
}

__global__ void __launch_bounds__(1024, 1)
qme_ell_global(QmeEllArgs a) {
    __shared__ cplx red[32];
    const int N = a.N, NN = N * N;
    const int b = blockIdx.x;
    cplx* rho = a.rho + (size_t)b * NN;
    cplx* y0 = a.ybuf + (size_t)b * NN;
    cplx* y1 = a.ybuf + ((size_t)a.B + b) * NN;
    cplx* acc = a.accbuf + (size_t)b * NN;
    const double dt = a.dt, hdt = 0.5 * a.dt;
    for (int step = 0; step < a.nsteps; ++step) {
        // driven problems (_lindblad_driven with CSR operands, lime/oqs.py:1691-1800): one set of generator values per
        // step, frozen over the four stages
        const size_t vb = a.step_vals ? (size_t)step : ((a.nb > 1) ? (size_t)b : 0);
        for (int stage = 0; stage < 4; ++stage) {
            const cplx* yin = (stage == 0) ? rho : ((stage == 2) ? y1 : y0);
            cplx* yout = (stage == 1) ? y1 : y0;
            for (int idx = threadIdx.x; idx < NN; idx += blockDim.x) {
                const int i = idx / N, j = idx - i * N;
                cplx k = ell_rhs_elem<true>(a, vb, i, j, yin, N, 0);
                cplx r = rho[idx];
                if (stage == 0) {
                    acc[idx] = k;
                    yout[idx] = cmake(fma(hdt, k.x, r.x), fma(hdt, k.y, r.y));
                } else if (stage == 1) {
                    cplx ac = acc[idx]; rfma(ac, 2.0, k); acc[idx] = ac;
                    yout[idx] = cmake(fma(hdt, k.x, r.x), fma(hdt, k.y, r.y));
                } else if (stage == 2) {
                    cplx ac = acc[idx]; rfma(ac, 2.0, k); acc[idx] = ac;
                    yout[idx] = cmake(fma(dt, k.x, r.x), fma(dt, k.y, r.y));
                } else {
                    cplx tot = cadd(acc[idx], k);
                    r.x += tot.x / 6.0 * dt;
                    r.y += tot.y / 6.0 * dt;
                    rho[idx] = r;
                }
            }
            __syncthreads();
        }
        if (a.obs) {
            for (int e = 0; e < a.E; ++e) {
                cplx v = cmake(0, 0);
                for (int n = a.eptr[e] + threadIdx.x; n < a.eptr[e + 1]; n += blockDim.x)
                    cfma(v, a.eval[n], rho[a.eidx[n]]);
                cplx s = block_reduce_cplx(v, red);
                if (threadIdx.x == 0) a.obs[((size_t)step * a.B + b) * a.E + e] = s;
            }
        }
        if (a.traj && ((step + 1) % a.traj_every) == 0) {
            cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * NN;
            for (int idx = threadIdx.x; idx < NN; idx += blockDim.x) dst[idx] = rho[idx];
        }
    }
}

// one RHS evaluation (drop-in for lime's liouvillian(rho, H, c_ops), lime/oqs.py:706-713)
__global__ void __launch_bounds__(256)
qme_ell_rhs(QmeEllArgs a, const cplx* __restrict__ in, cplx* __restrict__ out) {
    const int N = a.N, NN = N * N;
    const int b = blockIdx.y;
    const size_t vb = (a.nb > 1) ? (size_t)b : 0;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= NN) return;
    const int i = idx / N, j = idx - i * N;
    out[(size_t)b * NN + idx] = ell_rhs_elem<true>(a, vb, i, j, in + (size_t)b * NN, N, 0);
}
