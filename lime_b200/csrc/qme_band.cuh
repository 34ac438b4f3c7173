// Register-tiled cluster propagator for structured (few entries per row) operators:
//
//     d rho/dt = G rho + rho G^H + sum_s X_s rho Z_s^H        (lime/oqs.py:706-723 in
//     generator form; Jaynes-Cummings / Rabi ladders, tight-binding chains, ...)
//
// Same residency scheme as qme_ell_cluster (one thread-block cluster per density matrix,
// row strips per CTA, ping-pong stage vectors in shared memory, halo rows pushed to the
// neighbour CTAs through distributed shared memory, one cluster barrier per RK4 stage, all
// nsteps fused in one launch) but with a static work assignment that removes the per-element
// index arithmetic of the ELL kernel:
//
//   * a warp owns TR consecutive rows; lane l owns the columns j_u = l + 32 u (u < TC), so
//     every shared-memory access of a warp is a contiguous 512-byte row segment (no bank
//     conflicts) and TR*TC elements of rho and of the RK4 accumulator live in registers;
//   * the operators are split on the host into the diagonal of G (merged left/right term
//     (G_ii + conj G_jj) y_ij), NOFF off-diagonal slots of G and one entry per row of each
//     X_s / Z_s; the column-side ("right") coefficients depend only on the lane and are
//     loaded into registers once per launch, the row-side ("left") ones are warp-uniform;
//   * purely imaginary off-diagonal G (real symmetric H: G = -iH - 1/2 sum l^H l) and real
//     X/Z halve the FP64 work (GT / XT template flags).
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
namespace cgb = cooperative_groups;

#define QME_BAND_MAXS 2

struct QmeBandArgs {
    int N, E, B, nsteps, traj_every, nb;
    int R, h, C;                 // rows per CTA, halo rows, CTAs per cluster
    const cplx* gd;              // [nb][N]        diagonal of G
    const int* gcol;             // [NOFF][N]      off-diagonal column (padding: col = row, val = 0)
    const cplx* gval;            // [nb][NOFF][N]
    const int* xcol[QME_BAND_MAXS];    // [N]      single entry per row (padding: col = row, val = 0)
    const cplx* xval[QME_BAND_MAXS];   // [nb][N]
    const int* zcol[QME_BAND_MAXS];
    const cplx* zval[QME_BAND_MAXS];
    const int* perm;             // [N] new -> old or null
    const int* eptr;             // observables as COO over the permuted linear index
    const int* eidx;
    const cplx* eval;
    cplx* rho;                   // [B][N][N] in/out
    cplx* obs;                   // [nsteps][B][E] or null
    cplx* traj;                  // [nsteps/traj_every][B][N][N] or null
    double dt;
};

// GT: 0 complex off-diagonal G, 1 purely imaginary.  XT: 0 complex X/Z, 1 real.
template <int TR, int TC, int NOFF, int S, int GT, int XT>
__global__ void __launch_bounds__(256, 1)
qme_band_kernel(QmeBandArgs a) {
    extern __shared__ double2 smem[];
    cgb::cluster_group cluster = cgb::this_cluster();
    const int N = a.N, R = a.R, h = a.h, C = a.C;
    const int rank = (C > 1) ? (int)cluster.block_rank() : 0;
    const int b = blockIdx.x / C;
    const size_t vb = (a.nb > 1) ? (size_t)b : 0;
    const int row_lo = rank * R;
    const int row_hi = min(N, row_lo + R);
    const int buf_rows = R + 2 * h;
    const int row0 = row_lo - h;
    const int T = blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    cplx* ybuf0 = smem;
    cplx* ybuf1 = ybuf0 + (size_t)buf_rows * N;
    cplx* red = ybuf1 + (size_t)buf_rows * N;          // [32]
    cplx* part = red + 32;                             // [2][C][E] (rank 0)

    // ---- load rho (own + halo rows, permuted basis) into stage buffer 0
    const cplx* grho = a.rho + (size_t)b * N * N;
    for (int l = threadIdx.x; l < buf_rows * N; l += T) {
        int r = row0 + l / N, c = l % N;
        cplx v = cmake(0, 0);
        if (r >= 0 && r < N) {
            int gr = a.perm ? a.perm[r] : r, gc = a.perm ? a.perm[c] : c;
            v = grho[(size_t)gr * N + gc];
        }
        ybuf0[l] = v;
        ybuf1[l] = cmake(0, 0);
    }
    // ---- column-side coefficients (registers, whole launch)
    int jc[TC];
    bool okc[TC];
    cplx gdj[TC];                    // conj(G_jj)
    int offR[NOFF][TC];              // byte offset of column c(j,q) inside a row
    cplx valR[NOFF][TC];             // conj(G[j][q])   (GT == 1: only .y is used)
    int offZ[S > 0 ? S : 1][TC];
    cplx valZ[S > 0 ? S : 1][TC];    // conj(Z[j])      (XT == 1: only .x is used)
#pragma unroll
    for (int u = 0; u < TC; ++u) {
        const int j = lane + 32 * u;
        okc[u] = j < N;
        jc[u] = okc[u] ? j : 0;
        gdj[u] = okc[u] ? cconj(a.gd[vb * N + jc[u]]) : cmake(0, 0);
#pragma unroll
        for (int q = 0; q < NOFF; ++q) {
            offR[q][u] = a.gcol[q * N + jc[u]] * 16;
            valR[q][u] = okc[u] ? cconj(a.gval[(vb * NOFF + q) * N + jc[u]]) : cmake(0, 0);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            offZ[s][u] = a.zcol[s][jc[u]] * 16;
            valZ[s][u] = okc[u] ? cconj(a.zval[s][vb * N + jc[u]]) : cmake(0, 0);
        }
    }
    __syncthreads();
    cplx rho[TR][TC], acc[TR][TC];
    const int i0 = row_lo + warp * TR;
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int u = 0; u < TC; ++u) {
            const int i = i0 + r;
            rho[r][u] = (i < row_hi && okc[u]) ? ybuf0[(size_t)(i - row0) * N + jc[u]] : cmake(0, 0);
            acc[r][u] = cmake(0, 0);
        }
    cplx* up0 = nullptr; cplx* up1 = nullptr; cplx* dn0 = nullptr; cplx* dn1 = nullptr;
    if (C > 1) {
        if (rank > 0) { up0 = cluster.map_shared_rank(ybuf0, rank - 1); up1 = cluster.map_shared_rank(ybuf1, rank - 1); }
        if (rank < C - 1) { dn0 = cluster.map_shared_rank(ybuf0, rank + 1); dn1 = cluster.map_shared_rank(ybuf1, rank + 1); }
        cluster.sync();
    }
    cplx* part0 = (C > 1) ? cluster.map_shared_rank(part, 0) : part;

    const double dt = a.dt, hdt = 0.5 * a.dt, w6 = a.dt / 6.0;
    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const cplx* yin = (stage & 1) ? ybuf1 : ybuf0;
            cplx* yout = (stage & 1) ? ybuf0 : ybuf1;
            cplx* rup = (stage & 1) ? up0 : up1;
            cplx* rdn = (stage & 1) ? dn0 : dn1;
            const double cy = (stage == 2) ? dt : hdt;
#pragma unroll
            for (int r = 0; r < TR; ++r) {
                const int i = i0 + r;
                if (i >= row_hi) break;                           // warp-uniform
                const char* yrow = reinterpret_cast<const char*>(yin + (size_t)(i - row0) * N);
                const cplx gdi = __ldg(a.gd + vb * N + i);
                cplx k[TC];
#pragma unroll
                for (int u = 0; u < TC; ++u) {
                    const cplx y = *reinterpret_cast<const cplx*>(yrow + jc[u] * 16);
                    const double dr = gdi.x + gdj[u].x, di = gdi.y + gdj[u].y;
                    k[u].x = dr * y.x - di * y.y;
                    k[u].y = dr * y.y + di * y.x;
                }
#pragma unroll
                for (int q = 0; q < NOFF; ++q) {
                    // left: G[i][c] y[c][j]  (c, value warp-uniform)
                    const int c = __ldg(a.gcol + q * N + i);
                    const cplx v = __ldg(a.gval + (vb * NOFF + q) * N + i);
                    const cplx* yl = yin + (size_t)(c - row0) * N;
#pragma unroll
                    for (int u = 0; u < TC; ++u) {
                        const cplx y = yl[jc[u]];
                        if (GT == 1) {
                            k[u].x = fma(-v.y, y.y, k[u].x);
                            k[u].y = fma(v.y, y.x, k[u].y);
                        } else {
                            cfma(k[u], v, y);
                        }
                    }
                    // right: y[i][c(j,q)] conj(G[j][q])
#pragma unroll
                    for (int u = 0; u < TC; ++u) {
                        const cplx y = *reinterpret_cast<const cplx*>(yrow + offR[q][u]);
                        if (GT == 1) {
                            k[u].x = fma(-valR[q][u].y, y.y, k[u].x);
                            k[u].y = fma(valR[q][u].y, y.x, k[u].y);
                        } else {
                            cfma(k[u], valR[q][u], y);
                        }
                    }
                }
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int cx = __ldg(a.xcol[s] + i);
                    const cplx xv = __ldg(a.xval[s] + vb * N + i);
                    const char* ys = reinterpret_cast<const char*>(yin + (size_t)(cx - row0) * N);
#pragma unroll
                    for (int u = 0; u < TC; ++u) {
                        const cplx y = *reinterpret_cast<const cplx*>(ys + offZ[s][u]);
                        if (XT == 1) {
                            const double cf = xv.x * valZ[s][u].x;
                            k[u].x = fma(cf, y.x, k[u].x);
                            k[u].y = fma(cf, y.y, k[u].y);
                        } else {
                            cfma(k[u], cmul(xv, valZ[s][u]), y);
                        }
                    }
                }
                // RK4 stage algebra (lime/phys.py:636-649) and write-out of the next stage vector
                const int lr = i - row0;
                const bool pu = rup && (i - row_lo < h);
                const bool pd = rdn && (i >= row_lo + R - h);
#pragma unroll
                for (int u = 0; u < TC; ++u) {
                    cplx yn;
                    if (stage == 0) {
                        acc[r][u] = k[u];
                        yn = cmake(fma(cy, k[u].x, rho[r][u].x), fma(cy, k[u].y, rho[r][u].y));
                    } else if (stage < 3) {
                        rfma(acc[r][u], 2.0, k[u]);
                        yn = cmake(fma(cy, k[u].x, rho[r][u].x), fma(cy, k[u].y, rho[r][u].y));
                    } else {
                        rho[r][u].x = fma(w6, acc[r][u].x + k[u].x, rho[r][u].x);
                        rho[r][u].y = fma(w6, acc[r][u].y + k[u].y, rho[r][u].y);
                        yn = rho[r][u];
                    }
                    if (okc[u]) {
                        yout[(size_t)lr * N + jc[u]] = yn;
                        if (pu) rup[(size_t)(i - row_lo + R + h) * N + jc[u]] = yn;
                        if (pd) rdn[(size_t)(i - row_lo - R + h) * N + jc[u]] = yn;
                    }
                }
            }
            if (C > 1) cluster.sync(); else __syncthreads();
            if (stage == 0 && a.obs && step > 0 && rank == 0 && threadIdx.x < a.E) {
                const cplx* pp = part + (size_t)((step - 1) & 1) * C * a.E;
                cplx sum = cmake(0, 0);
                for (int c = 0; c < C; ++c) sum = cadd(sum, pp[c * a.E + threadIdx.x]);
                a.obs[((size_t)(step - 1) * a.B + b) * a.E + threadIdx.x] = sum;
            }
        }
        // ybuf0 now holds rho_{n+1} (own + halo rows)
        if (a.obs) {
            for (int e = 0; e < a.E; ++e) {
                cplx v = cmake(0, 0);
                for (int n = a.eptr[e] + threadIdx.x; n < a.eptr[e + 1]; n += T) {
                    int idx = a.eidx[n];
                    int i = idx / N;
                    if (i >= row_lo && i < row_hi) cfma(v, a.eval[n], ybuf0[(size_t)(i - row0) * N + (idx - i * N)]);
                }
                for (int off = 16; off > 0; off >>= 1) {
                    v.x += __shfl_down_sync(0xffffffffu, v.x, off);
                    v.y += __shfl_down_sync(0xffffffffu, v.y, off);
                }
                __syncthreads();
                if (lane == 0) red[warp] = v;
                __syncthreads();
                if (threadIdx.x == 0) {
                    cplx sum = cmake(0, 0);
                    for (int w = 0; w < (T + 31) / 32; ++w) sum = cadd(sum, red[w]);
                    part0[(size_t)(step & 1) * C * a.E + rank * a.E + e] = sum;
                }
            }
        }
        if (a.traj && ((step + 1) % a.traj_every) == 0) {
            cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * N * N;
#pragma unroll
            for (int r = 0; r < TR; ++r)
#pragma unroll
                for (int u = 0; u < TC; ++u) {
                    const int i = i0 + r;
                    if (i < row_hi && okc[u]) {
                        int gr = a.perm ? a.perm[i] : i, gc = a.perm ? a.perm[jc[u]] : jc[u];
                        dst[(size_t)gr * N + gc] = rho[r][u];
                    }
                }
        }
    }
    if (a.obs && a.nsteps > 0) {
        if (C > 1) cluster.sync(); else __syncthreads();
        if (rank == 0 && threadIdx.x < a.E) {
            const cplx* pp = part + (size_t)((a.nsteps - 1) & 1) * C * a.E;
            cplx sum = cmake(0, 0);
            for (int c = 0; c < C; ++c) sum = cadd(sum, pp[c * a.E + threadIdx.x]);
            a.obs[((size_t)(a.nsteps - 1) * a.B + b) * a.E + threadIdx.x] = sum;
        }
    }
    cplx* out = a.rho + (size_t)b * N * N;
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int u = 0; u < TC; ++u) {
            const int i = i0 + r;
            if (i < row_hi && okc[u]) {
                int gr = a.perm ? a.perm[i] : i, gc = a.perm ? a.perm[jc[u]] : jc[u];
                out[(size_t)gr * N + gc] = rho[r][u];
            }
        }
    if (C > 1) cluster.sync();      // keep shared memory alive until the neighbours' remote stores are done
}

// geometry: smallest cluster with R = ceil(N/C) <= 32 rows per CTA that fits shared memory
static inline bool qme_band_geometry(int N, int E, int bandwidth, long long smem_optin,
                                     int* Cout, int* Rout, size_t* smem_out) {
    if (N > 128) return false;
    const int h = bandwidth;
    for (int c = 1; c <= 8; c *= 2) {
        int r = (N + c - 1) / c;
        if (r > 32) continue;
        if (c > 1 && h > r) break;
        size_t need = (size_t)2 * (r + 2 * h) * N * 16 + (size_t)(32 + 2 * c * (E > 0 ? E : 1)) * 16;
        if (need <= (size_t)smem_optin) { *Cout = c; *Rout = r; *smem_out = need; return true; }
    }
    return false;
}

// implemented in qme_band_inst_*.cu (one translation unit per (TC, NOFF) so that make -j compiles them in parallel)
int qme_band_launch_tc2_n2(const QmeBandArgs& a, int S, int GT, int XT, size_t smem, cudaStream_t st);
int qme_band_launch_tc2_n4(const QmeBandArgs& a, int S, int GT, int XT, size_t smem, cudaStream_t st);
int qme_band_launch_tc4_n2(const QmeBandArgs& a, int S, int GT, int XT, size_t smem, cudaStream_t st);
int qme_band_launch_tc4_n4(const QmeBandArgs& a, int S, int GT, int XT, size_t smem, cudaStream_t st);

template <int TC, int NOFF, int S, int GT, int XT>
static int qme_band_launch_one(const QmeBandArgs& a, size_t smem, cudaStream_t st) {
    constexpr int TR = 4;
    auto kern = qme_band_kernel<TR, TC, NOFF, S, GT, XT>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int W = (a.R + TR - 1) / TR;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.B * a.C);
    cfg.blockDim = dim3(32 * W);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    LB_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return LB_OK;
}

#define QME_BAND_DEFINE_LAUNCH(NAME, TC, NOFF)                                                        \
    int NAME(const QmeBandArgs& a, int S, int GT, int XT, size_t smem, cudaStream_t st) {             \
        if (S == 0) return GT ? qme_band_launch_one<TC, NOFF, 0, 1, 1>(a, smem, st)                   \
                              : qme_band_launch_one<TC, NOFF, 0, 0, 1>(a, smem, st);                  \
        if (S == 1) {                                                                                 \
            if (GT) return XT ? qme_band_launch_one<TC, NOFF, 1, 1, 1>(a, smem, st)                   \
                              : qme_band_launch_one<TC, NOFF, 1, 1, 0>(a, smem, st);                  \
            return XT ? qme_band_launch_one<TC, NOFF, 1, 0, 1>(a, smem, st)                           \
                      : qme_band_launch_one<TC, NOFF, 1, 0, 0>(a, smem, st);                          \
        }                                                                                             \
        if (GT) return XT ? qme_band_launch_one<TC, NOFF, 2, 1, 1>(a, smem, st)                       \
                          : qme_band_launch_one<TC, NOFF, 2, 1, 0>(a, smem, st);                      \
        return XT ? qme_band_launch_one<TC, NOFF, 2, 0, 1>(a, smem, st)                               \
                  : qme_band_launch_one<TC, NOFF, 2, 0, 0>(a, smem, st);                              \
    }
