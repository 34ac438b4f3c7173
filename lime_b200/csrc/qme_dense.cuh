// Dense-operator kernels for the quantum master equation in generator/sandwich form
//
//     d rho/dt = G rho + rho G^H + sum_s X_s rho Z_s^H
//
// which covers lime's Lindblad RHS (lime/oqs.py:706-723: G = -iH - 1/2 sum l^H l,
// X_s = Z_s = l_s) and the operator form of its Redfield RHS (lime/oqs.py:840-850,
// the same generator lime/oqs.py:528-579 expands into an N^2 x N^2 CSR matrix:
// G = -i diag(eps) - sum_k A_k Lam_k, sandwiches (A_k, Lam_k) and (Lam_k, A_k)).
//
// Two kernels:
//   qme_dense_onchip : N <= 64.  rho, the RK4 accumulator and the stage vector stay in
//                      registers / shared memory for ALL nsteps steps; HBM sees rho once
//                      on the way in and once on the way out (+ observables).
//   qme_dense_stage  : any N.  One launch per RK stage, 32x32 output tiles, the RK4
//                      axpy chain fused into the epilogue.
#pragma once
#include "common.cuh"

struct QmeDenseArgs {
    int N, S, E, nd, B, nsteps, traj_every;
    int nb;                 // operator batch: 1 (shared) or B (per-item values)
    const cplx* G;          // [nb][N*N]
    const cplx* Gh;         // [nb][N*N]   right generator (G^H unless set explicitly)
    const cplx* X;          // [nb][S][N*N]
    const cplx* Zh;         // [nb][S][N*N] Z_s^H
    const cplx* D;          // [nd][N*N]   drive generators:  G_k = G + sum_i coef[k][i] D_i
    const cplx* Dh;         // [nd][N*N]   right drives     Gr_k = Gr + sum_i (conj?)(coef) Dr_i
    const cplx* eT;         // [E][N*N]    observables, transposed: Tr(e rho) = sum eT[idx] rho[idx]
    const cplx* coef;       // [nsteps][nd]
    cplx* rho;              // [B][N*N] in/out
    cplx* obs;              // [nsteps][B][E] or null
    cplx* traj;             // [nsteps/traj_every][B][N*N] or null
    double dt;
    int slots;              // density matrices per CTA
    int tps;                // threads per slot
    int ops_in_smem;        // G,Gh,X,Zh staged in shared memory (nb == 1 only)
    int drive_conj;         // 1: Gr_k += conj(c) Dh ; 0: Gr_k += c Dh
};

// --------------------------------------------------------------------------------
// on-chip kernel
// --------------------------------------------------------------------------------
template <int EPT>
__global__ void __launch_bounds__(1024, 1)
qme_dense_onchip(QmeDenseArgs a) {
    extern __shared__ double2 smem[];
    const int N = a.N, NN = N * N, S = a.S;
    const int tps = a.tps;
    const int slot = threadIdx.x / tps;
    const int t = threadIdx.x - slot * tps;
    const int b = blockIdx.x * a.slots + slot;
    const bool active = b < a.B;

    cplx* y = smem + (size_t)slot * 2 * NN;
    cplx* tmp = y + NN;
    cplx* red = smem + (size_t)a.slots * 2 * NN;             // [slots][32]
    cplx* opsm = red + a.slots * 32;

    const size_t ob = (a.nb > 1 && active) ? (size_t)b : 0;
    const cplx* Gs = a.G + ob * NN;
    const cplx* Ghs = a.Gh + ob * NN;
    const cplx* Xs = a.X + ob * S * NN;
    const cplx* Zhs = a.Zh + ob * S * NN;
    if (a.ops_in_smem) {
        cplx* g = opsm;
        cplx* gh = g + NN;
        cplx* x = gh + NN;
        cplx* zh = x + (size_t)S * NN;
        for (int i = threadIdx.x; i < NN; i += blockDim.x) { g[i] = a.G[i]; gh[i] = a.Gh[i]; }
        for (int i = threadIdx.x; i < S * NN; i += blockDim.x) { x[i] = a.X[i]; zh[i] = a.Zh[i]; }
        Gs = g; Ghs = gh; Xs = x; Zhs = zh;
    }

    int ei[EPT], ej[EPT];
    bool ok[EPT];
    cplx rho[EPT], acc[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        int idx = t + e * tps;
        ok[e] = idx < NN;
        ei[e] = ok[e] ? idx / N : 0;
        ej[e] = ok[e] ? idx - ei[e] * N : 0;
        rho[e] = (ok[e] && active) ? a.rho[(size_t)b * NN + idx] : cmake(0, 0);
        acc[e] = cmake(0, 0);
        if (ok[e]) y[idx] = rho[e];
    }
    __syncthreads();

    const double dt = a.dt, hdt = 0.5 * a.dt;
    for (int step = 0; step < a.nsteps; ++step) {
        if (a.nd > 0) {          // time-dependent generator for this step (ops_in_smem guaranteed)
            cplx* g = opsm;
            cplx* gh = g + NN;
            for (int i = threadIdx.x; i < NN; i += blockDim.x) {
                cplx v = a.G[i], vh = a.Gh[i];
                for (int d = 0; d < a.nd; ++d) {
                    cplx c = a.coef[(size_t)step * a.nd + d];
                    cfma(v, c, a.D[(size_t)d * NN + i]);
                    cfma(vh, a.drive_conj ? cconj(c) : c, a.Dh[(size_t)d * NN + i]);
                }
                g[i] = v; gh[i] = vh;
            }
            __syncthreads();
        }
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            cplx k[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                cplx s = cmake(0, 0);
                if (ok[e]) {
                    const cplx* gr = Gs + ei[e] * N;
                    const cplx* yr = y + ei[e] * N;
                    const int j = ej[e];
                    for (int q = 0; q < N; ++q) {
                        cfma(s, gr[q], y[q * N + j]);
                        cfma(s, yr[q], Ghs[q * N + j]);
                    }
                }
                k[e] = s;
            }
            for (int sw = 0; sw < S; ++sw) {
                const cplx* zh = Zhs + (size_t)sw * NN;
                const cplx* xs = Xs + (size_t)sw * NN;
                cplx tv[EPT];
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    cplx s = cmake(0, 0);
                    if (ok[e]) {
                        const cplx* yr = y + ei[e] * N;
                        const int j = ej[e];
                        for (int q = 0; q < N; ++q) cfma(s, yr[q], zh[q * N + j]);
                    }
                    tv[e] = s;
                }
                __syncthreads();
#pragma unroll
                for (int e = 0; e < EPT; ++e)
                    if (ok[e]) tmp[t + e * tps] = tv[e];
                __syncthreads();
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    if (ok[e]) {
                        const cplx* xr = xs + ei[e] * N;
                        const int j = ej[e];
                        cplx s = k[e];
                        for (int q = 0; q < N; ++q) cfma(s, xr[q], tmp[q * N + j]);
                        k[e] = s;
                    }
                }
            }
            cplx yn[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                if (stage == 0) {
                    acc[e] = k[e];
                    yn[e] = cmake(fma(hdt, k[e].x, rho[e].x), fma(hdt, k[e].y, rho[e].y));
                } else if (stage == 1) {
                    rfma(acc[e], 2.0, k[e]);
                    yn[e] = cmake(fma(hdt, k[e].x, rho[e].x), fma(hdt, k[e].y, rho[e].y));
                } else if (stage == 2) {
                    rfma(acc[e], 2.0, k[e]);
                    yn[e] = cmake(fma(dt, k[e].x, rho[e].x), fma(dt, k[e].y, rho[e].y));
                } else {
                    cplx tot = cadd(acc[e], k[e]);
                    rho[e].x += tot.x / 6.0 * dt;
                    rho[e].y += tot.y / 6.0 * dt;
                    yn[e] = rho[e];
                }
            }
            __syncthreads();
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (ok[e]) y[t + e * tps] = yn[e];
            __syncthreads();
        }
        // ---- observables Tr(e rho) of the state AFTER this step (lime/oqs.py:1674-1682)
        if (a.obs) {
            for (int eo = 0; eo < a.E; ++eo) {
                cplx v = cmake(0, 0);
#pragma unroll
                for (int e = 0; e < EPT; ++e)
                    if (ok[e]) cfma(v, a.eT[(size_t)eo * NN + t + e * tps], rho[e]);
                if (tps >= 32) {
                    for (int off = 16; off > 0; off >>= 1) {
                        v.x += __shfl_down_sync(0xffffffffu, v.x, off);
                        v.y += __shfl_down_sync(0xffffffffu, v.y, off);
                    }
                    if ((t & 31) == 0) red[slot * 32 + (t >> 5)] = v;
                    __syncthreads();
                    if (t == 0 && active) {
                        cplx s = cmake(0, 0);
                        for (int w = 0; w < tps / 32; ++w) s = cadd(s, red[slot * 32 + w]);
                        a.obs[((size_t)step * a.B + b) * a.E + eo] = s;
                    }
                    __syncthreads();
                } else {
                    for (int off = tps >> 1; off > 0; off >>= 1) {
                        v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
                        v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
                    }
                    if (t == 0 && active) a.obs[((size_t)step * a.B + b) * a.E + eo] = v;
                }
            }
        }
        if (a.traj && ((step + 1) % a.traj_every) == 0 && active) {
            cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * NN;
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (ok[e]) dst[t + e * tps] = rho[e];
        }
    }
    if (active) {
#pragma unroll
        for (int e = 0; e < EPT; ++e)
            if (ok[e]) a.rho[(size_t)b * NN + t + e * tps] = rho[e];
    }
}

// --------------------------------------------------------------------------------
// stage-wise tiled kernel (any N): out = sum_p A_p B_p with fused RK4 epilogue
// --------------------------------------------------------------------------------
#define QME_MAXPROD 18
struct QmeStageArgs {
    int N, B, nprod;
    int Mr, Nc, Kd;                  // DMMA kernel: products are [Mr x Kd] * [Kd x Nc] (0 = N: the square stage case)
    const cplx* A[QME_MAXPROD];      // left factors
    const cplx* Bm[QME_MAXPROD];     // right factors
    long long sA[QME_MAXPROD];       // batch strides (elements); 0 = shared
    long long sB[QME_MAXPROD];
    // epilogue
    int mode;                        // 0: out = k ; 1..4: RK4 stage
    cplx* out;                       // mode 0: [B][N*N] (stride sOut)
    long long sOut;
    cplx* rho;                       // [B][N*N]
    cplx* acc;                       // [B][N*N]
    cplx* ynext;                     // [B][N*N]
    double dt;
};

__global__ void __launch_bounds__(256)
qme_dense_stage(QmeStageArgs a) {
    // 32x32 output tile, 256 threads, 2x2 outputs per thread, k-tile 16
    __shared__ cplx As[32][17];
    __shared__ cplx Bs[16][33];
    const int N = a.N;
    const int b = blockIdx.z;
    const int ti = blockIdx.y * 32, tj = blockIdx.x * 32;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16
    cplx c00 = cmake(0, 0), c01 = c00, c10 = c00, c11 = c00;
    for (int p = 0; p < a.nprod; ++p) {
        const cplx* Ap = a.A[p] + (size_t)b * a.sA[p];
        const cplx* Bp = a.Bm[p] + (size_t)b * a.sB[p];
        for (int k0 = 0; k0 < N; k0 += 16) {
            for (int l = threadIdx.x; l < 32 * 16; l += 256) {
                int r = l >> 4, c = l & 15;
                int gi = ti + r, gk = k0 + c;
                As[r][c] = (gi < N && gk < N) ? Ap[(size_t)gi * N + gk] : cmake(0, 0);
            }
            for (int l = threadIdx.x; l < 16 * 32; l += 256) {
                int r = l >> 5, c = l & 31;
                int gk = k0 + r, gj = tj + c;
                Bs[r][c] = (gk < N && gj < N) ? Bp[(size_t)gk * N + gj] : cmake(0, 0);
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                cplx a0 = As[ty][q], a1 = As[ty + 16][q];
                cplx b0 = Bs[q][tx], b1 = Bs[q][tx + 16];
                cfma(c00, a0, b0); cfma(c01, a0, b1);
                cfma(c10, a1, b0); cfma(c11, a1, b1);
            }
            __syncthreads();
        }
    }
    cplx kk[4] = {c00, c01, c10, c11};
    const double dt = a.dt, hdt = 0.5 * a.dt;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int gi = ti + ty + ((q >> 1) ? 16 : 0);
        int gj = tj + tx + ((q & 1) ? 16 : 0);
        if (gi >= N || gj >= N) continue;
        size_t idx = (size_t)gi * N + gj;
        cplx k = kk[q];
        if (a.mode == 0) {
            a.out[(size_t)b * a.sOut + idx] = k;
        } else {
            size_t o = (size_t)b * N * N + idx;
            cplx r = a.rho[o];
            if (a.mode == 1) {
                a.acc[o] = k;
                a.ynext[o] = cmake(fma(hdt, k.x, r.x), fma(hdt, k.y, r.y));
            } else if (a.mode == 2) {
                cplx ac = a.acc[o]; rfma(ac, 2.0, k); a.acc[o] = ac;
                a.ynext[o] = cmake(fma(hdt, k.x, r.x), fma(hdt, k.y, r.y));
            } else if (a.mode == 3) {
                cplx ac = a.acc[o]; rfma(ac, 2.0, k); a.acc[o] = ac;
                a.ynext[o] = cmake(fma(dt, k.x, r.x), fma(dt, k.y, r.y));
            } else {
                cplx tot = cadd(a.acc[o], k);
                r.x += tot.x / 6.0 * dt;
                r.y += tot.y / 6.0 * dt;
                a.rho[o] = r;
                a.ynext[o] = r;
            }
        }
    }
}

// obs[b][e] = sum_idx eT[e][idx] * rho[b][idx]; one CTA per (e, b)
// Same contract as qme_dense_stage with a 64x64 output tile and a 4x4 register tile per thread (rows
// ty + 16 r, columns tx + 16 c): per k step a warp issues 4 broadcast loads of A and 4 conflict-free loads of B
// for 16 complex FMAs per thread, which moves the kernel from shared-memory bound to FP64-pipe bound.
// Global -> shared copies are software-pipelined through registers (next k-tile loaded while the current one
// is consumed).
__global__ void __launch_bounds__(256, 2)
qme_dense_stage64(QmeStageArgs a) {
    constexpr int KT = 8;
    __shared__ cplx As[2][KT][64];       // [k][row]   (transposed: row index fastest)
    __shared__ cplx Bs[2][KT][64];       // [k][col]
    const int N = a.N;
    const int b = blockIdx.z;
    const int ti = blockIdx.y * 64, tj = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    cplx c[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) c[r][q] = cmake(0, 0);
    // loader mapping: A tile 64 rows x KT cols (2 elements per thread), B tile KT rows x 64 cols (2 per thread)
    const int la_r = threadIdx.x >> 2, la_c = (threadIdx.x & 3) * 2;      // A: row, first of 2 consecutive k
    const int lb_r = threadIdx.x >> 5, lb_c = (threadIdx.x & 31) * 2;     // B: k row, first of 2 consecutive cols
    const int nk = (N + KT - 1) / KT;
    const int total = a.nprod * nk;
    cplx ra[2], rb[2];
    auto fetch = [&](int it) {
        const int p = it / nk, k0 = (it - p * nk) * KT;
        const cplx* Ap = a.A[p] + (size_t)b * a.sA[p];
        const cplx* Bp = a.Bm[p] + (size_t)b * a.sB[p];
        const int gi = ti + la_r;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int gk = k0 + la_c + u;
            ra[u] = (gi < N && gk < N) ? Ap[(size_t)gi * N + gk] : cmake(0, 0);
            const int gkb = k0 + lb_r, gj = tj + lb_c + u;
            rb[u] = (gkb < N && gj < N) ? Bp[(size_t)gkb * N + gj] : cmake(0, 0);
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            As[buf][la_c + u][la_r] = ra[u];
            Bs[buf][lb_r][lb_c + u] = rb[u];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int buf = it & 1;
        if (it + 1 < total) fetch(it + 1);
#pragma unroll
        for (int q = 0; q < KT; ++q) {
            cplx av[4], bv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { av[r] = As[buf][q][ty + 16 * r]; bv[r] = Bs[buf][q][tx + 16 * r]; }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) cfma(c[r][cc], av[r], bv[cc]);
        }
        if (it + 1 < total) stash(buf ^ 1);
        __syncthreads();
    }
    const double dt = a.dt, hdt = 0.5 * a.dt;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int gi = ti + ty + 16 * r, gj = tj + tx + 16 * cc;
            if (gi >= N || gj >= N) continue;
            const size_t idx = (size_t)gi * N + gj;
            const cplx k = c[r][cc];
            if (a.mode == 0) {
                a.out[(size_t)b * a.sOut + idx] = k;
            } else {
                const size_t o = (size_t)b * N * N + idx;
                cplx rr = a.rho[o];
                if (a.mode == 1) {
                    a.acc[o] = k;
                    a.ynext[o] = cmake(fma(hdt, k.x, rr.x), fma(hdt, k.y, rr.y));
                } else if (a.mode == 2) {
                    cplx ac = a.acc[o]; rfma(ac, 2.0, k); a.acc[o] = ac;
                    a.ynext[o] = cmake(fma(hdt, k.x, rr.x), fma(hdt, k.y, rr.y));
                } else if (a.mode == 3) {
                    cplx ac = a.acc[o]; rfma(ac, 2.0, k); a.acc[o] = ac;
                    a.ynext[o] = cmake(fma(dt, k.x, rr.x), fma(dt, k.y, rr.y));
                } else {
                    const cplx tot = cadd(a.acc[o], k);
                    rr.x += tot.x / 6.0 * dt;
                    rr.y += tot.y / 6.0 * dt;
                    a.rho[o] = rr;
                    a.ynext[o] = rr;
                }
            }
        }
}

// FP64 tensor-core variant (DMMA, mma.sync m8n8k4): same contract as qme_dense_stage.
// CTA tile 64x64 complex, 4 warps, each warp a 32x32 complex sub-tile = 4x4 MMA tiles of 8x8, four real
// MMAs per complex tile and k-step (C_re += A_re B_re + (-A_im) B_im, C_im += A_re B_im + A_im B_re).
// Operand tiles are stored planar (re / im) in shared memory, k-major, so that the m8n8k4 fragments
// (A: row = lane/4, k = lane%4; B: k = lane%4, col = lane/4) are single 8-byte loads.
__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

#define QDM_KT 8
#define QDM_LD 68        // 64 + 4 doubles of padding per k row: rows k, k+1, k+2, k+3 start 4 double-banks apart, so the 4 (k) x 4
                         // (row) doubles a half-warp reads for one MMA fragment fall into 16 distinct banks (72 made k and k+2 collide)
__global__ void __launch_bounds__(128, 2)
qme_dense_stage_dmma(QmeStageArgs a) {
    __shared__ __align__(16) double Ar[2][QDM_KT][QDM_LD], Ai[2][QDM_KT][QDM_LD];     // [buf][k][row]
    __shared__ __align__(16) double Br[2][QDM_KT][QDM_LD], Bi[2][QDM_KT][QDM_LD];     // [buf][k][col]
    const int N = a.N;
    const int Mr = a.Mr ? a.Mr : N, Nc = a.Nc ? a.Nc : N, Kd = a.Kd ? a.Kd : N;
    const int b = blockIdx.z;
    const int ti = blockIdx.y * 64, tj = blockIdx.x * 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;              // warp sub-tile origin
    const int fr = lane >> 2, fk = lane & 3;                            // fragment row/col and k index
    double cr[4][4][2], ci[4][4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { cr[r][c][0] = cr[r][c][1] = 0.0; ci[r][c][0] = ci[r][c][1] = 0.0; }
    // loaders: A tile 64 rows x 8 k (4 elements per thread), B tile 8 k x 64 cols (4 per thread)
    const int la_r = threadIdx.x >> 1, la_k = (threadIdx.x & 1) * 4;     // A: row, 4 consecutive k
    const int lb_k = threadIdx.x >> 4, lb_c = (threadIdx.x & 15) * 4;    // B: k row, 4 consecutive cols
    const int nk = (Kd + QDM_KT - 1) / QDM_KT;
    const int total = a.nprod * nk;
    cplx ra[4], rb[4];
    auto fetch = [&](int it) {
        const int p = it / nk, k0 = (it - p * nk) * QDM_KT;
        const cplx* Ap = a.A[p] + (size_t)b * a.sA[p];
        const cplx* Bp = a.Bm[p] + (size_t)b * a.sB[p];
        const int gi = ti + la_r, gkb = k0 + lb_k;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int gk = k0 + la_k + u, gj = tj + lb_c + u;
            ra[u] = (gi < Mr && gk < Kd) ? Ap[(size_t)gi * Kd + gk] : cmake(0, 0);
            rb[u] = (gkb < Kd && gj < Nc) ? Bp[(size_t)gkb * Nc + gj] : cmake(0, 0);
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { Ar[buf][la_k + u][la_r] = ra[u].x; Ai[buf][la_k + u][la_r] = ra[u].y; }
        // the 4 consecutive columns of a B row go out as two 16-byte stores per plane (scalar stores 32 bytes apart
        // between lanes were a 4-way bank conflict)
        *reinterpret_cast<double2*>(&Br[buf][lb_k][lb_c]) = make_double2(rb[0].x, rb[1].x);
        *reinterpret_cast<double2*>(&Br[buf][lb_k][lb_c + 2]) = make_double2(rb[2].x, rb[3].x);
        *reinterpret_cast<double2*>(&Bi[buf][lb_k][lb_c]) = make_double2(rb[0].y, rb[1].y);
        *reinterpret_cast<double2*>(&Bi[buf][lb_k][lb_c + 2]) = make_double2(rb[2].y, rb[3].y);
    };
    fetch(0);
    stash(0);
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int buf = it & 1;
        if (it + 1 < total) fetch(it + 1);
#pragma unroll
        for (int kb = 0; kb < QDM_KT / 4; ++kb) {
            double are[4], aim[4], ain[4], bre[4], bim[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                are[t] = Ar[buf][kb * 4 + fk][wr + t * 8 + fr];
                aim[t] = Ai[buf][kb * 4 + fk][wr + t * 8 + fr];
                ain[t] = -aim[t];
                bre[t] = Br[buf][kb * 4 + fk][wc + t * 8 + fr];
                bim[t] = Bi[buf][kb * 4 + fk][wc + t * 8 + fr];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    dmma8x8x4(cr[r][c][0], cr[r][c][1], are[r], bre[c]);
                    dmma8x8x4(cr[r][c][0], cr[r][c][1], ain[r], bim[c]);
                    dmma8x8x4(ci[r][c][0], ci[r][c][1], are[r], bim[c]);
                    dmma8x8x4(ci[r][c][0], ci[r][c][1], aim[r], bre[c]);
                }
        }
        if (it + 1 < total) stash(buf ^ 1);
        __syncthreads();
    }
    // epilogue: C fragment element (row = lane/4, cols = 2*(lane%4) + {0,1}) of every 8x8 tile
    const double dt = a.dt, hdt = 0.5 * a.dt;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gi = ti + wr + r * 8 + fr, gj = tj + wc + c * 8 + fk * 2 + e;
                if (gi >= Mr || gj >= Nc) continue;
                const size_t idx = (size_t)gi * Nc + gj;
                const cplx k = cmake(cr[r][c][e], ci[r][c][e]);
                if (a.mode == 0) {
                    a.out[(size_t)b * a.sOut + idx] = k;
                } else {
                    const size_t o = (size_t)b * Mr * Nc + idx;
                    cplx rr = a.rho[o];
                    if (a.mode == 1) {
                        a.acc[o] = k;
                        a.ynext[o] = cmake(fma(hdt, k.x, rr.x), fma(hdt, k.y, rr.y));
                    } else if (a.mode == 2) {
                        cplx ac = a.acc[o]; rfma(ac, 2.0, k); a.acc[o] = ac;
                        a.ynext[o] = cmake(fma(hdt, k.x, rr.x), fma(hdt, k.y, rr.y));
                    } else if (a.mode == 3) {
                        cplx ac = a.acc[o]; rfma(ac, 2.0, k); a.acc[o] = ac;
                        a.ynext[o] = cmake(fma(dt, k.x, rr.x), fma(dt, k.y, rr.y));
                    } else {
                        const cplx tot = cadd(a.acc[o], k);
                        rr.x += tot.x / 6.0 * dt;
                        rr.y += tot.y / 6.0 * dt;
                        a.rho[o] = rr;
                        a.ynext[o] = rr;
                    }
                }
            }
}

__global__ void __launch_bounds__(256)
qme_trace_obs(const cplx* __restrict__ eT, const cplx* __restrict__ rho, cplx* __restrict__ obs,
              int NN, int E, long long obs_stride_b) {
    __shared__ cplx red[8];
    const int e = blockIdx.x, b = blockIdx.y;
    const cplx* et = eT + (size_t)e * NN;
    const cplx* r = rho + (size_t)b * NN;
    cplx v = cmake(0, 0);
    for (int i = threadIdx.x; i < NN; i += 256) cfma(v, et[i], r[i]);
    for (int off = 16; off > 0; off >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, off);
        v.y += __shfl_down_sync(0xffffffffu, v.y, off);
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        cplx s = cmake(0, 0);
        for (int w = 0; w < 8; ++w) s = cadd(s, red[w]);
        obs[(size_t)b * obs_stride_b + e] = s;
    }
}
