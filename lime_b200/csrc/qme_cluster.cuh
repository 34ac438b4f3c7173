// Cluster-resident sparse (ELL) RK4 propagator: one thread-block cluster per density matrix.
//
//   * rows of rho are split across the C CTAs of the cluster (C = 1,2,4,8);
//   * the two ping-pong stage vectors (own rows + `halo` rows above/below) live in shared
//     memory, rho_n and the RK4 accumulator in registers, the ELL operators in shared memory;
//   * after each stage the boundary rows of the new stage vector are pushed into the
//     neighbour CTAs' halo regions through distributed shared memory, then ONE cluster
//     barrier per stage;
//   * all nsteps steps run inside one launch: HBM traffic is rho in, rho out, observables.
//
// The host permutes the basis (reverse Cuthill-McKee) so that the operators are banded;
// `perm[new] = old` maps back on load/store so the caller never sees the permutation.
#pragma once
#include "common.cuh"
#include "qme_sparse.cuh"

struct QmeClusterGeom {
    int C;          // CTAs per cluster
    int R;          // rows per CTA (ceil(N / C))
    int h;          // halo rows
    int T;          // threads per CTA
    const int* perm;   // [N] new -> old, or null
};


template <int EPT>
__global__ void __launch_bounds__(512, 1)
qme_ell_cluster(QmeEllArgs a, QmeClusterGeom g) {
    extern __shared__ double2 smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int N = a.N, R = g.R, h = g.h, C = g.C, T = g.T;
    const int rank = (C > 1) ? (int)cluster.block_rank() : 0;
    const int b = blockIdx.x / C;
    const size_t vb = (a.nb > 1) ? (size_t)b : 0;
    const int row_lo = rank * R;
    const int row_hi = min(N, row_lo + R);
    const int buf_rows = R + 2 * h;
    const int row0 = row_lo - h;                       // global row of local row 0

    // ---- shared memory carve-up
    cplx* ybuf0 = smem;
    cplx* ybuf1 = ybuf0 + (size_t)buf_rows * N;
    cplx* red = ybuf1 + (size_t)buf_rows * N;          // [32]
    cplx* part = red + 32;                             // [2][C][E]  (used on rank 0)
    cplx* vals = part + 2 * C * max(a.E, 1);
    QmeEllArgs s = a;                                  // operator views redirected to smem
    {
        cplx* v = vals;
        auto stage_vals = [&](EllOp& op) {
            const cplx* src = op.val + vb * N * op.w;
            for (int i = threadIdx.x; i < N * op.w; i += T) v[i] = src[i];
            op.val = v;
            v += N * op.w;
        };
        stage_vals(s.G);
        for (int sw = 0; sw < a.S; ++sw) { stage_vals(s.X[sw]); stage_vals(s.Z[sw]); }
        int* c = reinterpret_cast<int*>(v);
        auto stage_cols = [&](EllOp& op) {
            for (int i = threadIdx.x; i < N * op.w; i += T) c[i] = op.col[i];
            op.col = c;
            c += N * op.w;
        };
        stage_cols(s.G);
        for (int sw = 0; sw < a.S; ++sw) { stage_cols(s.X[sw]); stage_cols(s.Z[sw]); }
        s.nb = 1;
    }

    // ---- load rho: own rows into registers, own + halo rows into stage buffer 0
    const cplx* grho = a.rho + (size_t)b * N * N;
    for (int l = threadIdx.x; l < buf_rows * N; l += T) {
        int r = row0 + l / N, c = l % N;
        cplx v = cmake(0, 0);
        if (r >= 0 && r < N) {
            int gr = g.perm ? g.perm[r] : r, gc = g.perm ? g.perm[c] : c;
            v = grho[(size_t)gr * N + gc];
        }
        ybuf0[l] = v;
    }
    int ei[EPT], ej[EPT];
    bool ok[EPT];
    cplx rho[EPT], acc[EPT];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        int idx = threadIdx.x + e * T;
        int i = row_lo + idx / N;
        ok[e] = (idx < R * N) && (i < row_hi);
        ei[e] = ok[e] ? i : row_lo;
        ej[e] = ok[e] ? idx % N : 0;
        rho[e] = ok[e] ? ybuf0[(size_t)(ei[e] - row0) * N + ej[e]] : cmake(0, 0);
        acc[e] = cmake(0, 0);
    }
    // remote halo targets
    cplx* up0 = nullptr; cplx* up1 = nullptr; cplx* dn0 = nullptr; cplx* dn1 = nullptr;
    if (C > 1) {
        if (rank > 0) { up0 = cluster.map_shared_rank(ybuf0, rank - 1); up1 = cluster.map_shared_rank(ybuf1, rank - 1); }
        if (rank < C - 1) { dn0 = cluster.map_shared_rank(ybuf0, rank + 1); dn1 = cluster.map_shared_rank(ybuf1, rank + 1); }
        cluster.sync();
    }
    cplx* part0 = (C > 1) ? cluster.map_shared_rank(part, 0) : part;

    const double dt = a.dt, hdt = 0.5 * a.dt;
    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const cplx* yin = (stage & 1) ? ybuf1 : ybuf0;
            cplx* yout = (stage & 1) ? ybuf0 : ybuf1;
            cplx* rup = (stage & 1) ? up0 : up1;
            cplx* rdn = (stage & 1) ? dn0 : dn1;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                if (!ok[e]) continue;
                const int i = ei[e], j = ej[e];
                cplx k = ell_rhs_elem<false>(s, 0, i, j, yin, N, row0);
                cplx yn;
                if (stage == 0) {
                    acc[e] = k;
                    yn = cmake(fma(hdt, k.x, rho[e].x), fma(hdt, k.y, rho[e].y));
                } else if (stage == 1) {
                    rfma(acc[e], 2.0, k);
                    yn = cmake(fma(hdt, k.x, rho[e].x), fma(hdt, k.y, rho[e].y));
                } else if (stage == 2) {
                    rfma(acc[e], 2.0, k);
                    yn = cmake(fma(dt, k.x, rho[e].x), fma(dt, k.y, rho[e].y));
                } else {
                    cplx tot = cadd(acc[e], k);
                    rho[e].x += tot.x / 6.0 * dt;
                    rho[e].y += tot.y / 6.0 * dt;
                    yn = rho[e];
                }
                const int lr = i - row0;                       // local row in own buffer
                yout[(size_t)lr * N + j] = yn;
                // own row i sits at local row (i - row_lo) + R + h in the upper neighbour's
                // buffer and at (i - row_lo) - R + h in the lower neighbour's
                if (rup && i - row_lo < h) rup[(size_t)(i - row_lo + R + h) * N + j] = yn;
                if (rdn && i >= row_lo + R - h) rdn[(size_t)(i - row_lo - R + h) * N + j] = yn;
            }
            if (C > 1) cluster.sync(); else __syncthreads();
            if (stage == 0 && a.obs && step > 0 && rank == 0 && threadIdx.x < a.E) {
                // partials of the previous step are complete (they were pushed before this barrier)
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
                cplx sum = block_reduce_cplx(v, red);
                if (threadIdx.x == 0) part0[(size_t)(step & 1) * C * a.E + rank * a.E + e] = sum;
            }
        }
        if (a.traj && ((step + 1) % a.traj_every) == 0) {
            cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * N * N;
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (ok[e]) {
                    int gr = g.perm ? g.perm[ei[e]] : ei[e], gc = g.perm ? g.perm[ej[e]] : ej[e];
                    dst[(size_t)gr * N + gc] = rho[e];
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
    for (int e = 0; e < EPT; ++e)
        if (ok[e]) {
            int gr = g.perm ? g.perm[ei[e]] : ei[e], gc = g.perm ? g.perm[ej[e]] : ej[e];
            out[(size_t)gr * N + gc] = rho[e];
        }
    // keep this CTA's shared memory alive until every neighbour has finished its remote stores
    if (C > 1) cluster.sync();
}

// cluster geometry for a given operator set; false when the problem does not fit
static bool qme_cluster_geometry(int N, int S, int wG, const int* wX, const int* wZ, int E, int bandwidth,
                                 long long smem_optin, int* Cout, int* Rout, size_t* smem_out) {
    size_t op_elems = (size_t)N * wG;
    for (int s = 0; s < S; ++s) op_elems += (size_t)N * (wX[s] + wZ[s]);
    const size_t op_bytes = op_elems * (16 + 4) + 16;
    const int h = bandwidth;
    for (int c = 1; c <= 8; c *= 2) {
        int r = ceil_div(N, c);
        if (c > 1 && h > r) break;                  // halo would reach beyond the adjacent CTA
        if ((long long)r * N > 4096) continue;      // 512 threads x 8 elements
        size_t need = (size_t)2 * (r + 2 * h) * N * 16 + (32 + 2 * c * std::max(E, 1)) * 16 + op_bytes;
        if (need <= (size_t)smem_optin) { *Cout = c; *Rout = r; *smem_out = need; return true; }
    }
    return false;
}

static int qme_cluster_launch(QmeEllArgs a, int bandwidth, const int* d_perm, long long smem_optin,
                              cudaStream_t st) {
    const int N = a.N;
    int wX[QME_MAXS], wZ[QME_MAXS];
    for (int s = 0; s < a.S; ++s) { wX[s] = a.X[s].w; wZ[s] = a.Z[s].w; }
    int C = 0, R = 0;
    size_t smem = 0;
    if (!qme_cluster_geometry(N, a.S, a.G.w, wX, wZ, a.E, bandwidth, smem_optin, &C, &R, &smem))
        return LB_ERR_UNSUPPORTED;
    const int elems = R * N;
    int T = std::min(512, ceil_div(elems, 32) * 32);
    int ept = ceil_div(elems, T);
    int EPT = ept <= 1 ? 1 : (ept <= 2 ? 2 : (ept <= 4 ? 4 : 8));
    QmeClusterGeom g;
    g.C = C; g.R = R; g.h = bandwidth; g.T = T; g.perm = d_perm;
    void (*kern)(QmeEllArgs, QmeClusterGeom) =
        EPT == 1 ? qme_ell_cluster<1> : EPT == 2 ? qme_ell_cluster<2> : EPT == 4 ? qme_ell_cluster<4> : qme_ell_cluster<8>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.B * C);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    LB_CUDA(cudaLaunchKernelEx(&cfg, kern, a, g));
    return LB_OK;
}
