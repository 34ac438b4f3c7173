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
#include <cstdlib>
namespace cgb = cooperative_groups;

#define QME_BAND_MAXS 2

// ---- mbarrier / st.async plumbing for the halo exchange ------------------------------------
// Halo rows are pushed into the neighbour CTA with st.async, which signals an mbarrier in the
// DESTINATION CTA as the bytes land; the consumer waits on its own mbarrier for the expected
// byte count.  No cluster-wide barrier, no fence (a cluster-scope release compiles to
// MEMBAR.ALL.GPU + CCTL.IVALL on sm_100a) in the time loop.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "LIMEB_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LIMEB_DONE_%=;\n\t"
        "bra LIMEB_WAIT_%=;\n"
        "LIMEB_DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
template <int IMM>
__device__ __forceinline__ void st_async_c128(unsigned raddr, cplx v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0+%4], {%1, %2}, [%3];"
                 ::"r"(raddr), "d"(v.x), "d"(v.y), "r"(rbar), "n"(IMM) : "memory");
}

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
    int debug_flags;             // bit 0: skip the halo exchange (timing experiments only; wrong results)
    int chain_rs;                // CH kernels: G couples row i to rows i - chain_rs (slot 0) and i + chain_rs (slot 1) only
};

// GT: 0 complex off-diagonal G, 1 purely imaginary.  XT: 0 complex X/Z, 1 real.
//
// Shared memory (element = double2, all indices 32-bit so that every access is an LDS/STS with
// a register base + immediate): two stage buffers of (R + 2h) rows x N (+128 pad: lanes whose
// column l + 32u lies beyond N read into the next row and are never stored), the reduction
// scratch, and the row-side coefficient tables of this CTA's rows.
// CH = 1 ("chain" structure, NOFF = 2): the off-diagonal part of G couples row i only to rows i -/+ RS, and a
// thread's TR rows are RS apart, so the left-hand neighbours of a row are the thread's previous / next row: a
// sliding window of three rows in registers replaces 3 of the 7 shared-memory loads per element by (TR + 2) / TR.
template <int TR, int TC, int NOFF, int S, int GT, int XT, int CH>
__global__ void __launch_bounds__(512, 1)
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
    const int CB = (N + 32 * TC - 1) / (32 * TC);      // column blocks of 32*TC columns; warp -> (row group, column block)
    const int jb = (warp % CB) * 32 * TC + lane;       // this thread's columns: jb + 32 u
    constexpr int SS = S > 0 ? S : 1;

    const int buf_elems = buf_rows * N + 128;
    const int o_buf1 = buf_elems;
    const int o_bar = 2 * buf_elems;                   // 4 mbarriers (8 B each): halo[2], partial sums[2]
    const int o_red = o_bar + 2;                       // [E][16]
    const int o_part = o_red + 16 * max(a.E, 1);        // [2][C][E] (rank 0)
    const int o_lgd = o_part + 2 * C * max(a.E, 1);    // [R]        G_ii
    const int o_lgv = o_lgd + R;                       // [NOFF][R]  G[i][c(i,q)]
    const int o_lxv = o_lgv + NOFF * R;                // [S][R]     X_s[i]
    int* lgo = reinterpret_cast<int*>(smem + o_lxv + SS * R);     // [NOFF][R] byte offset of row c(i,q) in a buffer
    int* lxo = lgo + NOFF * R;                                    // [S][R]
    cplx* part = smem + o_part;

    // ---- load rho (own + halo rows, permuted basis) into stage buffer 0
    const cplx* grho = a.rho + (size_t)b * N * N;
    for (int l = threadIdx.x; l < buf_elems; l += T) {
        int r = row0 + l / N, c = l % N;
        cplx v = cmake(0, 0);
        if (l < buf_rows * N && r >= 0 && r < N) {
            int gr = a.perm ? a.perm[r] : r, gc = a.perm ? a.perm[c] : c;
            v = grho[(size_t)gr * N + gc];
        }
        smem[l] = v;
        smem[o_buf1 + l] = cmake(0, 0);
    }
    // ---- row-side coefficient tables (rows beyond row_hi: zero coefficients, own row)
    for (int l = threadIdx.x; l < R; l += T) {
        const int i = row_lo + l;
        const bool ok = i < row_hi;
        smem[o_lgd + l] = ok ? a.gd[vb * N + i] : cmake(0, 0);
        for (int q = 0; q < NOFF; ++q) {
            smem[o_lgv + q * R + l] = ok ? a.gval[(vb * NOFF + q) * N + i] : cmake(0, 0);
            lgo[q * R + l] = ((ok ? a.gcol[q * N + i] : row_lo) - row0) * N * 16;
        }
        for (int s = 0; s < S; ++s) {
            smem[o_lxv + s * R + l] = ok ? a.xval[s][vb * N + i] : cmake(0, 0);
            lxo[s * R + l] = ((ok ? a.xcol[s][i] : row_lo) - row0) * N * 16;
        }
    }
    // ---- column-side coefficients (registers, whole launch)
    bool okc[TC];
    cplx gdj[TC];                    // conj(G_jj)
    int offR[NOFF][TC];              // byte offset of column c(j,q) inside a row
    cplx valR[NOFF][TC];             // conj(G[j][q])   (GT == 1: only .y is used)
    int offZ[SS][TC];
    cplx valZ[SS][TC];               // conj(Z[j])      (XT == 1: only .x is used)
#pragma unroll
    for (int u = 0; u < TC; ++u) {
        const int j = jb + 32 * u;
        okc[u] = j < N;
        const int jj = okc[u] ? j : 0;
        gdj[u] = okc[u] ? cconj(a.gd[vb * N + jj]) : cmake(0, 0);
#pragma unroll
        for (int q = 0; q < NOFF; ++q) {
            offR[q][u] = a.gcol[q * N + jj] * 16;
            valR[q][u] = okc[u] ? cconj(a.gval[(vb * NOFF + q) * N + jj]) : cmake(0, 0);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            offZ[s][u] = a.zcol[s][jj] * 16;
            valZ[s][u] = okc[u] ? cconj(a.zval[s][vb * N + jj]) : cmake(0, 0);
        }
    }
    __syncthreads();
    cplx rho[TR][TC], acc[TR][TC];
    const int RS = CH ? a.chain_rs : 1;                // distance between a thread's consecutive rows
    const int l0 = CH ? ((warp / CB) / RS) * (TR * RS) + (warp / CB) % RS
                      : (warp / CB) * TR;              // first own row, relative to row_lo
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int u = 0; u < TC; ++u) {
            const int i = row_lo + l0 + r * RS;
            rho[r][u] = (i < row_hi && okc[u]) ? smem[(l0 + r * RS + h) * N + jb + 32 * u] : cmake(0, 0);
            acc[r][u] = cmake(0, 0);
        }
    // ---- neighbours: mapped shared-memory windows and mbarriers
    const unsigned bar0 = smem_u32(smem + o_bar);      // halo mbarriers bar0 + 8*p, partial-sum mbarriers bar0 + 16 + 8*p
    const bool has_up = C > 1 && rank > 0 && !(a.debug_flags & 1), has_dn = C > 1 && rank < C - 1 && !(a.debug_flags & 1);
    unsigned up_base = 0, dn_base = 0, up_bar = 0, dn_bar = 0, r0_part = 0, r0_bar = 0, halo_bytes = 0;
    if (C > 1) {
        if (threadIdx.x == 0) {
            for (int m = 0; m < 4; ++m) mbar_init(bar0 + 8 * m, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        const unsigned sbase = smem_u32(smem);
        if (has_up) { up_base = mapa_u32(sbase, rank - 1); up_bar = mapa_u32(bar0, rank - 1); halo_bytes += h * N * 16; }
        if (has_dn) {
            dn_base = mapa_u32(sbase, rank + 1); dn_bar = mapa_u32(bar0, rank + 1);
            halo_bytes += min(h, min(N, row_lo + 2 * R) - (row_lo + R)) * N * 16;
        }
        r0_part = mapa_u32(smem_u32(smem + o_part), 0);
        r0_bar = mapa_u32(bar0 + 16, 0);
        cluster.sync();
    }

    const double dt = a.dt, hdt = 0.5 * a.dt, w6 = a.dt / 6.0;
    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const int yin = (stage & 1) ? o_buf1 : 0;
            const int yout = (stage & 1) ? 0 : o_buf1;
            const double cy = (stage == 2) ? dt : hdt;
            if (C > 1 && threadIdx.x == 0) mbar_arrive_expect_tx(bar0 + 8 * (stage & 1), halo_bytes);
            const char* yinb = reinterpret_cast<const char*>(smem + yin);          // stage input, row 0
            const char* yinl = yinb + jb * 16;                                      // ... at this thread's first column
            cplx wlo[TC], wcur[TC], whi[TC];                  // CH: rows li - RS, li, li + RS of the stage vector
            if (CH) {
#pragma unroll
                for (int u = 0; u < TC; ++u) {
                    wlo[u] = *reinterpret_cast<const cplx*>(yinl + (l0 - RS + h) * N * 16 + 512 * u);
                    wcur[u] = *reinterpret_cast<const cplx*>(yinl + (l0 + h) * N * 16 + 512 * u);
                }
            }
#pragma unroll
            for (int r = 0; r < TR; ++r) {
                const int li = l0 + r * RS;                       // row relative to row_lo (warp-uniform)
                if (row_lo + li < row_hi) {
                    const int ownoff = (li + h) * N * 16;
                    const char* ownb = yinb + ownoff;             // y[i][0]
                    const char* ownl = yinl + ownoff;             // y[i][jb]
                    const cplx gdi = smem[o_lgd + li];
                    cplx k[TC];
                    if (CH) {
#pragma unroll
                        for (int u = 0; u < TC; ++u) whi[u] = *reinterpret_cast<const cplx*>(ownl + RS * N * 16 + 512 * u);
                    }
#pragma unroll
                    for (int u = 0; u < TC; ++u) {
                        const cplx y = CH ? wcur[u] : *reinterpret_cast<const cplx*>(ownl + 512 * u);
                        const double dr = gdi.x + gdj[u].x, di = gdi.y + gdj[u].y;
                        k[u].x = dr * y.x - di * y.y;
                        k[u].y = dr * y.y + di * y.x;
                    }
#pragma unroll
                    for (int q = 0; q < NOFF; ++q) {
                        // left: G[i][c] y[c][j]  (c, value warp-uniform)
                        const char* lrow = yinl + lgo[q * R + li];
                        const cplx v = smem[o_lgv + q * R + li];
#pragma unroll
                        for (int u = 0; u < TC; ++u) {
                            const cplx y = CH ? (q == 0 ? wlo[u] : whi[u]) : *reinterpret_cast<const cplx*>(lrow + 512 * u);
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
                            const cplx y = *reinterpret_cast<const cplx*>(ownb + offR[q][u]);
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
                        const char* xrowb = yinb + lxo[s * R + li];
                        const cplx xv = smem[o_lxv + s * R + li];
#pragma unroll
                        for (int u = 0; u < TC; ++u) {
                            const cplx y = *reinterpret_cast<const cplx*>(xrowb + offZ[s][u]);
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
                    cplx yn[TC];
                    if (stage == 0) {
#pragma unroll
                        for (int u = 0; u < TC; ++u) {
                            acc[r][u] = k[u];
                            yn[u] = cmake(fma(cy, k[u].x, rho[r][u].x), fma(cy, k[u].y, rho[r][u].y));
                        }
                    } else if (stage < 3) {
#pragma unroll
                        for (int u = 0; u < TC; ++u) {
                            rfma(acc[r][u], 2.0, k[u]);
                            yn[u] = cmake(fma(cy, k[u].x, rho[r][u].x), fma(cy, k[u].y, rho[r][u].y));
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < TC; ++u) {
                            rho[r][u].x = fma(w6, acc[r][u].x + k[u].x, rho[r][u].x);
                            rho[r][u].y = fma(w6, acc[r][u].y + k[u].y, rho[r][u].y);
                            yn[u] = rho[r][u];
                        }
                    }
                    cplx* oo = smem + (yout + (li + h) * N + jb);
#pragma unroll
                    for (int u = 0; u < TC; ++u)
                        if (okc[u]) oo[32 * u] = yn[u];
                    // own row li sits at local row li + R + h in the upper neighbour's buffer and at
                    // li - R + h in the lower neighbour's; the bytes are counted on the neighbour's mbarrier
                    if (has_up && li < h) {
                        const unsigned d = up_base + (unsigned)(yout + (li + R + h) * N + jb) * 16u;
                        const unsigned rb = up_bar + 8 * (stage & 1);
                        if (okc[0]) st_async_c128<0>(d, yn[0], rb);
                        if (TC > 1 && okc[TC > 1 ? 1 : 0]) st_async_c128<512>(d, yn[TC > 1 ? 1 : 0], rb);
                        if (TC > 2 && okc[TC > 2 ? 2 : 0]) st_async_c128<1024>(d, yn[TC > 2 ? 2 : 0], rb);
                        if (TC > 3 && okc[TC > 3 ? 3 : 0]) st_async_c128<1536>(d, yn[TC > 3 ? 3 : 0], rb);
                    }
                    if (has_dn && li >= R - h) {
                        const unsigned d = dn_base + (unsigned)(yout + (li - R + h) * N + jb) * 16u;
                        const unsigned rb = dn_bar + 8 * (stage & 1);
                        if (okc[0]) st_async_c128<0>(d, yn[0], rb);
                        if (TC > 1 && okc[TC > 1 ? 1 : 0]) st_async_c128<512>(d, yn[TC > 1 ? 1 : 0], rb);
                        if (TC > 2 && okc[TC > 2 ? 2 : 0]) st_async_c128<1024>(d, yn[TC > 2 ? 2 : 0], rb);
                        if (TC > 3 && okc[TC > 3 ? 3 : 0]) st_async_c128<1536>(d, yn[TC > 3 ? 3 : 0], rb);
                    }
                    if (CH) {
#pragma unroll
                        for (int u = 0; u < TC; ++u) { wlo[u] = wcur[u]; wcur[u] = whi[u]; }
                    }
                }
            }
            // local writes visible / local reads of the old stage vector finished ...
            __syncthreads();
            // ... and the neighbours' halo rows have landed (phase (g >> 1) & 1 of mbarrier g & 1, g = 4 step + stage)
            if (C > 1) mbar_wait(bar0 + 8 * (stage & 1), (unsigned)((step * 2 + (stage >> 1)) & 1));
            if (C > 1 && stage == 0 && a.obs && step > 0 && rank == 0) {
                // partial sums of the previous step (pushed by the other CTAs with st.async) are complete
                const int sp = step - 1;
                for (int e = threadIdx.x; e < a.E; e += T) {
                    mbar_wait(bar0 + 16 + 8 * (sp & 1), (unsigned)((sp >> 1) & 1));
                    const cplx* pp = part + (size_t)(sp & 1) * C * a.E;
                    cplx sum = cmake(0, 0);
                    for (int c = 0; c < C; ++c) sum = cadd(sum, pp[c * a.E + e]);
                    a.obs[((size_t)sp * a.B + b) * a.E + e] = sum;
                }
            }
        }
        // buffer 0 now holds rho_{n+1} (own + halo rows)
        if (a.obs) {
            for (int e = 0; e < a.E; ++e) {
                cplx v = cmake(0, 0);
                for (int n = a.eptr[e] + threadIdx.x; n < a.eptr[e + 1]; n += T) {
                    int idx = a.eidx[n];
                    int i = idx / N;
                    if (i >= row_lo && i < row_hi) cfma(v, a.eval[n], smem[(i - row0) * N + (idx - i * N)]);
                }
                for (int off = 16; off > 0; off >>= 1) {
                    v.x += __shfl_down_sync(0xffffffffu, v.x, off);
                    v.y += __shfl_down_sync(0xffffffffu, v.y, off);
                }
                if (lane == 0) smem[o_red + e * 16 + warp] = v;
            }
            __syncthreads();
            if (C > 1 && rank == 0 && threadIdx.x == 0)
                mbar_arrive_expect_tx(bar0 + 16 + 8 * (step & 1), (unsigned)((C - 1) * a.E * 16));
            for (int e = threadIdx.x; e < a.E; e += T) {
                cplx sum = cmake(0, 0);
                for (int w = 0; w < (T + 31) / 32; ++w) sum = cadd(sum, smem[o_red + e * 16 + w]);
                if (C == 1) a.obs[((size_t)step * a.B + b) * a.E + e] = sum;
                else if (rank == 0) part[(size_t)(step & 1) * C * a.E + e] = sum;
                else st_async_c128<0>(r0_part + (unsigned)(((step & 1) * C + rank) * a.E + e) * 16u, sum,
                                      r0_bar + 8 * (step & 1));
            }
            __syncthreads();          // red[] is reused by the next step
        }
        if (a.traj && ((step + 1) % a.traj_every) == 0) {
            cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * N * N;
#pragma unroll
            for (int r = 0; r < TR; ++r)
#pragma unroll
                for (int u = 0; u < TC; ++u) {
                    const int i = row_lo + l0 + r * RS, j = jb + 32 * u;
                    if (i < row_hi && okc[u]) {
                        int gr = a.perm ? a.perm[i] : i, gc = a.perm ? a.perm[j] : j;
                        dst[(size_t)gr * N + gc] = rho[r][u];
                    }
                }
        }
    }
    if (C > 1 && a.obs && a.nsteps > 0 && rank == 0) {
        const int sp = a.nsteps - 1;
        for (int e = threadIdx.x; e < a.E; e += T) {
            mbar_wait(bar0 + 16 + 8 * (sp & 1), (unsigned)((sp >> 1) & 1));
            const cplx* pp = part + (size_t)(sp & 1) * C * a.E;
            cplx sum = cmake(0, 0);
            for (int c = 0; c < C; ++c) sum = cadd(sum, pp[c * a.E + e]);
            a.obs[((size_t)sp * a.B + b) * a.E + e] = sum;
        }
    }
    cplx* out = a.rho + (size_t)b * N * N;
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int u = 0; u < TC; ++u) {
            const int i = row_lo + l0 + r * RS, j = jb + 32 * u;
            if (i < row_hi && okc[u]) {
                int gr = a.perm ? a.perm[i] : i, gc = a.perm ? a.perm[j] : j;
                out[(size_t)gr * N + gc] = rho[r][u];
            }
        }
    if (C > 1) cluster.sync();      // keep shared memory alive until the neighbours' remote stores are done
}

// shared-memory bytes of qme_band_kernel for a given geometry
static inline size_t qme_band_smem(int N, int R, int h, int C, int E, int NOFF, int S) {
    const size_t SS = S > 0 ? S : 1;
    size_t elems = (size_t)2 * ((size_t)(R + 2 * h) * N + 128) + 2 + (size_t)(16 + 2 * C) * (E > 0 ? E : 1) +
                   (size_t)R * (1 + NOFF + SS);
    return elems * 16 + (size_t)R * (NOFF + SS) * 4 + 16;
}

// geometry: smallest cluster with R = ceil(N/C) <= 32 rows per CTA that fits shared memory
static inline bool qme_band_geometry(int N, int E, int bandwidth, int NOFF, int S, long long smem_optin,
                                     int* Cout, int* Rout, size_t* smem_out) {
    if (N > 128) return false;
    const int h = bandwidth;
    const char* force = getenv("LIMEB200_BAND_C");           // tuning/experiments: force the cluster size
    for (int c = 1; c <= 8; c *= 2) {
        int r = (N + c - 1) / c;
        if (r > 32) continue;
        if (force && atoi(force) != c) continue;
        if (c > 1 && h > r) break;
        size_t need = qme_band_smem(N, r, h, c, E, NOFF, S);
        if (need <= (size_t)smem_optin) { *Cout = c; *Rout = r; *smem_out = need; return true; }
    }
    return false;
}

// implemented in qme_band_inst_*.cu (one translation unit per (TC, NOFF) so that make -j compiles them in parallel)
int qme_band_launch_tc2_n2(const QmeBandArgs& a, int S, int GT, int XT, size_t smem, cudaStream_t st);
int qme_band_launch_tc2_n4(const QmeBandArgs& a, int S, int GT, int XT, size_t smem, cudaStream_t st);

template <int TC, int NOFF, int S, int GT, int XT, int CH = 0>
static int qme_band_launch_one(const QmeBandArgs& a, size_t smem, cudaStream_t st) {
    constexpr int TR = 4;
    auto kern = qme_band_kernel<TR, TC, NOFF, S, GT, XT, CH>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int W = ((a.R + TR - 1) / TR) * ((a.N + 32 * TC - 1) / (32 * TC));
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

// chain-structured variants (NOFF = 2, purely imaginary off-diagonal G, real X/Z: the ladder-operator case)
#define QME_BAND_DEFINE_LAUNCH_CHAIN(NAME, TC)                                                        \
    int NAME(const QmeBandArgs& a, int S, size_t smem, cudaStream_t st) {                             \
        if (S == 0) return qme_band_launch_one<TC, 2, 0, 1, 1, 1>(a, smem, st);                       \
        if (S == 1) return qme_band_launch_one<TC, 2, 1, 1, 1, 1>(a, smem, st);                       \
        return qme_band_launch_one<TC, 2, 2, 1, 1, 1>(a, smem, st);                                   \
    }
int qme_band_launch_chain_tc2(const QmeBandArgs& a, int S, size_t smem, cudaStream_t st);
