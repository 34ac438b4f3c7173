// Register-patch cluster propagator for chain-structured operators (Jaynes-Cummings / Rabi
// ladders, tight-binding chains with ladder-type collapse operators):
//
//     d rho/dt = G rho + rho G^H + sum_s X_s rho Z_s^H        (lime/oqs.py:706-723 in generator
//     form; RK4 of lime/phys.py:636-649, all nsteps of a launch fused)
//
// Structure exploited (checked on the host, qme.cu:build_tile_host; anything else runs on
// qme_band_kernel): in a suitable basis order ("chain order": the off-diagonal graph of G is a
// union of P simple paths) G is tridiagonal with purely imaginary off-diagonals and every X_s /
// Z_s has one real entry per row.  The propagation is then a 5-point stencil on rho plus one
// shifted pick-up per sandwich term, and the kernel is organised like a stencil code:
//
//   * one thread-block cluster per density matrix; CTA c owns, from each path, the chunk of
//     `chunk` consecutive chain rows [c*chunk, (c+1)*chunk) and keeps them together with one halo
//     row on either side of every chunk: buffer row (p, t), t = 0 .. chunk+1;
//   * a thread owns a PATCH of TR chain-consecutive rows x 2 chain-consecutive columns.  Columns
//     are stored in a lane-interleaved order (chain column q = 64 cb + 2 lane + u sits at position
//     64 cb + 32 u + lane) so that the patch, its left and its right neighbour columns are each one
//     conflict-free 512-byte warp access; the TR + 2 rows of the patch's row window slide through
//     registers.  Per element and stage that is 4.5 shared-memory accesses of 16 B (window
//     (TR+2)/TR, left/right 1, sandwich source 1, store 1) against 8 in qme_band_kernel, and no
//     address arithmetic: every access is base register + immediate;
//   * rho, the RK4 accumulator and the row-side coefficients live in TENSOR MEMORY (tcgen05.ld /
//     tcgen05.st, 32x32b shape: TMEM lane = thread, columns = the thread's private words), used
//     as a second register file: 32 words per patch row = [rho 8][acc 8][G_ii, Im G_i,i-1,
//     Im G_i,i+1, X_0, X_1: 12][pad 4], fetched with one tcgen05.ld.x32 per row and stage.  That
//     frees 8 TR registers per thread for the sliding window and takes the 3 coefficient loads
//     per row (LDS.128 broadcasts, 2 wavefronts each) off the shared-memory pipe;
//   * the two stage vectors ping-pong in shared memory; halo rows are pushed into the neighbour
//     CTAs with st.async (complete_tx on the consumer's mbarrier), one __syncthreads + one
//     mbarrier wait per stage, no cluster barrier in the time loop;
//   * the four RK4 stages are unrolled with compile-time buffer addresses and stage algebra.
#pragma once
#include "common.cuh"
#include "qme_band.cuh"          // mbarrier / st.async helpers
#include <cooperative_groups.h>

#define QME_TILE_MAXS 2
#define QME_TILE_ROWC 6           // doubles per row-coefficient record: gd.x gd.y gup gdn xv0 xv1
#define QME_TILE_COLC 6           // doubles per column-coefficient record: conj(gd).x conj(gd).y cL cR zv0 zv1

struct QmeTileArgs {
    int N, E, B, nsteps, traj_every, nb;
    int C, P, chunk;             // CTAs per cluster, paths, chain rows of one path owned by one CTA
    const int* brow;             // [C][P*(chunk+2)]  old basis index of buffer row (p, t), -1 = zero row
    const int* colold;           // [NP]              old basis index of column position, -1 = padding
    const int* colpos;           // [NP][4]           byte offsets inside a row: left nbr, right nbr, sandwich source 0, 1
    const double* colc;          // [nb][NP][QME_TILE_COLC]
    const double* rowc;          // [nb][C][R][QME_TILE_ROWC],  R = P*chunk own rows per CTA
    const int* xs0;              // [C][R/TR][QME_TILE_MAXS]  buffer row of the sandwich source of a patch's first row
    const int* eptr;             // observables as COO: entries [eptr[e], eptr[e+1])
    const int* erank;            //   owner CTA of the entry's row
    const int* eoff;             //   element offset (buffer row * NP + position) in the owner's buffer
    const cplx* eval;
    cplx* rho;                   // [B][N][N] in/out (caller's basis order)
    cplx* obs;                   // [nsteps][B][E] or null
    cplx* traj;                  // [nsteps/traj_every][B][N][N] or null
    double dt;
};

// ---- tensor memory as a per-thread scratch file ------------------------------------------------
#define QME_TILE_TMW 32           // tensor-memory words per patch row: [rho 8][acc 8][coefficients 12][pad 4]
struct TmemRow { unsigned r[32]; };
// issue the load of a patch row's 32 words (asynchronous) ...
__device__ __forceinline__ void tmem_ld32_issue(unsigned taddr, TmemRow& w) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(w.r[0]), "=r"(w.r[1]), "=r"(w.r[2]), "=r"(w.r[3]), "=r"(w.r[4]), "=r"(w.r[5]), "=r"(w.r[6]), "=r"(w.r[7]),
          "=r"(w.r[8]), "=r"(w.r[9]), "=r"(w.r[10]), "=r"(w.r[11]), "=r"(w.r[12]), "=r"(w.r[13]), "=r"(w.r[14]), "=r"(w.r[15]),
          "=r"(w.r[16]), "=r"(w.r[17]), "=r"(w.r[18]), "=r"(w.r[19]), "=r"(w.r[20]), "=r"(w.r[21]), "=r"(w.r[22]), "=r"(w.r[23]),
          "=r"(w.r[24]), "=r"(w.r[25]), "=r"(w.r[26]), "=r"(w.r[27]), "=r"(w.r[28]), "=r"(w.r[29]), "=r"(w.r[30]), "=r"(w.r[31])
        : "r"(taddr));
}
// ... and wait for it; the words are operands of the wait so that no consumer can be scheduled above it
__device__ __forceinline__ void tmem_ld32_wait(TmemRow& w) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(w.r[0]), "+r"(w.r[1]), "+r"(w.r[2]), "+r"(w.r[3]), "+r"(w.r[4]), "+r"(w.r[5]), "+r"(w.r[6]), "+r"(w.r[7]),
          "+r"(w.r[8]), "+r"(w.r[9]), "+r"(w.r[10]), "+r"(w.r[11]), "+r"(w.r[12]), "+r"(w.r[13]), "+r"(w.r[14]), "+r"(w.r[15]),
          "+r"(w.r[16]), "+r"(w.r[17]), "+r"(w.r[18]), "+r"(w.r[19]), "+r"(w.r[20]), "+r"(w.r[21]), "+r"(w.r[22]), "+r"(w.r[23]),
          "+r"(w.r[24]), "+r"(w.r[25]), "+r"(w.r[26]), "+r"(w.r[27])
        :: "memory");
}
// Clobber-free forms used by the tensor-memory-window variant (tensor memory does not alias shared memory; the
// TMEM accesses stay ordered among themselves because they are volatile).
__device__ __forceinline__ void tmem_st8_nc(unsigned taddr, const double (&d)[4]) {
    unsigned r[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { r[2 * i] = (unsigned)__double2loint(d[i]); r[2 * i + 1] = (unsigned)__double2hiint(d[i]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
template <int IMM>
__device__ __forceinline__ void st_async_c128_nc(unsigned raddr, cplx v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0+%4], {%1, %2}, [%3];"
                 ::"r"(raddr), "d"(v.x), "d"(v.y), "r"(rbar), "n"(IMM));
}
// 16-word load (rho + accumulator)
struct TmemRow16 { unsigned r[16]; };
__device__ __forceinline__ void tmem_ld16_issue(unsigned taddr, TmemRow16& w) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(w.r[0]), "=r"(w.r[1]), "=r"(w.r[2]), "=r"(w.r[3]), "=r"(w.r[4]), "=r"(w.r[5]), "=r"(w.r[6]), "=r"(w.r[7]),
          "=r"(w.r[8]), "=r"(w.r[9]), "=r"(w.r[10]), "=r"(w.r[11]), "=r"(w.r[12]), "=r"(w.r[13]), "=r"(w.r[14]), "=r"(w.r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_wait(TmemRow16& w) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(w.r[0]), "+r"(w.r[1]), "+r"(w.r[2]), "+r"(w.r[3]), "+r"(w.r[4]), "+r"(w.r[5]), "+r"(w.r[6]), "+r"(w.r[7]),
          "+r"(w.r[8]), "+r"(w.r[9]), "+r"(w.r[10]), "+r"(w.r[11]), "+r"(w.r[12]), "+r"(w.r[13]), "+r"(w.r[14]), "+r"(w.r[15]));
}
__device__ __forceinline__ double tmem_dbl(const TmemRow16& w, int i) {
    return __hiloint2double((int)w.r[2 * i + 1], (int)w.r[2 * i]);
}
// 8-word load (one patch row of the stage vector)
struct TmemRow8 { unsigned r[8]; };
__device__ __forceinline__ void tmem_ld8_issue(unsigned taddr, TmemRow8& w) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w.r[0]), "=r"(w.r[1]), "=r"(w.r[2]), "=r"(w.r[3]), "=r"(w.r[4]), "=r"(w.r[5]), "=r"(w.r[6]), "=r"(w.r[7])
        : "r"(taddr));
}
// wait for every outstanding tensor-memory load of the thread; both records are operands so that no consumer moves up
__device__ __forceinline__ void tmem_ld_wait(TmemRow16& a, TmemRow8& b) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(a.r[4]), "+r"(a.r[5]), "+r"(a.r[6]), "+r"(a.r[7]),
          "+r"(a.r[8]), "+r"(a.r[9]), "+r"(a.r[10]), "+r"(a.r[11]), "+r"(a.r[12]), "+r"(a.r[13]), "+r"(a.r[14]), "+r"(a.r[15]),
          "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]), "+r"(b.r[4]), "+r"(b.r[5]), "+r"(b.r[6]), "+r"(b.r[7]));
}
__device__ __forceinline__ void tmem_ld_wait(TmemRow8& b) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]), "+r"(b.r[4]), "+r"(b.r[5]), "+r"(b.r[6]), "+r"(b.r[7]));
}
__device__ __forceinline__ double tmem_dbl(const TmemRow8& w, int i) {
    return __hiloint2double((int)w.r[2 * i + 1], (int)w.r[2 * i]);
}
__device__ __forceinline__ double tmem_dbl(const TmemRow& w, int i) {      // i-th double of the row record
    return __hiloint2double((int)w.r[2 * i + 1], (int)w.r[2 * i]);
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const double (&d)[4]) {
    unsigned r[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { r[2 * i] = (unsigned)__double2loint(d[i]); r[2 * i + 1] = (unsigned)__double2hiint(d[i]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

template <int NP, int TR, int S>
struct QmeTileCtx {
    static constexpr int ROWB = NP * 16;
    // per-thread constants of the launch
    unsigned own, nl, nr;              // byte offset (buffer 0, from the start of shared memory) of window row 0 at: own column 0, left nbr, right nbr
    unsigned xs[S > 0 ? S : 1][2];     // ... of the sandwich source of patch row 0 at the source columns of u = 0, 1
    unsigned up_dst, dn_dst, up_bar, dn_bar;   // remote (shared::cluster) addresses, buffer 0, of the pushed rows; 0 = no push
    unsigned trho;                     // tensor-memory address of this thread's row records
    unsigned rowc;                     // TW: byte offset (from the start of shared memory) of the patch's row coefficients
    double cdr[2], cdi[2], cL[2], cR[2], zv[S > 0 ? S : 1][2];
    double hdt, dt, w6;
};

// One RK4 stage of one thread's patch.  STAGE is compile time: input buffer = STAGE & 1, output = the other one.
template <int NP, int TR, int S, int STAGE>
__device__ __forceinline__ void qme_tile_stage(const QmeTileCtx<NP, TR, S>& c, char* smem, unsigned bufb) {
    constexpr int ROWB = NP * 16;
    // ordinary shared-memory accesses (base register + immediate) that the compiler is free to schedule
    const unsigned in_off = (STAGE & 1) ? bufb : 0u;
    const unsigned out_off = (STAGE & 1) ? 0u : bufb;
    const char* pown = smem + (c.own + in_off);
    const char* pl = smem + (c.nl + in_off);
    const char* pr = smem + (c.nr + in_off);
    char* pout = smem + (c.own + out_off);
    const double cy = (STAGE == 2) ? c.dt : c.hdt;

    cplx wp[2], wc[2], wn[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        wp[u] = *reinterpret_cast<const cplx*>(pown + 512 * u);
        wc[u] = *reinterpret_cast<const cplx*>(pown + ROWB + 512 * u);
    }
#pragma unroll
    for (int r = 0; r < TR; ++r) {
        TmemRow tw;
        tmem_ld32_issue(c.trho + QME_TILE_TMW * r, tw);
#pragma unroll
        for (int u = 0; u < 2; ++u) wn[u] = *reinterpret_cast<const cplx*>(pown + (r + 2) * ROWB + 512 * u);
        const cplx yl = *reinterpret_cast<const cplx*>(pl + (r + 1) * ROWB);
        const cplx yr = *reinterpret_cast<const cplx*>(pr + (r + 1) * ROWB);
        cplx k[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            // right multiplication: y[i][j-1] conj(G[j][j-1]) + y[i][j+1] conj(G[j][j+1]), coefficients i cL, i cR
            const cplx a = (u == 0) ? yl : wc[0];
            const cplx b = (u == 0) ? wc[1] : yr;
            k[u].x = -c.cL[u] * a.y;
            k[u].y = c.cL[u] * a.x;
            k[u].x = fma(-c.cR[u], b.y, k[u].x);
            k[u].y = fma(c.cR[u], b.x, k[u].y);
        }
        tmem_ld32_wait(tw);
        const double gdx = tmem_dbl(tw, 8), gdy = tmem_dbl(tw, 9), gup = tmem_dbl(tw, 10), gdn = tmem_dbl(tw, 11);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            // diagonal: (G_ii + conj G_jj) y_ij
            const double dr = gdx + c.cdr[u], di = gdy + c.cdi[u];
            k[u].x = fma(dr, wc[u].x, k[u].x);
            k[u].x = fma(-di, wc[u].y, k[u].x);
            k[u].y = fma(dr, wc[u].y, k[u].y);
            k[u].y = fma(di, wc[u].x, k[u].y);
            // left multiplication: i gup y[i-1][j] + i gdn y[i+1][j]
            k[u].x = fma(-gup, wp[u].y, k[u].x);
            k[u].y = fma(gup, wp[u].x, k[u].y);
            k[u].x = fma(-gdn, wn[u].y, k[u].x);
            k[u].y = fma(gdn, wn[u].x, k[u].y);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double xv = tmem_dbl(tw, 12 + s);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const cplx ys = *reinterpret_cast<const cplx*>(smem + (c.xs[s][u] + in_off) + r * ROWB);
                const double cf = xv * c.zv[s][u];
                k[u].x = fma(cf, ys.x, k[u].x);
                k[u].y = fma(cf, ys.y, k[u].y);
            }
        }
        // RK4 stage algebra (lime/phys.py:636-649); record words 0-3 = rho, 4-7 = accumulator
        cplx yn[2];
        double st4[4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const double kx = k[u].x, ky = k[u].y;
            if (STAGE == 0) {                     // the stage input IS rho
                st4[2 * u] = kx; st4[2 * u + 1] = ky;
                yn[u] = cmake(fma(cy, kx, wc[u].x), fma(cy, ky, wc[u].y));
            } else if (STAGE < 3) {
                st4[2 * u] = fma(2.0, kx, tmem_dbl(tw, 4 + 2 * u));
                st4[2 * u + 1] = fma(2.0, ky, tmem_dbl(tw, 5 + 2 * u));
                yn[u] = cmake(fma(cy, kx, tmem_dbl(tw, 2 * u)), fma(cy, ky, tmem_dbl(tw, 2 * u + 1)));
            } else {
                yn[u].x = fma(c.w6, tmem_dbl(tw, 4 + 2 * u) + kx, tmem_dbl(tw, 2 * u));
                yn[u].y = fma(c.w6, tmem_dbl(tw, 5 + 2 * u) + ky, tmem_dbl(tw, 2 * u + 1));
                st4[2 * u] = yn[u].x; st4[2 * u + 1] = yn[u].y;
            }
        }
        tmem_st8(c.trho + QME_TILE_TMW * r + (STAGE == 3 ? 0 : 8), st4);
#pragma unroll
        for (int u = 0; u < 2; ++u) *reinterpret_cast<cplx*>(pout + (r + 1) * ROWB + 512 * u) = yn[u];
        // boundary rows of a chunk also go into the neighbour CTA's halo row; the bytes are counted on its mbarrier
        if (r == 0 && c.up_dst) {
            st_async_c128<0>(c.up_dst + out_off, yn[0], c.up_bar + 8 * (STAGE & 1));
            st_async_c128<512>(c.up_dst + out_off, yn[1], c.up_bar + 8 * (STAGE & 1));
        }
        if (r == TR - 1 && c.dn_dst) {
            st_async_c128<0>(c.dn_dst + out_off, yn[0], c.dn_bar + 8 * (STAGE & 1));
            st_async_c128<512>(c.dn_dst + out_off, yn[1], c.dn_bar + 8 * (STAGE & 1));
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) { wp[u] = wc[u]; wc[u] = wn[u]; }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// TW ("tensor-memory window", opt-in: LIMEB200_TILE_V=8): the thread's OWN rows of the stage vector come from tensor
// memory, and the warp-uniform row coefficients from shared memory as three 16-byte broadcasts.  Tensor memory is the
// cheap port (tools/ubench/tmem_bw.cu: tcgen05.ld 439 B/clk per SM, tcgen05.st 619 B/clk, against 128 B/clk of shared
// memory).  Of the (TR + 2) x 2 window loads of a stage, TR x 2 re-read values the SAME thread stored in the previous
// stage; here they live in a third tensor-memory slot per row ([rho 8][acc 8][y 8]), updated in place (row r + 1 is
// fetched one row ahead, before row r is overwritten), and only the two halo rows of the window, the left / right /
// sandwich neighbours and the row coefficients still come from shared memory: 31 wavefronts per warp and patch row
// instead of 42.7.  The values still go to shared memory too -- the other threads read them there.
// Measured: 4.66e6 against 4.62e6 rho-steps/s, i.e. nothing -- the kernel is bound by the latency of a warp's dependent
// chain, not by the shared-memory pipe -- so this stays the opt-in form; identical bits.
template <int NP, int TR, int S, int STAGE>
__device__ __forceinline__ void qme_tile_stage_t(const QmeTileCtx<NP, TR, S>& c, char* smem, unsigned bufb) {
    constexpr int ROWB = NP * 16;
    const unsigned in_off = (STAGE & 1) ? bufb : 0u;
    const unsigned out_off = (STAGE & 1) ? 0u : bufb;
    const char* pown = smem + (c.own + in_off);
    const char* pl = smem + (c.nl + in_off);
    const char* pr = smem + (c.nr + in_off);
    const char* prc = smem + c.rowc;
    char* pout = smem + (c.own + out_off);
    const double cy = (STAGE == 2) ? c.dt : c.hdt;

    cplx wp[2], wc[2], wn[2];
    TmemRow8 ty;
    tmem_ld8_issue(c.trho + 16, ty);                       // own row 0 of the stage input
#pragma unroll
    for (int u = 0; u < 2; ++u) wp[u] = *reinterpret_cast<const cplx*>(pown + 512 * u);     // row above the patch
    tmem_ld_wait(ty);
#pragma unroll
    for (int u = 0; u < 2; ++u) wc[u] = cmake(tmem_dbl(ty, 2 * u), tmem_dbl(ty, 2 * u + 1));
#pragma unroll
    for (int r = 0; r < TR; ++r) {
        TmemRow16 tw;
        if (STAGE != 0) tmem_ld16_issue(c.trho + QME_TILE_TMW * r, tw);
        if (r + 1 < TR) {
            tmem_ld8_issue(c.trho + QME_TILE_TMW * (r + 1) + 16, ty);
        } else {
#pragma unroll
            for (int u = 0; u < 2; ++u) wn[u] = *reinterpret_cast<const cplx*>(pown + (TR + 1) * ROWB + 512 * u);   // row below
        }
        const cplx yl = *reinterpret_cast<const cplx*>(pl + (r + 1) * ROWB);
        const cplx yr = *reinterpret_cast<const cplx*>(pr + (r + 1) * ROWB);
        const double2 gd = *reinterpret_cast<const double2*>(prc + r * 48);
        const double2 gud = *reinterpret_cast<const double2*>(prc + r * 48 + 16);
        double2 xv2 = make_double2(0.0, 0.0);
        if (S > 0) xv2 = *reinterpret_cast<const double2*>(prc + r * 48 + 32);
        cplx ysrc[S > 0 ? S : 1][2];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int u = 0; u < 2; ++u)
                ysrc[s][u] = *reinterpret_cast<const cplx*>(smem + (c.xs[s][u] + in_off) + r * ROWB);
        cplx k[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const cplx a = (u == 0) ? yl : wc[0];
            const cplx b = (u == 0) ? wc[1] : yr;
            k[u].x = -c.cL[u] * a.y;
            k[u].y = c.cL[u] * a.x;
            k[u].x = fma(-c.cR[u], b.y, k[u].x);
            k[u].y = fma(c.cR[u], b.x, k[u].y);
        }
        if (STAGE != 0) {
            if (r + 1 < TR) tmem_ld_wait(tw, ty); else tmem_ld16_wait(tw);
        } else if (r + 1 < TR) {
            tmem_ld_wait(ty);
        }
        if (r + 1 < TR) {
#pragma unroll
            for (int u = 0; u < 2; ++u) wn[u] = cmake(tmem_dbl(ty, 2 * u), tmem_dbl(ty, 2 * u + 1));
        }
        const double gdx = gd.x, gdy = gd.y, gup = gud.x, gdn = gud.y;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const double dr = gdx + c.cdr[u], di = gdy + c.cdi[u];
            k[u].x = fma(dr, wc[u].x, k[u].x);
            k[u].x = fma(-di, wc[u].y, k[u].x);
            k[u].y = fma(dr, wc[u].y, k[u].y);
            k[u].y = fma(di, wc[u].x, k[u].y);
            k[u].x = fma(-gup, wp[u].y, k[u].x);
            k[u].y = fma(gup, wp[u].x, k[u].y);
            k[u].x = fma(-gdn, wn[u].y, k[u].x);
            k[u].y = fma(gdn, wn[u].x, k[u].y);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double xv = (s == 0) ? xv2.x : xv2.y;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double cf = xv * c.zv[s][u];
                k[u].x = fma(cf, ysrc[s][u].x, k[u].x);
                k[u].y = fma(cf, ysrc[s][u].y, k[u].y);
            }
        }
        cplx yn[2];
        double st4[4], sy4[4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const double kx = k[u].x, ky = k[u].y;
            if (STAGE == 0) {
                st4[2 * u] = kx; st4[2 * u + 1] = ky;
                yn[u] = cmake(fma(cy, kx, wc[u].x), fma(cy, ky, wc[u].y));
            } else if (STAGE < 3) {
                st4[2 * u] = fma(2.0, kx, tmem_dbl(tw, 4 + 2 * u));
                st4[2 * u + 1] = fma(2.0, ky, tmem_dbl(tw, 5 + 2 * u));
                yn[u] = cmake(fma(cy, kx, tmem_dbl(tw, 2 * u)), fma(cy, ky, tmem_dbl(tw, 2 * u + 1)));
            } else {
                yn[u].x = fma(c.w6, tmem_dbl(tw, 4 + 2 * u) + kx, tmem_dbl(tw, 2 * u));
                yn[u].y = fma(c.w6, tmem_dbl(tw, 5 + 2 * u) + ky, tmem_dbl(tw, 2 * u + 1));
                st4[2 * u] = yn[u].x; st4[2 * u + 1] = yn[u].y;
            }
            sy4[2 * u] = yn[u].x; sy4[2 * u + 1] = yn[u].y;
        }
        tmem_st8_nc(c.trho + QME_TILE_TMW * r + (STAGE == 3 ? 0 : 8), st4);
        tmem_st8_nc(c.trho + QME_TILE_TMW * r + 16, sy4);
#pragma unroll
        for (int u = 0; u < 2; ++u) *reinterpret_cast<cplx*>(pout + (r + 1) * ROWB + 512 * u) = yn[u];
        if (r == 0 && c.up_dst) {
            st_async_c128_nc<0>(c.up_dst + out_off, yn[0], c.up_bar + 8 * (STAGE & 1));
            st_async_c128_nc<512>(c.up_dst + out_off, yn[1], c.up_bar + 8 * (STAGE & 1));
        }
        if (r == TR - 1 && c.dn_dst) {
            st_async_c128_nc<0>(c.dn_dst + out_off, yn[0], c.dn_bar + 8 * (STAGE & 1));
            st_async_c128_nc<512>(c.dn_dst + out_off, yn[1], c.dn_bar + 8 * (STAGE & 1));
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) { wp[u] = wc[u]; wc[u] = wn[u]; }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// shared memory: [buffer 0][buffer 1][4 mbarriers][red E*32][part 2*C*E][tmem base][row coefficients R x 48 B]
static inline size_t qme_tile_smem(int NP, int P, int chunk, int C, int E) {
    const size_t nbr = (size_t)P * (chunk + 2);
    const size_t e = E > 0 ? E : 1;
    return 2 * nbr * NP * 16 + 64 + e * 32 * 16 + 2 * (size_t)C * e * 16 + 16 + (size_t)P * chunk * 48;
}

template <int NP, int TR, int S, bool TW>
__global__ void __launch_bounds__(512, 1)
qme_tile_kernel(QmeTileArgs a) {
    extern __shared__ __align__(16) char smem_raw[];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int ROWB = NP * 16;
    constexpr int CB = NP / 64;
    constexpr int SS = S > 0 ? S : 1;
    const int C = a.C, P = a.P, chunk = a.chunk;
    const int rank = (C > 1) ? (int)cluster.block_rank() : 0;
    const int b = blockIdx.x / C;
    const size_t vb = (a.nb > 1) ? (size_t)b : 0;
    const int NBR = P * (chunk + 2);                  // buffer rows
    const int R = P * chunk;                          // own rows
    const int T = blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cb = warp % CB, g = warp / CB;          // column block, row group (patch rows g*TR .. g*TR+TR-1 of the own rows)
    const int Ee = max(a.E, 1);

    const unsigned bufb = (unsigned)NBR * ROWB;
    const unsigned o_bar = 2 * bufb;                  // 4 mbarriers: halo[2], partial sums[2]
    const unsigned o_red = o_bar + 64;                // [E][32] cplx
    const unsigned o_part = o_red + Ee * 32 * 16;     // [2][C][E] cplx (rank 0)
    const unsigned o_tm = o_part + 2 * C * Ee * 16;
    const unsigned o_rowc = o_tm + 16;                // [R][6] doubles: row coefficients (TW)
    cplx* buf0 = reinterpret_cast<cplx*>(smem_raw);
    cplx* red = reinterpret_cast<cplx*>(smem_raw + o_red);
    cplx* part = reinterpret_cast<cplx*>(smem_raw + o_part);
    unsigned* tmbase = reinterpret_cast<unsigned*>(smem_raw + o_tm);

    // ---- tensor memory: QME_TILE_TMW * TR words per thread; warps sharing a lane quadrant take consecutive column
    // ranges.  All 512 columns are taken: the launch requests > half of the shared memory, so one CTA per SM.
    constexpr int TM_PER_WARP = QME_TILE_TMW * TR;
    constexpr int TM_COLS = 512;
    static_assert(TM_PER_WARP * 4 <= TM_COLS, "tensor memory budget: 16 warps");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmbase)), "n"(TM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }

    // ---- load rho (own + halo rows, chain order / lane-interleaved columns) into buffer 0, zero buffer 1
    const cplx* grho = a.rho + (size_t)b * a.N * a.N;
    const int* brow = a.brow + rank * NBR;
    for (int l = threadIdx.x; l < NBR * NP; l += T) {
        const int br = l / NP, pos = l - br * NP;
        const int orow = brow[br], ocol = a.colold[pos];
        cplx v = cmake(0, 0);
        if (orow >= 0 && ocol >= 0) v = grho[(size_t)orow * a.N + ocol];
        buf0[l] = v;
        buf0[NBR * NP + l] = cmake(0, 0);
    }

    if (TW) {
        const double* rc = a.rowc + ((size_t)vb * C + rank) * R * QME_TILE_ROWC;
        double* rs = reinterpret_cast<double*>(smem_raw + o_rowc);
        for (int l = threadIdx.x; l < R * QME_TILE_ROWC; l += T) rs[l] = rc[l];
    }

    QmeTileCtx<NP, TR, S> c;
    const int own0 = g * TR;                           // first own row of the patch
    const int p = own0 / chunk, t0 = 1 + own0 % chunk; // path, buffer row inside the chunk block
    const int wb = p * (chunk + 2) + t0 - 1;           // window row 0
    const int pos0 = 64 * cb + lane;                   // column position of u = 0 (u = 1: + 32)
    const unsigned sbase = smem_u32(smem_raw);
    c.own = wb * ROWB + pos0 * 16;
    c.nl = wb * ROWB + a.colpos[pos0 * 4 + 0];
    c.nr = wb * ROWB + a.colpos[(pos0 + 32) * 4 + 1];
#pragma unroll
    for (int s = 0; s < SS; ++s) {
        const int xr = (S > 0) ? a.xs0[(rank * (R / TR) + g) * QME_TILE_MAXS + s] : 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            c.xs[s][u] = (S > 0) ? (unsigned)(xr * ROWB + a.colpos[(pos0 + 32 * u) * 4 + 2 + s]) : 0u;
            c.zv[s][u] = (S > 0) ? a.colc[(vb * NP + pos0 + 32 * u) * QME_TILE_COLC + 4 + s] : 0.0;
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const double* cc = a.colc + (vb * NP + pos0 + 32 * u) * QME_TILE_COLC;
        c.cdr[u] = cc[0]; c.cdi[u] = cc[1]; c.cL[u] = cc[2]; c.cR[u] = cc[3];
    }
    c.hdt = 0.5 * a.dt; c.dt = a.dt; c.w6 = a.dt / 6.0;
    c.rowc = o_rowc + (unsigned)own0 * (QME_TILE_ROWC * 8);

    // ---- neighbours: mapped shared-memory windows and mbarriers
    const unsigned bar0 = sbase + o_bar;
    const bool first_tile = (t0 == 1), last_tile = (t0 + TR - 1 == chunk);
    c.up_dst = c.dn_dst = c.up_bar = c.dn_bar = 0;
    unsigned r0_part = 0, r0_bar = 0, halo_bytes = 0;
    if (C > 1) {
        if (threadIdx.x == 0) {
            for (int m = 0; m < 4; ++m) mbar_init(bar0 + 8 * m, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (rank > 0) {
            halo_bytes += P * ROWB;
            if (first_tile) {       // own row t = 1 is the neighbour's halo row t = chunk + 1 of the same path
                c.up_dst = mapa_u32(sbase + (p * (chunk + 2) + chunk + 1) * ROWB + pos0 * 16, rank - 1);
                c.up_bar = mapa_u32(bar0, rank - 1);
            }
        }
        if (rank < C - 1) {
            halo_bytes += P * ROWB;
            if (last_tile) {        // own row t = chunk is the neighbour's halo row t = 0
                c.dn_dst = mapa_u32(sbase + (p * (chunk + 2)) * ROWB + pos0 * 16, rank + 1);
                c.dn_bar = mapa_u32(bar0, rank + 1);
            }
        }
        r0_part = mapa_u32(sbase + o_part, 0);
        r0_bar = mapa_u32(bar0 + 16, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    c.trho = *tmbase + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(warp >> 2) * TM_PER_WARP;

#pragma unroll
    for (int r = 0; r < TR; ++r) {
        double rh[4], z4[4] = {0.0, 0.0, 0.0, 0.0}, c4[4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const cplx v = *reinterpret_cast<const cplx*>(smem_raw + c.own + (r + 1) * ROWB + 512 * u);
            rh[2 * u] = v.x; rh[2 * u + 1] = v.y;
        }
        const double* rcp = a.rowc + (((size_t)vb * C + rank) * R + own0 + r) * QME_TILE_ROWC;
        tmem_st8(c.trho + QME_TILE_TMW * r, rh);
        tmem_st8(c.trho + QME_TILE_TMW * r + 8, z4);
        c4[0] = rcp[0]; c4[1] = rcp[1]; c4[2] = rcp[2]; c4[3] = rcp[3];
        if (TW) tmem_st8(c.trho + QME_TILE_TMW * r + 16, rh);      // third slot: the stage input (= rho before stage 0)
        else tmem_st8(c.trho + QME_TILE_TMW * r + 16, c4);
        c4[0] = rcp[4]; c4[1] = rcp[5]; c4[2] = 0.0; c4[3] = 0.0;
        tmem_st8(c.trho + QME_TILE_TMW * r + 24, c4);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");

    if (C > 1) cluster.sync();

    for (int step = 0; step < a.nsteps; ++step) {
#define QME_TILE_STAGE(ST)                                                                           \
        if (C > 1 && threadIdx.x == 0) mbar_arrive_expect_tx(bar0 + 8 * ((ST) & 1), halo_bytes);      \
        if (TW) qme_tile_stage_t<NP, TR, S, ST>(c, smem_raw, bufb);                                   \
        else qme_tile_stage<NP, TR, S, ST>(c, smem_raw, bufb);                                        \
        __syncthreads();                                                                              \
        if (C > 1) mbar_wait(bar0 + 8 * ((ST) & 1), (unsigned)(((ST) >> 1) & 1));
        QME_TILE_STAGE(0)
        if (C > 1 && a.obs && step > 0 && rank == 0) {
            // partial sums of the previous step (pushed by the other CTAs with st.async) are complete
            const int sp = step - 1;
            for (int e = threadIdx.x; e < a.E; e += T) {
                mbar_wait(bar0 + 16 + 8 * (sp & 1), (unsigned)((sp >> 1) & 1));
                const cplx* pp = part + (size_t)(sp & 1) * C * a.E;
                cplx sum = cmake(0, 0);
                for (int q = 0; q < C; ++q) sum = cadd(sum, pp[q * a.E + e]);
                a.obs[((size_t)sp * a.B + b) * a.E + e] = sum;
            }
        }
        QME_TILE_STAGE(1)
        QME_TILE_STAGE(2)
        QME_TILE_STAGE(3)
#undef QME_TILE_STAGE
        // buffer 0 now holds rho_{n+1} (own + halo rows)
        if (a.obs) {
            for (int e = 0; e < a.E; ++e) {
                cplx v = cmake(0, 0);
                for (int n = a.eptr[e] + threadIdx.x; n < a.eptr[e + 1]; n += T)
                    if (a.erank[n] == rank) cfma(v, a.eval[n], buf0[a.eoff[n]]);
                for (int off = 16; off > 0; off >>= 1) {
                    v.x += __shfl_down_sync(0xffffffffu, v.x, off);
                    v.y += __shfl_down_sync(0xffffffffu, v.y, off);
                }
                if (lane == 0) red[e * 32 + warp] = v;
            }
            __syncthreads();
            if (C > 1 && rank == 0 && threadIdx.x == 0)
                mbar_arrive_expect_tx(bar0 + 16 + 8 * (step & 1), (unsigned)((C - 1) * a.E * 16));
            for (int e = threadIdx.x; e < a.E; e += T) {
                cplx sum = cmake(0, 0);
                for (int w = 0; w < (T + 31) / 32; ++w) sum = cadd(sum, red[e * 32 + w]);
                if (C == 1) a.obs[((size_t)step * a.B + b) * a.E + e] = sum;
                else if (rank == 0) part[(size_t)(step & 1) * C * a.E + e] = sum;
                else st_async_c128<0>(r0_part + (unsigned)(((step & 1) * C + rank) * a.E + e) * 16u, sum,
                                      r0_bar + 8 * (step & 1));
            }
            __syncthreads();          // red[] is reused by the next step
        }
        if (a.traj && ((step + 1) % a.traj_every) == 0) {
            cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * a.N * a.N;
#pragma unroll
            for (int r = 0; r < TR; ++r) {
                const int orow = brow[wb + 1 + r];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int ocol = a.colold[pos0 + 32 * u];
                    if (orow >= 0 && ocol >= 0) dst[(size_t)orow * a.N + ocol] = *reinterpret_cast<const cplx*>(smem_raw + c.own + (r + 1) * ROWB + 512 * u);
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
            for (int q = 0; q < C; ++q) sum = cadd(sum, pp[q * a.E + e]);
            a.obs[((size_t)sp * a.B + b) * a.E + e] = sum;
        }
    }
    cplx* out = a.rho + (size_t)b * a.N * a.N;
#pragma unroll
    for (int r = 0; r < TR; ++r) {
        const int orow = brow[wb + 1 + r];
        TmemRow tw;
        tmem_ld32_issue(c.trho + QME_TILE_TMW * r, tw);
        tmem_ld32_wait(tw);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int ocol = a.colold[pos0 + 32 * u];
            if (orow >= 0 && ocol >= 0) out[(size_t)orow * a.N + ocol] = cmake(tmem_dbl(tw, 2 * u), tmem_dbl(tw, 2 * u + 1));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmbase), "n"(TM_COLS));
    if (C > 1) cluster.sync();      // keep shared memory alive until the neighbours' remote stores are done
}

template <int NP, int TR, int S, bool TW>
static int qme_tile_launch_one(const QmeTileArgs& a, size_t smem, cudaStream_t st) {
    auto kern = qme_tile_kernel<NP, TR, S, TW>;
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int W = (a.P * a.chunk / TR) * (NP / 64);
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

// implemented in qme_tile_inst.cu
int qme_tile_launch(const QmeTileArgs& a, int NP, int S, size_t smem, cudaStream_t st);
