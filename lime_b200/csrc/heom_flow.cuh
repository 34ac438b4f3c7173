// Dataflow-synchronised persistent HEOM propagator (diagonal coupling operators, one hierarchy).
//
// The barrier-per-stage kernels (heom_persist_*) spend most of a stage waiting: a grid barrier costs ~3 us on one
// GPU, the cross-GPU barrier ~13 us (peer-store acknowledgement + remote arrival count + poll = 2.5 NVLink
// traversals on the critical path of EVERY stage), while the arithmetic of a stage of the 3060-ADO FMO hierarchy is
// ~1 us.  Here there is no barrier at all:
//
//   * the stage vectors are TAGGED: element e of stage vector y_s lives in entry e of buffer s & 1 as two 16-byte
//     words {bits(re), tag}, {bits(im), tag} with tag = tag0 + s, a 64-bit stage counter that is monotonic over the
//     launches of a plan.  A 16-byte word is written with ONE store and read with ONE load, so whoever sees the
//     tag sees the value (the LL idea of collective libraries, applied per matrix element);
//   * a thread that needs a neighbour value polls that entry (volatile loads, all neighbours of an element in
//     flight together) until both tags say "stage s" -- it waits for exactly the <= 8 values it reads and for nothing
//     else; producers never wait for consumers;
//   * on several GPUs the owner of an ADO stores the new entry into its own buffer and, through peer pointers over
//     NVLink, into the buffers of the ranks that read it (halo masks); the consumer polls its OWN memory, so a
//     stage costs one one-way NVLink traversal instead of a fence + a flag + a barrier;
//   * write-after-read safety of the two-buffer ping-pong needs no extra synchronisation: the coupling graph is
//     symmetric at element level (n reads n -+ e_k  <=>  n -+ e_k reads n; zero-coefficient directions are still
//     waited for to keep it symmetric), so by the time (n, ij) overwrites y_{s-1}(n, ij) it has seen y_s of every
//     reader of that entry, and a reader publishes y_s only after its reads of y_{s-1} have completed.
//
// Mapping: every CTA owns a contiguous block of ADOs and walks it in tiles of `apc` ADOs (one matrix element per
// thread); -i[H, .] goes through a shared-memory copy of the tile; rho and the RK4 accumulator stay in global memory
// (L2), touched once per stage with coalesced accesses.  All CTAs are co-resident (cooperative launch), which the
// spinning relies on.  Waits are bounded (~5 s): a missing producer sets err[0] instead of hanging the GPU.
#pragma once

struct HeomFlowArgs {
    HeomDev d;
    long long row_lo, row_hi;        // owned ADO range of this rank
    int apc;                         // ADOs per tile
    int maxn;                        // largest occupation number n_k of the hierarchy
    int nsteps, E, traj_every;
    double dt;
    ulonglong2* T[2];                // local tagged stage vectors: entry e -> words 2e, 2e + 1
    ulonglong2* Tp[2][7];            // the peers' buffers (slot q = q-th rank other than this one)
    int npeer;
    const unsigned char* peer_mask;  // [nhe] bit q: peer slot q reads this ADO (null: every peer)
    unsigned long long tag0;         // tag of the state held in T[0] on entry
    cplx* rho;                       // [nhe][nn] local: owned rows are read on entry and hold rho_final on exit
    cplx* acc;                       // [nhe][nn] RK4 accumulator scratch (owned rows)
    const cplx* eT; cplx* obs; cplx* traj;      // tier-0 outputs; written by the CTA that owns ADO 0 (or null)
    unsigned* err;
};

__device__ __forceinline__ ulonglong2 flow_ld(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void flow_st(ulonglong2* p, double v, unsigned long long tag) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((unsigned long long)__double_as_longlong(v)), "l"(tag) : "memory");
}

#define HEOM_FLOW_NE 8
template <int NN_, int MAXT>          // NN_ = n when known at compile time (2, 3, 7), 0 = runtime
__global__ void __launch_bounds__(MAXT, 1)
heom_flow_kernel(HeomFlowArgs a) {
    extern __shared__ double2 smem[];
    const HeomDev& d = a.d;
    const int n = NN_ ? NN_ : d.n, nn = n * n, T = blockDim.x;
    cplx* Hs = smem;                         // [nn]
    cplx* ys = Hs + nn;                      // [apc * nn]
    cplx* r0s = ys + (size_t)a.apc * nn;     // [nn] tier-0 rho of the finished step (observables)
    for (int l = threadIdx.x; l < nn; l += T) Hs[l] = d.H[l];
    const long long nown = a.row_hi - a.row_lo;
    // contiguous, balanced blocks of ADOs per CTA
    const long long per = (nown + gridDim.x - 1) / gridDim.x;
    const long long cta_lo = a.row_lo + per * blockIdx.x;
    const long long cta_hi = min(a.row_hi, cta_lo + per);
    const int g = threadIdx.x / nn, idx = threadIdx.x - g * nn;
    const int i = idx / n, j = idx - i * n;
    const bool lane_ok = g < a.apc;
    const bool obs_cta = (a.obs || a.traj) && cta_lo == 0 && cta_hi > 0;
    const int t_lo = lane_ok ? __ldg(d.em_start + idx) : 0;
    const int t_hi = lane_ok ? __ldg(d.em_start + idx + 1) : 0;
    const long long limit = 10000000000LL;
    __syncthreads();

    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const unsigned long long want = a.tag0 + 4ull * step + stage;
            const ulonglong2* Tin = a.T[stage & 1];
            ulonglong2* Tout = a.T[(stage + 1) & 1];
            const bool last = (step == a.nsteps - 1 && stage == 3);
#pragma unroll 1
            for (long long tile = cta_lo; tile < cta_hi; tile += a.apc) {
                const long long ado = tile + g;
                const bool act = lane_ok && ado < cta_hi;
                const size_t own = (size_t)(act ? ado : cta_lo) * nn + idx;
                // ---- neighbour entries of this element: fixed slots (mode t, down / up); a missing neighbour points at
                // the element's own entry (always tagged `want`) and later gets coefficient 0
                const unsigned own32 = (unsigned)own;
                unsigned off[HEOM_FLOW_NE];
#pragma unroll
                for (int s = 0; s < HEOM_FLOW_NE; ++s) off[s] = own32;
                if (act) {
                    const int* dn = d.dn + ado * d.nmodes;
                    const int* up = d.up + ado * d.nmodes;
#pragma unroll
                    for (int q = 0; q < HEOM_FLOW_NE / 2; ++q) {
                        const int t = t_lo + q;
                        if (t < t_hi) {
                            const int m = __ldg(d.em_mode + t);
                            const int id = __ldg(dn + m), iu = __ldg(up + m);
                            if (id >= 0) off[2 * q] = (unsigned)id * (unsigned)nn + (unsigned)idx;
                            if (iu >= 0) off[2 * q + 1] = (unsigned)iu * (unsigned)nn + (unsigned)idx;
                        }
                    }
                }
                // ---- wait for the own value and the neighbours: all loads in flight together, re-poll until tagged
                double vx[HEOM_FLOW_NE + 1], vy[HEOM_FLOW_NE + 1];
                {
                    const long long t0 = clock64();
                    bool ready;
                    unsigned spins = 0;
                    do {
                        ready = true;
#pragma unroll
                        for (int s = 0; s <= HEOM_FLOW_NE; ++s) {
                            const ulonglong2* e = Tin + 2 * (size_t)(s < HEOM_FLOW_NE ? off[s] : own32);
                            const ulonglong2 w0 = flow_ld(e), w1 = flow_ld(e + 1);
                            vx[s] = __longlong_as_double((long long)w0.x);
                            vy[s] = __longlong_as_double((long long)w1.x);
                            ready = ready && w0.y == want && w1.y == want;
                        }
                        if (!ready && ((++spins & 1023u) == 0) &&
                            (clock64() - t0 > limit || *(volatile unsigned*)a.err)) {
                            atomicExch(a.err, 1u);
                            ready = true;
                        }
                    } while (!ready);
                }
                const cplx yv = cmake(vx[HEOM_FLOW_NE], vy[HEOM_FLOW_NE]);
                if (lane_ok) ys[threadIdx.x] = yv;
                __syncthreads();
                if (act) {
                    cplx k = NN_ ? heom_sys_t<(NN_ ? NN_ : 2)>(Hs, ys + (size_t)g * nn, i, j)
                                 : heom_sys(Hs, n, ys + (size_t)g * nn, i, j);
                    const double damp = heom_damp(d, 0, ado);
                    k.x = fma(-damp, yv.x, k.x);
                    k.y = fma(-damp, yv.y, k.y);
                    const int* st = d.states + ado * d.nmodes;
#pragma unroll
                    for (int q = 0; q < HEOM_FLOW_NE / 2; ++q) {
                        const int t = t_lo + q;
                        if (t < t_hi) {
                            const int m = __ldg(d.em_mode + t);
                            const double2 v = __ldg(d.em_v + t);
                            if (off[2 * q] != own32)
                                cfma(k, heom_dn_coef(d, 0, m, (double)__ldg(st + m), v), cmake(vx[2 * q], vy[2 * q]));
                            if (off[2 * q + 1] != own32)
                                cfma(k, cscale(v.x - v.y, d.pref_up), cmake(vx[2 * q + 1], vy[2 * q + 1]));
                        }
                    }
                    cplx r = a.rho[own];
                    cplx ac = (stage == 0) ? cmake(0, 0) : a.acc[own];
                    const cplx yn = heom_rk_update(stage, k, r, ac, a.dt);
                    if (stage < 3) a.acc[own] = ac; else a.rho[own] = r;
                    flow_st(Tout + 2 * own, yn.x, want + 1);
                    flow_st(Tout + 2 * own + 1, yn.y, want + 1);
                    if (a.npeer) {
                        const unsigned m = (last || !a.peer_mask) ? 0xffu : a.peer_mask[ado];
                        for (int q = 0; q < a.npeer; ++q)
                            if ((m >> q) & 1u) {
                                ulonglong2* tp = a.Tp[(stage + 1) & 1][q];
                                flow_st(tp + 2 * own, yn.x, want + 1);
                                flow_st(tp + 2 * own + 1, yn.y, want + 1);
                            }
                    }
                    if (stage == 3 && obs_cta && ado == 0) r0s[idx] = r;
                }
                __syncthreads();          // ys is overwritten by the next tile
            }
        }
        // tier-0 observables / trajectory of the finished step: the CTA that owns ADO 0
        if (obs_cta) {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = T >> 5;
            if (a.obs)
                for (int e = warp; e < a.E; e += nw) {
                    cplx v = cmake(0, 0);
                    for (int l = lane; l < nn; l += 32) cfma(v, __ldg(a.eT + (size_t)e * nn + l), r0s[l]);
                    for (int o = 16; o > 0; o >>= 1) {
                        v.x += __shfl_down_sync(0xffffffffu, v.x, o);
                        v.y += __shfl_down_sync(0xffffffffu, v.y, o);
                    }
                    if (lane == 0) a.obs[(size_t)step * a.E + e] = v;
                }
            if (a.traj && ((step + 1) % a.traj_every) == 0)
                for (int l = threadIdx.x; l < nn; l += T) a.traj[(size_t)(step / a.traj_every) * nn + l] = r0s[l];
            __syncthreads();
        }
    }
}

// T0 <- tagged copy of a full state vector (run on every rank before the propagation; stream-ordered)
__global__ void __launch_bounds__(256)
heom_flow_pack_kernel(const cplx* __restrict__ y, long long count, unsigned long long tag, ulonglong2* __restrict__ T0) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
        const cplx v = y[e];
        T0[2 * e] = make_ulonglong2((unsigned long long)__double_as_longlong(v.x), tag);
        T0[2 * e + 1] = make_ulonglong2((unsigned long long)__double_as_longlong(v.y), tag);
    }
}
// full state vector <- values of a tagged buffer; entries whose tag differs from `tag` raise err[1]
__global__ void __launch_bounds__(256)
heom_flow_unpack_kernel(const ulonglong2* __restrict__ T0, long long count, unsigned long long tag, cplx* __restrict__ y,
                        unsigned* err) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
        const ulonglong2 a = T0[2 * e], b = T0[2 * e + 1];
        if (a.y != tag || b.y != tag) atomicExch(err + 1, 1u);
        y[e] = cmake(__longlong_as_double((long long)a.x), __longlong_as_double((long long)b.x));
    }
}

// ---- register / shared-memory resident variant: EPT matrix elements per thread for the whole run ---------------
// The latency-bound case (up to ~2000 elements per SM: the 3060-ADO FMO hierarchy on 1-8 GPUs, the 38 760-ADO one on
// 8).  rho, the RK4 accumulator and the current stage value of an element live in registers; its <= 8 neighbour
// entries are cached in shared memory as ONE packed word each (entry index | n_k << 26; the finished coefficient is
// looked up in a per-(matrix element, slot, n_k) table of a few tens of KB), computed once.  A stage is: publish the own
// ADOs through shared memory for -i[H, .], then per element poll the 8 tagged neighbour entries (all in flight, only
// untagged ones re-read), and publish the new tagged entry locally and into the reading peers.  No barrier, no round
// trip through global memory for the state.
#define HEOM_FLOW_IDXBITS 26
template <int NN_, int MAXT, int MINB, int EPT>
__global__ void __launch_bounds__(MAXT, MINB)
heom_flow_cached_kernel(HeomFlowArgs a) {
    extern __shared__ double2 smem[];
    const HeomDev& d = a.d;
    const int n = NN_ ? NN_ : d.n, nn = n * n, T = blockDim.x;
    cplx* Hs = smem;                                        // [nn]
    cplx* ys = Hs + nn;                                     // [apc * nn]
    cplx* r0s = ys + (size_t)a.apc * nn;                    // [nn]
    const int NK = a.maxn + 1;
    cplx* ctab = r0s + nn;                                  // [nn][NE][NK] finished coefficient for n_k = 0 .. maxn
    // EPT == 1: there is room to cache the finished coefficients too ([NE][T]), which saves the n_k conversion and
    // two multiplications per neighbour and stage
    constexpr bool FULLCF = (EPT == 1);
    cplx* ecf = ctab + (size_t)nn * HEOM_FLOW_NE * NK;      // [NE][T] (FULLCF only)
    unsigned* eoff = reinterpret_cast<unsigned*>(ecf + (FULLCF ? (size_t)HEOM_FLOW_NE * T : 0));     // [EPT][NE][T]
    for (int l = threadIdx.x; l < nn; l += T) Hs[l] = d.H[l];
    for (int l = threadIdx.x; l < nn * HEOM_FLOW_NE * NK; l += T) ctab[l] = cmake(0, 0);
    __syncthreads();
    for (int l = threadIdx.x; l < nn; l += T) {
        int q = 0;
        for (int t = d.em_start[l]; t < d.em_start[l + 1] && q < HEOM_FLOW_NE / 2; ++t, ++q) {
            const int m = d.em_mode[t];
            const double2 v = d.em_v[t];
            for (int nk = 1; nk < NK; ++nk)     // down-coupling: n_k (q_i c_k - q_j conj c_k), exactly as the other kernels form it
                ctab[(l * HEOM_FLOW_NE + 2 * q) * NK + nk] = heom_dn_coef(d, 0, m, (double)nk, v);
            if (NK > 1) ctab[(l * HEOM_FLOW_NE + 2 * q + 1) * NK + 1] = cscale(v.x - v.y, d.pref_up);     // up-coupling: slot "n_k" = 1
        }
    }
    const long long nown = a.row_hi - a.row_lo;
    const bool obs_cta = (a.obs || a.traj) && a.row_lo == 0 && blockIdx.x == 0;
    bool act[EPT];
    unsigned own[EPT], pmask[EPT];
    int idx[EPT], gg[EPT];
    cplx rreg[EPT], areg[EPT], ycur[EPT];
    double damp[EPT];
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        const int l = threadIdx.x + u * T;
        gg[u] = l / nn;
        idx[u] = l - gg[u] * nn;
        const long long item = (long long)blockIdx.x * a.apc + gg[u];
        act[u] = gg[u] < a.apc && item < nown;
        const long long ado = a.row_lo + (act[u] ? item : 0);
        own[u] = (unsigned)(ado * nn + idx[u]);
        rreg[u] = act[u] ? a.rho[own[u]] : cmake(0, 0);
        areg[u] = cmake(0, 0);
        ycur[u] = rreg[u];
        damp[u] = act[u] ? heom_damp(d, 0, ado) : 0.0;
        pmask[u] = (a.peer_mask && act[u]) ? a.peer_mask[ado] : 0xffu;
#pragma unroll
        for (int s = 0; s < HEOM_FLOW_NE; ++s) eoff[(u * HEOM_FLOW_NE + s) * T + threadIdx.x] = own[u];
        if (act[u]) {
            const int* st = d.states + ado * d.nmodes;
            const int* dn = d.dn + ado * d.nmodes;
            const int* up = d.up + ado * d.nmodes;
            int q = 0;
            for (int t = d.em_start[idx[u]]; t < d.em_start[idx[u] + 1] && q < HEOM_FLOW_NE / 2; ++t, ++q) {
                const int m = d.em_mode[t];
                const int id = dn[m], iu = up[m];
                if (id >= 0)
                    eoff[(u * HEOM_FLOW_NE + 2 * q) * T + threadIdx.x] =
                        ((unsigned)id * (unsigned)nn + (unsigned)idx[u]) | ((unsigned)st[m] << HEOM_FLOW_IDXBITS);
                if (iu >= 0)        // waited for even when the coefficient vanishes: keeps the coupling graph symmetric
                    eoff[(u * HEOM_FLOW_NE + 2 * q + 1) * T + threadIdx.x] =
                        ((unsigned)iu * (unsigned)nn + (unsigned)idx[u]) | (1u << HEOM_FLOW_IDXBITS);
            }
        }
    }
    const long long limit = 10000000000LL;
    __syncthreads();
    if (FULLCF) {
#pragma unroll
        for (int s = 0; s < HEOM_FLOW_NE; ++s) {
            const unsigned w = eoff[s * T + threadIdx.x];
            ecf[s * T + threadIdx.x] = ctab[(idx[0] * HEOM_FLOW_NE + s) * NK + (w >> HEOM_FLOW_IDXBITS)];
        }
    }
    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const unsigned long long want = a.tag0 + 4ull * step + stage;
            const ulonglong2* Tin = a.T[stage & 1];
            ulonglong2* Tout = a.T[(stage + 1) & 1];
            const bool last = (step == a.nsteps - 1 && stage == 3);
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (gg[u] < a.apc) ys[threadIdx.x + u * T] = ycur[u];
            __syncthreads();
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                if (!act[u]) continue;
                // first poll round: all 16 loads are ISSUED, then -i[H, Y_a] and the damping (which need only this CTA's own
                // ADOs, through shared memory) are computed while they are in flight, then the tags are checked
                unsigned word[HEOM_FLOW_NE];
#pragma unroll
                for (int s = 0; s < HEOM_FLOW_NE; ++s) word[s] = eoff[(u * HEOM_FLOW_NE + s) * T + threadIdx.x];
                ulonglong2 w0[HEOM_FLOW_NE], w1[HEOM_FLOW_NE];
#pragma unroll
                for (int s = 0; s < HEOM_FLOW_NE; ++s) {
                    const ulonglong2* e = Tin + 2 * (size_t)(word[s] & ((1u << HEOM_FLOW_IDXBITS) - 1u));
                    w0[s] = flow_ld(e);
                    w1[s] = flow_ld(e + 1);
                }
                const int i = idx[u] / n, j = idx[u] - i * n;
                cplx k = NN_ ? heom_sys_t<(NN_ ? NN_ : 2)>(Hs, ys + (size_t)gg[u] * nn, i, j)
                             : heom_sys(Hs, n, ys + (size_t)gg[u] * nn, i, j);
                k.x = fma(-damp[u], ycur[u].x, k.x);
                k.y = fma(-damp[u], ycur[u].y, k.y);
                double vx[HEOM_FLOW_NE], vy[HEOM_FLOW_NE];
                bool ready = true;
#pragma unroll
                for (int s = 0; s < HEOM_FLOW_NE; ++s) {
                    vx[s] = __longlong_as_double((long long)w0[s].x);
                    vy[s] = __longlong_as_double((long long)w1[s].x);
                    ready = ready && w0[s].y == want && w1[s].y == want;
                }
                if (!ready) {
                    // further rounds: all loads of a round in flight together, repeated until every entry is tagged
                    // (re-reading only the untagged entries was measured slower: the per-entry branches serialise the loads)
                    const long long t0 = clock64();
                    unsigned spins = 0;
                    do {
                        ready = true;
#pragma unroll
                        for (int s = 0; s < HEOM_FLOW_NE; ++s) {
                            const ulonglong2* e = Tin + 2 * (size_t)(word[s] & ((1u << HEOM_FLOW_IDXBITS) - 1u));
                            const ulonglong2 a0 = flow_ld(e), a1 = flow_ld(e + 1);
                            vx[s] = __longlong_as_double((long long)a0.x);
                            vy[s] = __longlong_as_double((long long)a1.x);
                            ready = ready && a0.y == want && a1.y == want;
                        }
                        if (!ready && ((++spins & 1023u) == 0) && (clock64() - t0 > limit || *(volatile unsigned*)a.err)) {
                            atomicExch(a.err, 1u);
                            ready = true;
                        }
                    } while (!ready);
                }
#pragma unroll
                for (int s = 0; s < HEOM_FLOW_NE; ++s) {
                    if (FULLCF) {
                        cfma(k, ecf[s * T + threadIdx.x], cmake(vx[s], vy[s]));
                    } else {
                        cfma(k, ctab[(idx[u] * HEOM_FLOW_NE + s) * NK + (word[s] >> HEOM_FLOW_IDXBITS)], cmake(vx[s], vy[s]));
                    }
                }
                const cplx yn = heom_rk_update(stage, k, rreg[u], areg[u], a.dt);
                flow_st(Tout + 2 * (size_t)own[u], yn.x, want + 1);
                flow_st(Tout + 2 * (size_t)own[u] + 1, yn.y, want + 1);
                if (a.npeer) {
                    const unsigned m = last ? 0xffu : pmask[u];
                    for (int q = 0; q < a.npeer; ++q)
                        if ((m >> q) & 1u) {
                            ulonglong2* tp = a.Tp[(stage + 1) & 1][q];
                            flow_st(tp + 2 * (size_t)own[u], yn.x, want + 1);
                            flow_st(tp + 2 * (size_t)own[u] + 1, yn.y, want + 1);
                        }
                }
                ycur[u] = yn;
                if (stage == 3 && obs_cta && own[u] < (unsigned)nn) r0s[idx[u]] = rreg[u];
            }
            __syncthreads();
        }
        if (obs_cta) {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = T >> 5;
            if (a.obs)
                for (int e = warp; e < a.E; e += nw) {
                    cplx v = cmake(0, 0);
                    for (int l = lane; l < nn; l += 32) cfma(v, __ldg(a.eT + (size_t)e * nn + l), r0s[l]);
                    for (int o = 16; o > 0; o >>= 1) {
                        v.x += __shfl_down_sync(0xffffffffu, v.x, o);
                        v.y += __shfl_down_sync(0xffffffffu, v.y, o);
                    }
                    if (lane == 0) a.obs[(size_t)step * a.E + e] = v;
                }
            if (a.traj && ((step + 1) % a.traj_every) == 0)
                for (int l = threadIdx.x; l < nn; l += T) a.traj[(size_t)(step / a.traj_every) * nn + l] = r0s[l];
            __syncthreads();
        }
    }
#pragma unroll
    for (int u = 0; u < EPT; ++u)
        if (act[u]) a.rho[own[u]] = rreg[u];
}
