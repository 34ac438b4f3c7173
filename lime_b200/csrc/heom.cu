// HEOM: index tables (host), multi-index hierarchy RK4 (device), _heom_dl-exact sweep.
//
// Rules, lime/heom/heom.py:156-216 (+ system term lime/oqs.py:1854):
//   d rho_n/dt = -i[H,rho_n] - (sum_k n_k nu_k) rho_n
//              + sum_k pref_dn n_k (c_k Q_k rho_{n-e_k} - conj(c_k) rho_{n-e_k} Q_k)
//              + sum_k pref_up [Q_k, rho_{n+e_k}]
// grouped per distinct coupling operator q (modes k with qmap[k] == q):
//   L_q = sum_k (pref_dn n_k c_k       rho_{n-e_k} + pref_up rho_{n+e_k})
//   R_q = sum_k (pref_dn n_k conj(c_k) rho_{n-e_k} + pref_up rho_{n+e_k})
//   d rho_n/dt += Q_q L_q - R_q Q_q
// so that an ADO costs 2 + 2*nq small matrix products instead of 2 + 4*(neighbours); for a
// diagonal Q_q (site-basis system-bath coupling: sigma_z, |j><j|) the products collapse to
// q_i L_q[i,j] - R_q[i,j] q_j and the kernel is a pure neighbour gather.
#include "../../include/lime_b200.h"
#include "common.cuh"
#include <algorithm>
#include <memory>

// ====================================================================================
// host: index tables
// ====================================================================================
namespace {

struct Enumerator {
    const int* dims; int nm; int exc;
    long long count = 0;
    int* out = nullptr;                  // [count][nm] or null (count only)
    std::vector<int> state;
    void rec(int idx, long long tot) {
        // lime/heom/heom.py:61: prune when `excitations` is truthy and the prefix sum exceeds it
        if (exc != 0 && tot > exc) return;
        if (idx == nm) {
            if (out) std::copy(state.begin(), state.end(), out + count * nm);
            ++count;
            return;
        }
        for (int n = 0; n < dims[idx]; ++n) {
            state[idx] = n;
            rec(idx + 1, tot + n);
        }
        state[idx] = 0;
    }
};

inline int lexcmp(const int* a, const int* b, int nm) {
    for (int k = 0; k < nm; ++k) {
        if (a[k] < b[k]) return -1;
        if (a[k] > b[k]) return 1;
    }
    return 0;
}

long long find_state(const int* states, long long nhe, int nm, const int* key) {
    long long lo = 0, hi = nhe - 1;
    while (lo <= hi) {
        long long mid = (lo + hi) >> 1;
        int c = lexcmp(states + mid * nm, key, nm);
        if (c == 0) return mid;
        if (c < 0) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

}  // namespace

extern "C" long long limeb200_heom_count_states(const int* dims, int nmodes, int excitations) {
    if (!dims || nmodes < 1 || excitations < 0) { limeb200::set_error("bad arguments"); return LB_ERR_ARG; }
    for (int k = 0; k < nmodes; ++k)
        if (dims[k] < 1) { limeb200::set_error("dims[%d] < 1", k); return LB_ERR_ARG; }
    Enumerator e{dims, nmodes, excitations};
    e.state.assign(nmodes, 0);
    e.rec(0, 0);
    return e.count;
}

extern "C" int limeb200_heom_build_tables(const int* dims, int nmodes, int excitations, long long nhe,
                                          int* states, int* dn, int* up) {
    LB_REQUIRE(dims && states && nmodes >= 1 && excitations >= 0, "bad arguments");
    LB_REQUIRE(nhe < (1LL << 31), "hierarchy too large for 32-bit tables");
    Enumerator e{dims, nmodes, excitations};
    e.state.assign(nmodes, 0);
    e.rec(0, 0);
    LB_REQUIRE(e.count == nhe, "nhe=%lld does not match the enumeration (%lld)", nhe, e.count);
    e.count = 0; e.out = states;
    e.rec(0, 0);
    if (!dn && !up) return LB_OK;
    std::vector<int> key(nmodes);
    for (long long a = 0; a < nhe; ++a) {
        const int* s = states + a * nmodes;
        long long tot = 0;
        for (int k = 0; k < nmodes; ++k) tot += s[k];
        std::copy(s, s + nmodes, key.begin());
        for (int k = 0; k < nmodes; ++k) {
            if (dn) {
                long long r = -1;
                if (s[k] >= 1) { key[k] = s[k] - 1; r = find_state(states, nhe, nmodes, key.data()); key[k] = s[k]; }
                dn[a * nmodes + k] = (int)r;
            }
            if (up) {
                long long r = -1;
                // lime/heom/heom.py:200: coupled upwards iff sum(n) <= N_c - 1
                if (tot <= (long long)excitations - 1 && s[k] + 1 < dims[k]) {
                    key[k] = s[k] + 1; r = find_state(states, nhe, nmodes, key.data()); key[k] = s[k];
                }
                up[a * nmodes + k] = (int)r;
            }
        }
    }
    return LB_OK;
}

// ====================================================================================
// device: hierarchy RK4
// ====================================================================================
struct HeomDev {
    int n, nn, nmodes, nq, npar;
    long long nhe;
    const cplx* H;          // [n][n]
    const cplx* Q;          // [nq][n][n]
    const double* qdiag;    // [nq][n] real diagonal (valid when diagq)
    int diagq;
    const int* qstart;      // [nq+1] into qmodes
    const int* qmodes;      // modes sorted by q
    const cplx* cdn;        // [npar][nmodes]  pref_dn * c_k
    const cplx* cdnR;       // [npar][nmodes]  pref_dn * conj(c_k)
    const double* nu;       // [npar][nmodes]
    cplx pref_up;
    const int* states;      // [nhe][nmodes]
    const int* dn;          // [nhe][nmodes]
    const int* up;          // [nhe][nmodes]
};

struct HeomStageArgs {
    HeomDev d;
    int B, stage;           // stage -1: plain RHS into ynext
    long long row_lo, row_hi;
    cplx* rho; const cplx* yin; cplx* ynext; cplx* acc;
    double dt;
    int apc;                // ADOs per CTA
};

// L_q / R_q element (idx) of ADO a from the neighbours in stage vector y (hierarchy base)
__device__ __forceinline__ void heom_combos(const HeomDev& d, int par, long long a, int q, int idx,
                                            const cplx* y, cplx& L, cplx& R) {
    L = cmake(0, 0); R = cmake(0, 0);
    const int* st = d.states + a * d.nmodes;
    const int* dn = d.dn + a * d.nmodes;
    const int* up = d.up + a * d.nmodes;
    for (int m = __ldg(d.qstart + q); m < __ldg(d.qstart + q + 1); ++m) {
        const int k = __ldg(d.qmodes + m);
        const int id = __ldg(dn + k);
        if (id >= 0) {
            const double nk = (double)__ldg(st + k);
            const cplx v = y[(size_t)id * d.nn + idx];
            cfma(L, cscale(nk, __ldg(d.cdn + (size_t)par * d.nmodes + k)), v);
            cfma(R, cscale(nk, __ldg(d.cdnR + (size_t)par * d.nmodes + k)), v);
        }
        const int iu = __ldg(up + k);
        if (iu >= 0) {
            const cplx v = y[(size_t)iu * d.nn + idx];
            cfma(L, d.pref_up, v);
            cfma(R, d.pref_up, v);
        }
    }
}

__device__ __forceinline__ double heom_damp(const HeomDev& d, int par, long long a) {
    double s = 0.0;
    const int* st = d.states + a * d.nmodes;
    for (int k = 0; k < d.nmodes; ++k) s = fma((double)__ldg(st + k), __ldg(d.nu + (size_t)par * d.nmodes + k), s);
    return s;
}

// -i [H, Y]_{ij} with Y a full n x n matrix at ya
__device__ __forceinline__ cplx heom_sys(const HeomDev& d, const cplx* ya, int i, int j) {
    cplx s = cmake(0, 0);
    const int n = d.n;
    for (int m = 0; m < n; ++m) {
        cfma(s, __ldg(d.H + i * n + m), ya[m * n + j]);
        cplx t = cmul(ya[i * n + m], __ldg(d.H + m * n + j));
        s.x -= t.x; s.y -= t.y;
    }
    return cmake(s.y, -s.x);          // -i * s
}

// stage-wise kernel: grid (ceil(nown/apc), B); block apc*nn threads; dynamic smem 3*apc*nn cplx
__global__ void __launch_bounds__(1024)
heom_stage_kernel(HeomStageArgs a) {
    extern __shared__ double2 smem[];
    const HeomDev& d = a.d;
    const int nn = d.nn, n = d.n;
    const int g = threadIdx.x / nn;
    const int idx = threadIdx.x - g * nn;
    const int i = idx / n, j = idx - i * n;
    const long long ado = a.row_lo + (long long)blockIdx.x * a.apc + g;
    const bool act = ado < a.row_hi;
    const int b = blockIdx.y;
    const int par = d.npar > 1 ? b : 0;
    const cplx* y = a.yin + (size_t)b * d.nhe * nn;
    cplx* ys = smem + (size_t)g * nn;
    cplx* Ls = smem + (size_t)(a.apc + g) * nn;
    cplx* Rs = smem + (size_t)(2 * a.apc + g) * nn;

    cplx yv = act ? y[(size_t)ado * nn + idx] : cmake(0, 0);
    ys[idx] = yv;
    __syncthreads();
    cplx k = cmake(0, 0);
    if (act) {
        k = heom_sys(d, ys, i, j);
        const double damp = heom_damp(d, par, ado);
        k.x = fma(-damp, yv.x, k.x);
        k.y = fma(-damp, yv.y, k.y);
    }
    for (int q = 0; q < d.nq; ++q) {
        cplx L = cmake(0, 0), R = cmake(0, 0);
        if (act) heom_combos(d, par, ado, q, idx, y, L, R);
        if (d.diagq) {
            rfma(k, __ldg(d.qdiag + q * n + i), L);
            rfma(k, -__ldg(d.qdiag + q * n + j), R);
        } else {
            __syncthreads();
            Ls[idx] = L; Rs[idx] = R;
            __syncthreads();
            const cplx* Qq = d.Q + (size_t)q * nn;
            for (int m = 0; m < n; ++m) {
                cfma(k, __ldg(Qq + i * n + m), Ls[m * n + j]);
                cplx t = cmul(Rs[i * n + m], __ldg(Qq + m * n + j));
                k.x -= t.x; k.y -= t.y;
            }
        }
    }
    if (!act) return;
    const size_t o = ((size_t)b * d.nhe + ado) * nn + idx;
    const double dt = a.dt, hdt = 0.5 * a.dt;
    if (a.stage < 0) { a.ynext[o] = k; return; }
    cplx r = a.rho[o];
    if (a.stage == 0) {
        a.acc[o] = k;
        a.ynext[o] = cmake(fma(hdt, k.x, r.x), fma(hdt, k.y, r.y));
    } else if (a.stage == 1) {
        cplx ac = a.acc[o]; rfma(ac, 2.0, k); a.acc[o] = ac;
        a.ynext[o] = cmake(fma(hdt, k.x, r.x), fma(hdt, k.y, r.y));
    } else if (a.stage == 2) {
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

// on-chip kernel: one CTA per hierarchy, all nsteps fused.  smem: y0,y1 [nhe*nn] (+ L,R if dense Q)
struct HeomChipArgs {
    HeomDev d;
    int B, nsteps, traj_every, E, T;
    cplx* ado; const cplx* eT; cplx* obs; cplx* traj;
    double dt;
};

template <int EPT>
__global__ void __launch_bounds__(1024, 1)
heom_onchip_kernel(HeomChipArgs a) {
    extern __shared__ double2 smem[];
    __shared__ cplx red[32];
    const HeomDev& d = a.d;
    const int nn = d.nn, n = d.n, T = a.T;
    const int total = (int)d.nhe * nn;
    const int b = blockIdx.x;
    const int par = d.npar > 1 ? b : 0;
    cplx* y0 = smem;
    cplx* y1 = y0 + total;
    cplx* Ls = y1 + total;
    cplx* Rs = Ls + total;
    cplx* gado = a.ado + (size_t)b * total;
    cplx rho[EPT], acc[EPT];
    double damp[EPT];
    bool ok[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        int l = threadIdx.x + e * T;
        ok[e] = l < total;
        rho[e] = ok[e] ? gado[l] : cmake(0, 0);
        acc[e] = cmake(0, 0);
        damp[e] = ok[e] ? heom_damp(d, par, l / nn) : 0.0;
        if (ok[e]) y0[l] = rho[e];
    }
    __syncthreads();
    const double dt = a.dt, hdt = 0.5 * a.dt;
    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const cplx* yin = (stage & 1) ? y1 : y0;
            cplx* yout = (stage & 1) ? y0 : y1;
            cplx k[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                k[e] = cmake(0, 0);
                if (!ok[e]) continue;
                const int l = threadIdx.x + e * T;
                const int ado = l / nn, idx = l - ado * nn;
                const int i = idx / n, j = idx - i * n;
                cplx s = heom_sys(d, yin + (size_t)ado * nn, i, j);
                const cplx yv = yin[l];
                s.x = fma(-damp[e], yv.x, s.x);
                s.y = fma(-damp[e], yv.y, s.y);
                k[e] = s;
            }
            for (int q = 0; q < d.nq; ++q) {
                cplx L[EPT], R[EPT];
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    L[e] = cmake(0, 0); R[e] = cmake(0, 0);
                    if (!ok[e]) continue;
                    const int l = threadIdx.x + e * T;
                    const int ado = l / nn, idx = l - ado * nn;
                    heom_combos(d, par, ado, q, idx, yin, L[e], R[e]);
                    if (d.diagq) {
                        const int i = idx / n, j = idx - i * n;
                        rfma(k[e], __ldg(d.qdiag + q * n + i), L[e]);
                        rfma(k[e], -__ldg(d.qdiag + q * n + j), R[e]);
                    }
                }
                if (!d.diagq) {
                    __syncthreads();
#pragma unroll
                    for (int e = 0; e < EPT; ++e)
                        if (ok[e]) { Ls[threadIdx.x + e * T] = L[e]; Rs[threadIdx.x + e * T] = R[e]; }
                    __syncthreads();
                    const cplx* Qq = d.Q + (size_t)q * nn;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) {
                        if (!ok[e]) continue;
                        const int l = threadIdx.x + e * T;
                        const int ado = l / nn, idx = l - ado * nn;
                        const int i = idx / n, j = idx - i * n;
                        const cplx* Lm = Ls + (size_t)ado * nn;
                        const cplx* Rm = Rs + (size_t)ado * nn;
                        for (int m = 0; m < n; ++m) {
                            cfma(k[e], __ldg(Qq + i * n + m), Lm[m * n + j]);
                            cplx t = cmul(Rm[i * n + m], __ldg(Qq + m * n + j));
                            k[e].x -= t.x; k[e].y -= t.y;
                        }
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                if (!ok[e]) continue;
                cplx yn;
                if (stage == 0) {
                    acc[e] = k[e];
                    yn = cmake(fma(hdt, k[e].x, rho[e].x), fma(hdt, k[e].y, rho[e].y));
                } else if (stage == 1) {
                    rfma(acc[e], 2.0, k[e]);
                    yn = cmake(fma(hdt, k[e].x, rho[e].x), fma(hdt, k[e].y, rho[e].y));
                } else if (stage == 2) {
                    rfma(acc[e], 2.0, k[e]);
                    yn = cmake(fma(dt, k[e].x, rho[e].x), fma(dt, k[e].y, rho[e].y));
                } else {
                    cplx tot = cadd(acc[e], k[e]);
                    rho[e].x += tot.x / 6.0 * dt;
                    rho[e].y += tot.y / 6.0 * dt;
                    yn = rho[e];
                }
                yout[threadIdx.x + e * T] = yn;
            }
            __syncthreads();
        }
        // y0 holds the new hierarchy; tier 0 is ADO 0
        if (a.obs) {
            for (int eo = 0; eo < a.E; ++eo) {
                cplx v = cmake(0, 0);
                for (int l = threadIdx.x; l < nn; l += T) cfma(v, __ldg(a.eT + (size_t)eo * nn + l), y0[l]);
                for (int off = 16; off > 0; off >>= 1) {
                    v.x += __shfl_down_sync(0xffffffffu, v.x, off);
                    v.y += __shfl_down_sync(0xffffffffu, v.y, off);
                }
                if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
                __syncthreads();
                if (threadIdx.x == 0) {
                    cplx s = cmake(0, 0);
                    for (int w = 0; w < (T + 31) / 32; ++w) s = cadd(s, red[w]);
                    a.obs[((size_t)step * a.B + b) * a.E + eo] = s;
                }
                __syncthreads();
            }
        }
        if (a.traj && ((step + 1) % a.traj_every) == 0) {
            cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * nn;
            for (int l = threadIdx.x; l < nn; l += T) dst[l] = y0[l];
        }
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e)
        if (ok[e]) gado[threadIdx.x + e * T] = rho[e];
}

// tier-0 observables / trajectory for the stage-wise path
__global__ void heom_tier0_obs(const cplx* __restrict__ ado, long long hier_stride, int nn,
                               const cplx* __restrict__ eT, int E, cplx* obs, cplx* traj) {
    const int b = blockIdx.x;
    const cplx* r = ado + (size_t)b * hier_stride;
    if (traj)
        for (int l = threadIdx.x; l < nn; l += blockDim.x) traj[(size_t)b * nn + l] = r[l];
    if (obs && threadIdx.x < E) {
        cplx s = cmake(0, 0);
        for (int l = 0; l < nn; ++l) cfma(s, eT[(size_t)threadIdx.x * nn + l], r[l]);
        obs[(size_t)b * E + threadIdx.x] = s;
    }
}

// _heom_dl-exact: lime/oqs.py:1846-1857.  One CTA per hierarchy, nn threads.
struct HeomDlArgs {
    const cplx* H; const cplx* sz; int n, nado, B, nt;
    cplx* ado; const double* par; cplx* traj; double dt;
};

__device__ __forceinline__ cplx comm_elem(const cplx* A, const cplx* Y, int n, int i, int j) {   // (A Y - Y A)_{ij}
    cplx s = cmake(0, 0);
    for (int m = 0; m < n; ++m) {
        cfma(s, A[i * n + m], Y[m * n + j]);
        cplx t = cmul(Y[i * n + m], A[m * n + j]);
        s.x -= t.x; s.y -= t.y;
    }
    return s;
}
__device__ __forceinline__ cplx acomm_elem(const cplx* A, const cplx* Y, int n, int i, int j) {
    cplx s = cmake(0, 0);
    for (int m = 0; m < n; ++m) {
        cfma(s, A[i * n + m], Y[m * n + j]);
        cfma(s, Y[i * n + m], A[m * n + j]);
    }
    return s;
}

__global__ void __launch_bounds__(1024)
heom_dl_kernel(HeomDlArgs a) {
    extern __shared__ double2 smem[];
    const int n = a.n, nn = n * n, nado = a.nado;
    const int b = blockIdx.x;
    cplx* ado = smem;                       // [nado][nn]
    cplx* Hs = ado + (size_t)nado * nn;
    cplx* Ss = Hs + nn;
    const int idx = threadIdx.x;
    const int i = idx / n, j = idx - i * n;
    cplx* g = a.ado + (size_t)b * nado * nn;
    for (int t = 0; t < nado; ++t) ado[t * nn + idx] = g[t * nn + idx];
    Hs[idx] = a.H[idx]; Ss[idx] = a.sz[idx];
    const double gamma = a.par[b * 3 + 0], ca = a.par[b * 3 + 1], cb = a.par[b * 3 + 2];
    const double dt = a.dt;
    __syncthreads();
    for (int step = 0; step < a.nt; ++step) {
        // tier 0, first update (lime/oqs.py:1850-1851)
        {
            cplx c0 = comm_elem(Hs, ado, n, i, j);
            cplx c1 = comm_elem(Ss, ado + nn, n, i, j);
            cplx v = ado[idx];
            // += -1j*c0*dt - c1*dt
            v.x += c0.y * dt - c1.x * dt;
            v.y += -c0.x * dt - c1.y * dt;
            __syncthreads();
            ado[idx] = v;
            __syncthreads();
        }
        for (int t = 0; t < nado - 1; ++t) {
            const cplx* cur = ado + (size_t)t * nn;
            const cplx* nxt = ado + (size_t)(t + 1) * nn;
            // n = 0 multiplies the lower-tier term by zero; Python reads ado[:,:,-1] (the
            // last tier) there, which is harmless for finite values
            const cplx* prv = ado + (size_t)((t == 0) ? (nado - 1) : (t - 1)) * nn;
            cplx c0 = comm_elem(Hs, cur, n, i, j);
            cplx c1 = comm_elem(Ss, nxt, n, i, j);
            cplx c2 = comm_elem(Ss, prv, n, i, j);
            cplx a2 = acomm_elem(Ss, prv, n, i, j);
            cplx v = cur[idx];
            const double tn = (double)t;
            // inner = -c1 - n*gamma*cur + n*(a*c2 + 1j*b*a2)
            cplx inner;
            inner.x = -c1.x - tn * gamma * v.x + tn * (ca * c2.x - cb * a2.y);
            inner.y = -c1.y - tn * gamma * v.y + tn * (ca * c2.y + cb * a2.x);
            v.x += c0.y * dt + inner.x * dt;
            v.y += -c0.x * dt + inner.y * dt;
            __syncthreads();
            ado[(size_t)t * nn + idx] = v;
            __syncthreads();
        }
        if (a.traj) a.traj[((size_t)step * a.B + b) * nn + idx] = ado[idx];
    }
    for (int t = 0; t < nado; ++t) g[t * nn + idx] = ado[t * nn + idx];
}

// ====================================================================================
// plan
// ====================================================================================
struct limeb200_heom_s {
    int device = 0;
    int n = 0, nmodes = 0, nq = 0, npar = 1;
    long long nhe = 0, row_lo = 0, row_hi = 0;
    int path_req = 0, path = 0;
    bool diagq = false;
    cplx pref_up;
    DevBuf dH, dQ, dqdiag, dqstart, dqmodes, dcdn, dcdnR, dnu, dstates, ddn, dup;
    DevBuf s_y, s_acc;
    int scratch_B = 0;
    long long launches = 0;
    long long smem_optin = 0;
    int sm_count = 148;
    HeomDev dev() const {
        HeomDev d;
        d.n = n; d.nn = n * n; d.nmodes = nmodes; d.nq = nq; d.npar = npar; d.nhe = nhe;
        d.H = dH.as<cplx>(); d.Q = dQ.as<cplx>(); d.qdiag = dqdiag.as<double>(); d.diagq = diagq ? 1 : 0;
        d.qstart = dqstart.as<int>(); d.qmodes = dqmodes.as<int>();
        d.cdn = dcdn.as<cplx>(); d.cdnR = dcdnR.as<cplx>(); d.nu = dnu.as<double>();
        d.pref_up = pref_up;
        d.states = dstates.as<int>(); d.dn = ddn.as<int>(); d.up = dup.as<int>();
        return d;
    }
};

extern "C" {

int limeb200_heom_create_batched(limeb200_heom_t* plan, int device, int n, int nmodes, int nq, long long nhe,
                                 const double* h_H, const double* h_Q, const int* qmap,
                                 const double* h_c, const double* h_nu, int npar,
                                 const double* pref_dn, const double* pref_up,
                                 const int* states, const int* dn, const int* up,
                                 long long row_lo, long long row_hi);

int limeb200_heom_create(limeb200_heom_t* plan, int device, int n, int nmodes, int nq, long long nhe,
                         const double* h_H, const double* h_Q, const int* qmap,
                         const double* h_c, const double* h_nu,
                         const double* pref_dn, const double* pref_up,
                         const int* states, const int* dn, const int* up,
                         long long row_lo, long long row_hi) {
    return limeb200_heom_create_batched(plan, device, n, nmodes, nq, nhe, h_H, h_Q, qmap, h_c, h_nu, 1,
                                        pref_dn, pref_up, states, dn, up, row_lo, row_hi);
}

int limeb200_heom_create_batched(limeb200_heom_t* plan, int device, int n, int nmodes, int nq, long long nhe,
                                 const double* h_H, const double* h_Q, const int* qmap,
                                 const double* h_c, const double* h_nu, int npar,
                                 const double* pref_dn, const double* pref_up,
                                 const int* states, const int* dn, const int* up,
                                 long long row_lo, long long row_hi) {
    LB_REQUIRE(plan && h_H && h_Q && qmap && h_c && h_nu && pref_dn && pref_up && states && dn && up, "null argument");
    LB_REQUIRE(n >= 1 && n <= 32, "system dimension n=%d out of range (1..32)", n);
    LB_REQUIRE(nmodes >= 1 && nq >= 1 && nhe >= 1 && npar >= 1, "bad sizes");
    LB_REQUIRE(nhe < (1LL << 31), "hierarchy too large");
    LB_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= nhe, "bad ADO range");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        limeb200::set_error("no CUDA device available (%s): liblime_b200 has no CPU fallback",
                            ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
        return LB_ERR_CUDA;
    }
    LB_REQUIRE(device >= 0 && device < ndev, "device %d out of range", device);
    for (int k = 0; k < nmodes; ++k) LB_REQUIRE(qmap[k] >= 0 && qmap[k] < nq, "qmap[%d] out of range", k);
    LB_CUDA(cudaSetDevice(device));
    std::unique_ptr<limeb200_heom_s> p(new limeb200_heom_s());
    p->device = device; p->n = n; p->nmodes = nmodes; p->nq = nq; p->npar = npar; p->nhe = nhe;
    p->row_lo = row_lo; p->row_hi = row_hi;
    cudaDeviceProp prop;
    LB_CUDA(cudaGetDeviceProperties(&prop, device));
    p->smem_optin = (long long)prop.sharedMemPerBlockOptin;
    p->sm_count = prop.multiProcessorCount;
    const int nn = n * n;
    const hcplx* Q = reinterpret_cast<const hcplx*>(h_Q);
    bool diag = true;
    for (int q = 0; q < nq && diag; ++q)
        for (int i = 0; i < n && diag; ++i)
            for (int j = 0; j < n; ++j) {
                hcplx v = Q[(size_t)q * nn + i * n + j];
                if ((i != j && v != hcplx(0, 0)) || (i == j && v.imag() != 0.0)) { diag = false; break; }
            }
    p->diagq = diag;
    std::vector<double> qd((size_t)nq * n, 0.0);
    for (int q = 0; q < nq; ++q)
        for (int i = 0; i < n; ++i) qd[q * n + i] = Q[(size_t)q * nn + i * n + i].real();
    std::vector<int> qstart(nq + 1, 0), qmodes;
    for (int q = 0; q < nq; ++q) {
        for (int k = 0; k < nmodes; ++k)
            if (qmap[k] == q) qmodes.push_back(k);
        qstart[q + 1] = (int)qmodes.size();
    }
    const hcplx pd(pref_dn[0], pref_dn[1]);
    const hcplx* c = reinterpret_cast<const hcplx*>(h_c);
    std::vector<hcplx> cdn((size_t)npar * nmodes), cdnR((size_t)npar * nmodes);
    for (size_t i = 0; i < cdn.size(); ++i) { cdn[i] = pd * c[i]; cdnR[i] = pd * std::conj(c[i]); }
    p->pref_up = cmake(pref_up[0], pref_up[1]);
    LB_CUDA(p->dH.upload(h_H, (size_t)nn * 16));
    LB_CUDA(p->dQ.upload(h_Q, (size_t)nq * nn * 16));
    LB_CUDA(p->dqdiag.upload(qd.data(), qd.size() * 8));
    LB_CUDA(p->dqstart.upload(qstart.data(), qstart.size() * 4));
    LB_CUDA(p->dqmodes.upload(qmodes.data(), qmodes.size() * 4));
    LB_CUDA(p->dcdn.upload(cdn.data(), cdn.size() * 16));
    LB_CUDA(p->dcdnR.upload(cdnR.data(), cdnR.size() * 16));
    LB_CUDA(p->dnu.upload(h_nu, (size_t)npar * nmodes * 8));
    LB_CUDA(p->dstates.upload(states, (size_t)nhe * nmodes * 4));
    LB_CUDA(p->ddn.upload(dn, (size_t)nhe * nmodes * 4));
    LB_CUDA(p->dup.upload(up, (size_t)nhe * nmodes * 4));
    *plan = p.release();
    return LB_OK;
}

int limeb200_heom_destroy(limeb200_heom_t p) {
    if (p) { cudaSetDevice(p->device); delete p; }
    return LB_OK;
}
int limeb200_heom_set_path(limeb200_heom_t p, int path) {
    LB_REQUIRE(p && path >= 0 && path <= 2, "bad arguments");
    p->path_req = path;
    return LB_OK;
}
int limeb200_heom_get_path(limeb200_heom_t p) { return p ? p->path : LB_ERR_ARG; }
long long limeb200_heom_last_launches(limeb200_heom_t p) { return p ? p->launches : -1; }

static int heom_launch_stage(limeb200_heom_t p, int stage, cplx* rho, const cplx* yin, cplx* ynext, cplx* acc,
                             int B, double dt, cudaStream_t st) {
    const long long nown = p->row_hi - p->row_lo;
    if (nown <= 0) return LB_OK;
    const int nn = p->n * p->n;
    HeomStageArgs a;
    a.d = p->dev();
    a.B = B; a.stage = stage; a.row_lo = p->row_lo; a.row_hi = p->row_hi;
    a.rho = rho; a.yin = yin; a.ynext = ynext; a.acc = acc; a.dt = dt;
    a.apc = std::max(1, 256 / nn);
    size_t smem = (size_t)3 * a.apc * nn * 16;
    dim3 grid((unsigned)ceil_div(nown, (long long)a.apc), B);
    if (smem > 48 * 1024)
        LB_CUDA(cudaFuncSetAttribute(heom_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    heom_stage_kernel<<<grid, a.apc * nn, smem, st>>>(a);
    p->launches++;
    return LB_OK;
}

int limeb200_heom_stage(limeb200_heom_t p, int stage, double* d_rho, const double* d_yin,
                        double* d_ynext, double* d_acc, int B, double dt, void* stream) {
    LB_REQUIRE(p && d_rho && d_yin && d_ynext && d_acc, "null argument");
    LB_REQUIRE(stage >= 0 && stage <= 3 && B >= 1, "bad stage/B");
    LB_REQUIRE(p->npar == 1 || p->npar == B, "parameter batch %d != B %d", p->npar, B);
    LB_CUDA(cudaSetDevice(p->device));
    int r = heom_launch_stage(p, stage, (cplx*)d_rho, (const cplx*)d_yin, (cplx*)d_ynext, (cplx*)d_acc, B, dt,
                              (cudaStream_t)stream);
    if (r != LB_OK) return r;
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_heom_rhs(limeb200_heom_t p, const double* d_in, double* d_out, int B, void* stream) {
    LB_REQUIRE(p && d_in && d_out && B >= 1, "bad arguments");
    LB_REQUIRE(p->npar == 1 || p->npar == B, "parameter batch %d != B %d", p->npar, B);
    LB_CUDA(cudaSetDevice(p->device));
    p->launches = 0;
    int r = heom_launch_stage(p, -1, nullptr, (const cplx*)d_in, (cplx*)d_out, nullptr, B, 0.0, (cudaStream_t)stream);
    if (r != LB_OK) return r;
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_heom_run(limeb200_heom_t p, double* d_ado, int B, double dt, int nsteps,
                      const double* d_eT, int E, double* d_obs, double* d_traj, int traj_every,
                      void* stream) {
    LB_REQUIRE(p && d_ado && B >= 1 && nsteps >= 0 && E >= 0, "bad arguments");
    LB_REQUIRE(p->npar == 1 || p->npar == B, "parameter batch %d != B %d", p->npar, B);
    LB_REQUIRE(E == 0 || (d_eT && d_obs), "observables requested without buffers");
    LB_REQUIRE(E <= 32, "at most 32 observables");
    LB_REQUIRE(!d_traj || traj_every >= 1, "traj_every must be >= 1");
    LB_REQUIRE(p->row_lo == 0 && p->row_hi == p->nhe, "heom_run needs a plan that owns the whole hierarchy");
    LB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    p->launches = 0;
    if (nsteps == 0) return LB_OK;
    if (!d_traj) traj_every = 1;
    const int nn = p->n * p->n;
    const long long total = p->nhe * nn;
    // ---- on-chip path: whole hierarchy in one CTA's shared memory
    const int nbuf = p->diagq ? 2 : 4;
    const size_t smem_chip = (size_t)nbuf * total * 16;
    bool chip_ok = total <= 8 * 1024 && smem_chip + 1024 <= (size_t)p->smem_optin;
    int path = p->path_req;
    if (path == 0) path = (chip_ok && (B >= p->sm_count / 2 || total <= 4096)) ? 1 : 2;
    LB_REQUIRE(path != 1 || chip_ok, "hierarchy (%lld elements) does not fit the on-chip path", total);
    p->path = path;
    if (path == 1) {
        HeomChipArgs a;
        a.d = p->dev();
        a.B = B; a.nsteps = nsteps; a.traj_every = traj_every; a.E = E;
        a.ado = (cplx*)d_ado; a.eT = (const cplx*)d_eT; a.obs = E > 0 ? (cplx*)d_obs : nullptr; a.traj = (cplx*)d_traj;
        a.dt = dt;
        int T = (int)std::min<long long>(1024, ceil_div(total, 32LL) * 32);
        int ept = (int)ceil_div(total, (long long)T);
        int EPT = ept <= 1 ? 1 : ept <= 2 ? 2 : ept <= 4 ? 4 : 8;
        a.T = T;
        void (*kern)(HeomChipArgs) = EPT == 1 ? heom_onchip_kernel<1> : EPT == 2 ? heom_onchip_kernel<2>
                                   : EPT == 4 ? heom_onchip_kernel<4> : heom_onchip_kernel<8>;
        LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chip));
        kern<<<B, T, smem_chip, st>>>(a);
        LB_CUDA(cudaGetLastError());
        p->launches++;
        return LB_OK;
    }
    // ---- stage-wise path
    if (B > p->scratch_B || !p->s_y.p) {
        LB_CUDA(p->s_y.alloc((size_t)2 * B * total * 16));
        LB_CUDA(p->s_acc.alloc((size_t)B * total * 16));
        p->scratch_B = B;
    }
    cplx* rho = (cplx*)d_ado;
    cplx* y[2] = {p->s_y.as<cplx>(), p->s_y.as<cplx>() + (size_t)B * total};
    cplx* acc = p->s_acc.as<cplx>();
    LB_CUDA(cudaMemcpyAsync(y[0], rho, (size_t)B * total * 16, cudaMemcpyDeviceToDevice, st));
    for (int step = 0; step < nsteps; ++step) {
        for (int stage = 0; stage < 4; ++stage) {
            int r = heom_launch_stage(p, stage, rho, y[stage & 1], y[(stage + 1) & 1], acc, B, dt, st);
            if (r != LB_OK) return r;
        }
        const bool save = d_traj && ((step + 1) % traj_every) == 0;
        if (E > 0 || save) {
            heom_tier0_obs<<<B, 64, 0, st>>>(rho, total, nn, (const cplx*)d_eT, E,
                                            E > 0 ? (cplx*)d_obs + (size_t)step * B * E : nullptr,
                                            save ? (cplx*)d_traj + (size_t)(step / traj_every) * B * nn : nullptr);
            p->launches++;
        }
    }
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

int limeb200_heom_dl_euler(const double* h_H, const double* h_sz, int n, int nado,
                           double* d_ado, const double* d_par, int B, double dt, int nt,
                           double* d_traj, void* stream) {
    LB_REQUIRE(h_H && h_sz && d_ado && d_par, "null argument");
    LB_REQUIRE(n >= 1 && n <= 32 && nado >= 2 && B >= 1 && nt >= 0, "bad sizes (n<=32, nado>=2)");
    const int nn = n * n;
    size_t smem = ((size_t)nado * nn + 2 * nn) * 16;
    LB_REQUIRE(smem <= 200 * 1024, "hierarchy too deep for the on-chip _heom_dl kernel");
    DevBuf dH, dS;
    LB_CUDA(dH.upload(h_H, (size_t)nn * 16));
    LB_CUDA(dS.upload(h_sz, (size_t)nn * 16));
    HeomDlArgs a;
    a.H = dH.as<cplx>(); a.sz = dS.as<cplx>(); a.n = n; a.nado = nado; a.B = B; a.nt = nt;
    a.ado = (cplx*)d_ado; a.par = d_par; a.traj = (cplx*)d_traj; a.dt = dt;
    if (smem > 48 * 1024)
        LB_CUDA(cudaFuncSetAttribute(heom_dl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    heom_dl_kernel<<<B, nn, smem, (cudaStream_t)stream>>>(a);
    LB_CUDA(cudaGetLastError());
    LB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));     // dH/dS are freed on return
    return LB_OK;
}

}  // extern "C"
