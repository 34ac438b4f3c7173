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
#include <cstdlib>
#include <memory>
#include <cooperative_groups.h>

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
    int diagq;              // every Q_q is real diagonal
    const int* qstart;      // [nq+1] into qmodes              (dense-Q path)
    const int* qmodes;      // modes sorted by q
    // diagonal-Q path: per matrix element (i,j) the modes k with (Q_qk)_ii != 0 or (Q_qk)_jj != 0
    const int* em_start;    // [nn+1]
    const int* em_mode;     // mode index
    const double2* em_v;    // ((Q_q)_ii, (Q_q)_jj)
    const cplx* cdn;        // [npar][nmodes]  pref_dn * c_k
    const cplx* cdnR;       // [npar][nmodes]  pref_dn * conj(c_k)
    const double* nu;       // [npar][nmodes]
    const double* damp;     // [nhe] sum_k n_k nu_k, or null when npar > 1
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
    int npeer;              // sharded persistent kernels: the new stage value is also stored into
    cplx* peer_next[7];     // the same element of every peer GPU's stage vector (NVLink stores)
    const unsigned char* peer_mask;   // [nhe] bit q set: peer slot q reads this ADO (null or send_all: every peer)
    int send_all;
    // hybrid sharding (tagged halo): rows owned by OTHER ranks are not in yin but in a tagged inbox -- entry e = two
    // 16-byte words {value bits, 64-bit stage tag} written by the owner over NVLink -- and are polled there; new values
    // go to the peers' inboxes as tagged entries instead of plain elements.  Null tag_in: everything is in yin.
    const ulonglong2* tag_in;
    ulonglong2* tag_out_peer[7];
    ulonglong2* tag_out_own;          // own inbox, written at the last stage of a run only (uniform unpack)
    unsigned long long tag_want;      // tag of the stage input
    unsigned* tag_err;
    // packed neighbour lists (built once per plan, see heom_pack_kernel): per element 8 words entry | n_k << 26 and the
    // finished coefficients ctab[par][idx][slot][n_k]; null: walk the index tables
    const uint4* pk;
    const cplx* ctab;
    int NK;
};
#define HEOM_PK_NE 8
#define HEOM_PK_IDXBITS 26
// coefficient table layout ctab[par][slot][nk][idx], matrix element fastest (see heom_ctab_kernel)
__device__ __forceinline__ size_t heom_ctab_index(int par, int slot, int nk, int idx, int NK, int nn) {
    return (((size_t)par * HEOM_PK_NE + slot) * NK + nk) * nn + idx;
}

// L_q / R_q element (idx) of ADO a from the neighbours in stage vector y (dense-Q path)
// st/dn/up: the ADO's rows of the index tables (global memory, or a shared-memory copy)
__device__ __forceinline__ void heom_combos(const HeomDev& d, int par, const int* st, const int* dn, const int* up,
                                            int q, int idx, const cplx* y, cplx& L, cplx& R) {
    L = cmake(0, 0); R = cmake(0, 0);
    for (int m = __ldg(d.qstart + q); m < __ldg(d.qstart + q + 1); ++m) {
        const int k = __ldg(d.qmodes + m);
        const int id = dn[k];
        if (id >= 0) {
            const double nk = (double)st[k];
            const cplx v = y[(size_t)id * d.nn + idx];
            cfma(L, cscale(nk, __ldg(d.cdn + (size_t)par * d.nmodes + k)), v);
            cfma(R, cscale(nk, __ldg(d.cdnR + (size_t)par * d.nmodes + k)), v);
        }
        const int iu = up[k];
        if (iu >= 0) {
            const cplx v = y[(size_t)iu * d.nn + idx];
            cfma(L, d.pref_up, v);
            cfma(R, d.pref_up, v);
        }
    }
}

// coefficient of rho_{n-e_k}[i][j] in d rho_n[i][j]/dt for diagonal Q:  n_k (q_i pref_dn c_k - q_j pref_dn conj c_k)
__device__ __forceinline__ cplx heom_dn_coef(const HeomDev& d, int par, int k, double nk, double2 v) {
    const cplx cl = __ldg(d.cdn + (size_t)par * d.nmodes + k);
    const cplx cr = __ldg(d.cdnR + (size_t)par * d.nmodes + k);
    return cmake(nk * (v.x * cl.x - v.y * cr.x), nk * (v.x * cl.y - v.y * cr.y));
}

// bath part of the right-hand side, diagonal Q: a pure neighbour gather over the element's own mode list
__device__ __forceinline__ void heom_bath_diag(const HeomDev& d, int par, const int* st, const int* dn,
                                               const int* up, int idx, const cplx* y, cplx& k) {
    const int t1 = __ldg(d.em_start + idx + 1);
#pragma unroll 4
    for (int t = __ldg(d.em_start + idx); t < t1; ++t) {
        const int m = __ldg(d.em_mode + t);
        const double2 v = __ldg(d.em_v + t);
        const int id = dn[m], iu = up[m];
        // clamp instead of branching so that both neighbour loads are issued back to back
        const cplx yd = y[(size_t)max(id, 0) * d.nn + idx];
        const cplx yu = y[(size_t)max(iu, 0) * d.nn + idx];
        if (id >= 0) cfma(k, heom_dn_coef(d, par, m, (double)st[m], v), yd);
        if (iu >= 0) cfma(k, cscale(v.x - v.y, d.pref_up), yu);
    }
}

// value of element `e` of the stage input: from yin when its ADO is owned by this rank, else from the tagged inbox
// (polled until the owner's entry for this stage has arrived; bounded, sets *err on time-out)
__device__ __forceinline__ cplx heom_read_halo(const HeomStageArgs& a, const cplx* y, long long ado_nb, size_t e) {
    if (!a.tag_in || (ado_nb >= a.row_lo && ado_nb < a.row_hi)) return y[e];
    const ulonglong2* p = a.tag_in + 2 * e;
    ulonglong2 w0, w1;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (true) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0.x), "=l"(w0.y) : "l"(p) : "memory");
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w1.x), "=l"(w1.y) : "l"(p + 1) : "memory");
        if (w0.y == a.tag_want && w1.y == a.tag_want) break;
        if (((++spins & 1023u) == 0) && (clock64() - t0 > 10000000000LL || *(volatile unsigned*)a.tag_err)) {
            atomicExch(a.tag_err, 1u);
            break;
        }
    }
    return cmake(__longlong_as_double((long long)w0.x), __longlong_as_double((long long)w1.x));
}
__device__ __forceinline__ void heom_store_tagged(ulonglong2* T, size_t e, cplx v, unsigned long long tag) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(T + 2 * e), "l"((unsigned long long)__double_as_longlong(v.x)), "l"(tag) : "memory");
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(T + 2 * e + 1), "l"((unsigned long long)__double_as_longlong(v.y)), "l"(tag) : "memory");
}
// bath part with a tagged halo: as heom_bath_diag, but every existing neighbour is READ (and so waited for) even when
// its coefficient vanishes -- that keeps the read graph symmetric, which is what makes the inbox ping-pong safe
__device__ __forceinline__ void heom_bath_diag_halo(const HeomStageArgs& a, int par, const int* st, const int* dn,
                                                    const int* up, int idx, const cplx* y, cplx& k) {
    const HeomDev& d = a.d;
    const int t1 = __ldg(d.em_start + idx + 1);
    for (int t = __ldg(d.em_start + idx); t < t1; ++t) {
        const int m = __ldg(d.em_mode + t);
        const double2 v = __ldg(d.em_v + t);
        const int id = dn[m], iu = up[m];
        if (id >= 0) cfma(k, heom_dn_coef(d, par, m, (double)st[m], v), heom_read_halo(a, y, id, (size_t)id * d.nn + idx));
        if (iu >= 0) cfma(k, cscale(v.x - v.y, d.pref_up), heom_read_halo(a, y, iu, (size_t)iu * d.nn + idx));
    }
}

__device__ __forceinline__ double heom_damp(const HeomDev& d, int par, long long a) {
    if (d.damp) return __ldg(d.damp + a);
    double s = 0.0;
    const int* st = d.states + a * d.nmodes;
    for (int k = 0; k < d.nmodes; ++k) s = fma((double)__ldg(st + k), __ldg(d.nu + (size_t)par * d.nmodes + k), s);
    return s;
}

// -i [H, Y]_{ij} with Y a full n x n matrix at ya (shared memory), Hs the system Hamiltonian
__device__ __forceinline__ cplx heom_sys(const cplx* Hs, int n, const cplx* ya, int i, int j) {
    cplx s = cmake(0, 0);
    for (int m = 0; m < n; ++m) {
        cfma(s, Hs[i * n + m], ya[m * n + j]);
        cplx t = cmul(ya[i * n + m], Hs[m * n + j]);
        s.x -= t.x; s.y -= t.y;
    }
    return cmake(s.y, -s.x);          // -i * s
}

// same with the dimension known at compile time (fully unrolled, strength-reduced addressing)
template <int NN_>
__device__ __forceinline__ cplx heom_sys_t(const cplx* Hs, const cplx* ya, int i, int j) {
    cplx s = cmake(0, 0);
    const cplx* hrow = Hs + i * NN_;
    const cplx* yrow = ya + i * NN_;
#pragma unroll
    for (int m = 0; m < NN_; ++m) {
        cfma(s, hrow[m], ya[m * NN_ + j]);
        const cplx a = yrow[m], h = Hs[m * NN_ + j];
        s.x = fma(-a.x, h.x, s.x); s.x = fma(a.y, h.y, s.x);
        s.y = fma(-a.x, h.y, s.y); s.y = fma(-a.y, h.x, s.y);
    }
    return cmake(s.y, -s.x);          // -i * s
}
__device__ __forceinline__ cplx heom_sys_any(const cplx* Hs, int n, const cplx* ya, int i, int j) {
    switch (n) {
        case 2: return heom_sys_t<2>(Hs, ya, i, j);
        case 3: return heom_sys_t<3>(Hs, ya, i, j);
        case 7: return heom_sys_t<7>(Hs, ya, i, j);
        default: return heom_sys(Hs, n, ya, i, j);
    }
}

// RK4 stage algebra (lime/phys.py:636-649) for one element; returns the next stage vector element
__device__ __forceinline__ cplx heom_rk_update(int stage, cplx k, cplx& r, cplx& ac, double dt) {
    const double hdt = 0.5 * dt;
    if (stage == 0) { ac = k; return cmake(fma(hdt, k.x, r.x), fma(hdt, k.y, r.y)); }
    if (stage == 1) { rfma(ac, 2.0, k); return cmake(fma(hdt, k.x, r.x), fma(hdt, k.y, r.y)); }
    if (stage == 2) { rfma(ac, 2.0, k); return cmake(fma(dt, k.x, r.x), fma(dt, k.y, r.y)); }
    const double w6 = dt / 6.0;
    r.x = fma(w6, ac.x + k.x, r.x);
    r.y = fma(w6, ac.y + k.y, r.y);
    return r;
}

// one tile (apc ADOs x nn elements) of one RK4 stage; shared: Hs[nn], ys/Ls/Rs[apc*nn].
// Items are the owned ADOs of all B hierarchies in flat order: item = b * nown + (ado - row_lo).
// tabs: shared-memory copy [apc][3][nmodes] of the tile's (states, dn, up) rows, or null (read global).
// rreg/areg: the element's rho / RK accumulator kept in registers by a persistent caller, or null.
__device__ __forceinline__ void heom_stage_tile(const HeomStageArgs& a, long long item0, double2* smem,
                                                const int* tabs, cplx* rreg, cplx* areg) {
    const HeomDev& d = a.d;
    const int nn = d.nn, n = d.n;
    const int g = threadIdx.x / nn;
    const int idx = threadIdx.x - g * nn;
    const int i = idx / n, j = idx - i * n;
    const long long nown = a.row_hi - a.row_lo;
    const long long item = item0 + g;
    const bool act = g < a.apc && item < nown * a.B;
    int b = 0;
    if (act && a.B > 1)         // 32-bit division whenever the flat index fits (a 64-bit one costs ~100 instructions)
        b = (nown * a.B < (1LL << 31)) ? (int)((unsigned)item / (unsigned)nown) : (int)(item / nown);
    const long long ado = act ? a.row_lo + (item - (long long)b * nown) : a.row_lo;
    const int par = d.npar > 1 ? b : 0;
    const cplx* y = a.yin + (size_t)b * d.nhe * nn;
    cplx* Hs = smem;
    cplx* ys = Hs + nn + (size_t)g * nn;
    cplx* Ls = Hs + nn + (size_t)(a.apc + g) * nn;
    cplx* Rs = Hs + nn + (size_t)(2 * a.apc + g) * nn;

    cplx yv = act ? y[(size_t)ado * nn + idx] : cmake(0, 0);
    if (g < a.apc) ys[idx] = yv;
    __syncthreads();
    cplx k = cmake(0, 0);
    if (act) {
        k = heom_sys_any(Hs, n, ys, i, j);
        const double damp = heom_damp(d, par, ado);
        k.x = fma(-damp, yv.x, k.x);
        k.y = fma(-damp, yv.y, k.y);
    }
    const int nm = d.nmodes;
    const int* st = tabs ? tabs + (size_t)(3 * g) * nm : d.states + ado * nm;
    const int* dn = tabs ? tabs + (size_t)(3 * g + 1) * nm : d.dn + ado * nm;
    const int* up = tabs ? tabs + (size_t)(3 * g + 2) * nm : d.up + ado * nm;
    if (d.diagq) {
        if (act) {
            if (a.tag_in) {
                heom_bath_diag_halo(a, par, st, dn, up, idx, y, k);
            } else if (a.pk) {
                // packed gather: same neighbours, same coefficients, same order as heom_bath_diag
                const size_t e = (size_t)ado * nn + idx;
                const uint4 p0 = __ldg(a.pk + 2 * e), p1 = __ldg(a.pk + 2 * e + 1);
                const unsigned w[HEOM_PK_NE] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                cplx nb[HEOM_PK_NE];
#pragma unroll
                for (int s = 0; s < HEOM_PK_NE; ++s) nb[s] = y[w[s] & ((1u << HEOM_PK_IDXBITS) - 1u)];
#pragma unroll
                for (int s = 0; s < HEOM_PK_NE; ++s)
                    cfma(k, __ldg(a.ctab + heom_ctab_index(par, s, (int)(w[s] >> HEOM_PK_IDXBITS), idx, a.NK, nn)), nb[s]);
            } else {
                heom_bath_diag(d, par, st, dn, up, idx, y, k);
            }
        }
    } else {
        for (int q = 0; q < d.nq; ++q) {
            cplx L = cmake(0, 0), R = cmake(0, 0);
            if (act) heom_combos(d, par, st, dn, up, q, idx, y, L, R);
            __syncthreads();
            if (g < a.apc) { Ls[idx] = L; Rs[idx] = R; }
            __syncthreads();
            if (act) {
                const cplx* Qq = d.Q + (size_t)q * nn;
                for (int m = 0; m < n; ++m) {
                    cfma(k, __ldg(Qq + i * n + m), Ls[m * n + j]);
                    cplx t = cmul(Rs[i * n + m], __ldg(Qq + m * n + j));
                    k.x -= t.x; k.y -= t.y;
                }
            }
        }
    }
    if (!act) return;
    const size_t o = ((size_t)b * d.nhe + ado) * nn + idx;
    if (a.stage < 0) { a.ynext[o] = k; return; }
    if (rreg) {
        const cplx yn = heom_rk_update(a.stage, k, *rreg, *areg, a.dt);
        a.ynext[o] = yn;
        if (a.npeer) {
            const unsigned m = (a.send_all || !a.peer_mask) ? 0xffu : a.peer_mask[ado];
            for (int r = 0; r < a.npeer; ++r)
                if ((m >> r) & 1u) {
                    if (a.tag_in) heom_store_tagged(a.tag_out_peer[r], o, yn, a.tag_want + 1);
                    else a.peer_next[r][o] = yn;
                }
            if (a.tag_in && a.send_all) heom_store_tagged(a.tag_out_own, o, yn, a.tag_want + 1);
        }
        if (a.stage == 3) a.rho[o] = *rreg;
        return;
    }
    cplx r = a.rho[o];
    cplx ac = (a.stage == 0) ? cmake(0, 0) : a.acc[o];
    const cplx yn = heom_rk_update(a.stage, k, r, ac, a.dt);
    if (a.stage < 3) a.acc[o] = ac; else a.rho[o] = r;
    a.ynext[o] = yn;
    if (a.npeer) {
        const unsigned m = (a.send_all || !a.peer_mask) ? 0xffu : a.peer_mask[ado];
        for (int p = 0; p < a.npeer; ++p)
            if ((m >> p) & 1u) {
                if (a.tag_in) heom_store_tagged(a.tag_out_peer[p], o, yn, a.tag_want + 1);
                else a.peer_next[p][o] = yn;
            }
        if (a.tag_in && a.send_all) heom_store_tagged(a.tag_out_own, o, yn, a.tag_want + 1);
    }
}

// stage-wise kernel: grid ceil(items/apc); block >= apc*nn threads; dynamic smem (1 + 3*apc)*nn cplx
__global__ void __launch_bounds__(1024)
heom_stage_kernel(HeomStageArgs a) {
    extern __shared__ double2 smem[];
    for (int l = threadIdx.x; l < a.d.nn; l += blockDim.x) smem[l] = a.d.H[l];
    heom_stage_tile(a, (long long)blockIdx.x * a.apc, smem, nullptr, nullptr, nullptr);
}

// persistent variant: the whole RK4 loop in ONE cooperative launch; tiles are distributed
// grid-stride, stage vectors stay in L2, one grid barrier per stage.  Used when the hierarchy
// is too large for one CTA's shared memory but small enough that per-stage launches dominate.
struct HeomPersistArgs {
    HeomStageArgs s;        // rho, acc set; yin/ynext = ping-pong buffers y0 (holds rho on entry), y1
    cplx* y0; cplx* y1;
    int nsteps, E, traj_every;
    const cplx* eT; cplx* obs; cplx* traj;
    unsigned* barrier;      // zero-initialised: [0] arrival counter, [1] release word, [2] error word
    // ADO-sharded run (one process per GPU): peers' stage vectors and flag arrays (CUDA IPC mappings)
    int world, rank;
    unsigned epoch;         // flags are monotonic across launches: stage x of this launch signals epoch + x + 1
    cplx* y0p[7]; cplx* y1p[7];       // peers' y0 / y1 (order: every rank != this one)
    unsigned* flagp[7];               // peers' flag arrays; peer q's slot for this rank is flagp[q][rank]
    unsigned* flags;                  // this rank's flag array [world]: flags[q] counts the CTA arrivals of rank q
    unsigned peer_grid[7];            // CTAs of the persistent kernel on peer slot q (its arrivals per stage)
    // hybrid sharding: local grid barrier + tagged halo inboxes (T[b] local, Tp[b][q] the peers'), tags tag0 + stage count
    int hybrid;
    ulonglong2* T[2];
    ulonglong2* Tp[2][7];
    unsigned long long tag0;
    const unsigned char* peer_mask;   // [nhe] which peer slots need each owned ADO (null: all); the last stage of the
                                      // run always goes to every peer so that each rank ends with the full state
};

// grid-wide barrier on a monotonic counter (zeroed by the host before the launch): one release
// reduction per CTA, polling with relaxed L2 loads, ONE acquire fence (which invalidates L1) at
// the end.  cooperative_groups' grid.sync() polls through an L1-invalidating load (CCTL.IVALL per
// poll), measured at ~6 us per barrier with 444 CTAs; the kernel is still launched cooperatively so
// that all CTAs are co-resident.
__device__ __forceinline__ void heom_grid_barrier(unsigned* ctr, unsigned target) {
    // (a variant in which the CTA that completes the count -- atom.add with return -- publishes a release word on its
    //  own L2 line and the others spin on that word was measured 6 % slower: 8.78e7 vs 9.38e7 ADO-steps/s on config 4)
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

// Barrier across every CTA of every rank of an ADO-sharded run, one hop and no funnel: after its stores (local and
// peer) a CTA (1) arrives on the local counter (gpu-scope release: covers its local stores), (2) makes its peer stores
// visible system-wide and counts itself on EVERY peer with one remote reduction into that peer's flag word for this
// rank, (3) waits until the local counter and the flag word of every peer have reached the value that says "all CTAs
// of that rank have finished this stage" (flags are monotonic: peer_grid[q] arrivals per stage), (4) one system-scope
// acquire fence.  Remote signalling therefore overlaps the local collection, nobody polls over NVLink (every flag a
// rank polls lives in its own memory), and no stage is relayed through one CTA.  WAR on the ping-pong buffers is
// covered by the same counts: a peer's arrival for stage s implies it has finished READING the buffer that stage
// s + 1 overwrites.  Spins are bounded (~ 5 s) so that a missing peer produces an error instead of a hung GPU.
__device__ __forceinline__ void heom_world_barrier(const HeomPersistArgs& p, unsigned target, unsigned xcount) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        const long long limit = 10000000000LL;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.barrier) : "memory");
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        for (int q = 0; q < p.world - 1; ++q)
            asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p.flagp[q] + p.rank) : "memory");
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.barrier) : "memory");
        } while (v < target && clock64() - t0 < limit);
        int slot = 0;
        for (int q = 0; q < p.world; ++q) {
            if (q == p.rank) continue;
            const unsigned want = p.peer_grid[slot++] * xcount;
            do {
                asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.flags + q) : "memory");
            } while ((int)(v - want) < 0 && clock64() - t0 < limit);
            if ((int)(v - want) < 0) atomicExch(p.barrier + 2, 1u);
        }
        asm volatile("fence.acq_rel.sys;" ::: "memory");
    }
    __syncthreads();
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
heom_persist_kernel(HeomPersistArgs p) {
    extern __shared__ double2 smem[];
    HeomStageArgs a = p.s;
    unsigned bar_target = 0;
    const HeomDev& d = a.d;
    const int nn = d.nn;
    for (int l = threadIdx.x; l < nn; l += blockDim.x) smem[l] = d.H[l];
    const long long nown = a.row_hi - a.row_lo;
    const long long nitems = nown * a.B;
    const long long ntiles = (nitems + a.apc - 1) / a.apc;
    // one tile per CTA for the whole run: index tables of the tile in shared memory, rho and the RK4
    // accumulator of the thread's element in registers (only the stage vectors travel through L2)
    const bool fixed = ntiles <= (long long)gridDim.x;
    int* tabs = reinterpret_cast<int*>(smem + (size_t)(1 + 3 * a.apc) * nn);
    cplx rreg = cmake(0, 0), areg = cmake(0, 0);
    if (fixed) {
        const int nm = d.nmodes;
        for (int l = threadIdx.x; l < a.apc * nm; l += blockDim.x) {
            const int g = l / nm, k = l - g * nm;
            const long long item = (long long)blockIdx.x * a.apc + g;
            if (item < nitems) {
                const long long ado = a.row_lo + item % nown;
                tabs[(3 * g) * nm + k] = d.states[ado * nm + k];
                tabs[(3 * g + 1) * nm + k] = d.dn[ado * nm + k];
                tabs[(3 * g + 2) * nm + k] = d.up[ado * nm + k];
            }
        }
        const int g = threadIdx.x / nn;
        const long long item = (long long)blockIdx.x * a.apc + g;
        if (g < a.apc && item < nitems) {
            const long long b = item / nown;
            rreg = a.rho[((size_t)b * d.nhe + a.row_lo + (item - b * nown)) * nn + (threadIdx.x - g * nn)];
        }
    }
    __syncthreads();
    for (int step = 0; step < p.nsteps; ++step) {
        for (int stage = 0; stage < 4; ++stage) {
            a.stage = stage;
            a.yin = (stage & 1) ? p.y1 : p.y0;
            a.ynext = (stage & 1) ? p.y0 : p.y1;
            a.npeer = p.world - 1;
            a.peer_mask = p.peer_mask;
            a.send_all = (step == p.nsteps - 1 && stage == 3) ? 1 : 0;
            for (int q = 0; q < p.world - 1; ++q) a.peer_next[q] = (stage & 1) ? p.y0p[q] : p.y1p[q];
            if (p.hybrid) {
                a.tag_in = p.T[stage & 1];
                a.tag_out_own = p.T[(stage + 1) & 1];
                for (int q = 0; q < p.world - 1; ++q) a.tag_out_peer[q] = p.Tp[(stage + 1) & 1][q];
                a.tag_want = p.tag0 + 4ull * step + stage;
                a.tag_err = p.barrier + 2;
            }
            if (fixed) {
                if (blockIdx.x < ntiles) heom_stage_tile(a, (long long)blockIdx.x * a.apc, smem, tabs, &rreg, &areg);
            } else {
                for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                    heom_stage_tile(a, t * a.apc, smem, nullptr, nullptr, nullptr);
                    __syncthreads();
                }
            }
            bar_target += gridDim.x;
            if (p.world > 1 && !p.hybrid) heom_world_barrier(p, bar_target, p.epoch + 4u * step + stage + 1u);
            else heom_grid_barrier(p.barrier, bar_target);      // hybrid: the ranks are coupled through the tagged halo only
        }
        // y0 == rho_{n+1}; tier-0 observables / trajectory: one warp per (hierarchy, observable)
        const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        const int nwarps = (gridDim.x * blockDim.x) >> 5;
        if (p.obs)
            for (int w = warp; w < a.B * p.E; w += nwarps) {
                const int b = w / p.E, e = w - b * p.E;
                const cplx* r = a.rho + (size_t)b * d.nhe * nn;
                cplx v = cmake(0, 0);
                for (int l = lane; l < nn; l += 32) cfma(v, __ldg(p.eT + (size_t)e * nn + l), r[l]);
                for (int off = 16; off > 0; off >>= 1) {
                    v.x += __shfl_down_sync(0xffffffffu, v.x, off);
                    v.y += __shfl_down_sync(0xffffffffu, v.y, off);
                }
                if (lane == 0) p.obs[((size_t)step * a.B + b) * p.E + e] = v;
            }
        if (p.traj && ((step + 1) % p.traj_every) == 0)
            for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < (long long)a.B * nn;
                 l += (long long)gridDim.x * blockDim.x) {
                const long long b = l / nn;
                p.traj[((size_t)(step / p.traj_every) * a.B + b) * nn + (l - b * nn)] =
                    a.rho[(size_t)b * d.nhe * nn + (l - b * nn)];
            }
    }
}

// The same two barriers split into ARRIVE (after a CTA's stores) and WAIT (before it reads other CTAs' values), so that
// work which needs only the CTA's own data -- the -i[H, .] term of the NEXT stage -- runs while the arrivals propagate.
__device__ __forceinline__ void heom_barrier_arrive(const HeomPersistArgs& p) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.barrier) : "memory");
        if (p.world > 1) {
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            for (int q = 0; q < p.world - 1; ++q)
                asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p.flagp[q] + p.rank) : "memory");
        }
    }
}
__device__ __forceinline__ void heom_barrier_wait(const HeomPersistArgs& p, unsigned target, unsigned xcount) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        const long long limit = 10000000000LL;
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.barrier) : "memory");
        } while (v < target && (p.world == 1 || clock64() - t0 < limit));
        if (p.world > 1) {
            int slot = 0;
            for (int q = 0; q < p.world; ++q) {
                if (q == p.rank) continue;
                const unsigned want = p.peer_grid[slot++] * xcount;
                do {
                    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.flags + q) : "memory");
                } while ((int)(v - want) < 0 && clock64() - t0 < limit);
                if ((int)(v - want) < 0) atomicExch(p.barrier + 2, 1u);
            }
            asm volatile("fence.acq_rel.sys;" ::: "memory");
        } else {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
    }
    __syncthreads();
}

// Persistent kernel, diagonal Q, ONE element per thread for the whole run (apc ADOs per CTA,
// every item has its own CTA slot): the element's <= NE neighbour offsets and coefficients are
// computed once and kept in shared memory (structure-of-arrays, conflict-free), rho / the RK4
// accumulator / the damping rate in registers.  Per stage a thread issues its neighbour loads
// (independent L2 reads), exchanges its own stage value through shared memory for -i[H, .],
// and writes one element of the next stage vector; then the grid barrier.
#define HEOM_PC_NE 8
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
heom_persist_cached_kernel(HeomPersistArgs p) {
    extern __shared__ double2 smem[];
    const HeomStageArgs& a = p.s;
    const HeomDev& d = a.d;
    const int nn = d.nn, n = d.n, T = blockDim.x;
    cplx* Hs = smem;                                   // [nn]
    cplx* ys = Hs + nn;                                // [apc*nn]
    cplx* ecf = ys + (size_t)a.apc * nn;               // [NE][T]
    int* eoff = reinterpret_cast<int*>(ecf + (size_t)HEOM_PC_NE * T);   // [NE][T]
    for (int l = threadIdx.x; l < nn; l += T) Hs[l] = d.H[l];
    const long long nown = a.row_hi - a.row_lo;
    const long long nitems = nown * a.B;
    const int g = threadIdx.x / nn;
    const int idx = threadIdx.x - g * nn;
    const int i = idx / n, j = idx - i * n;
    const long long item = (long long)blockIdx.x * a.apc + g;
    const bool act = g < a.apc && item < nitems;
    const int b = act ? (int)(item / nown) : 0;
    const long long ado = act ? a.row_lo + (item - (long long)b * nown) : a.row_lo;
    const int par = d.npar > 1 ? b : 0;
    const size_t hb = (size_t)b * d.nhe * nn;          // hierarchy base
    const size_t own = hb + (size_t)ado * nn + idx;
    cplx rreg = act ? a.rho[own] : cmake(0, 0), areg = cmake(0, 0);
    const double damp = act ? heom_damp(d, par, ado) : 0.0;
    const unsigned pmask = (p.peer_mask && act) ? p.peer_mask[ado] : 0xffu;
    {
        int ne = 0;
        if (act) {
            const int* st = d.states + ado * d.nmodes;
            const int* dn = d.dn + ado * d.nmodes;
            const int* up = d.up + ado * d.nmodes;
            for (int t = d.em_start[idx]; t < d.em_start[idx + 1]; ++t) {
                const int m = d.em_mode[t];
                const double2 v = d.em_v[t];
                const int id = dn[m], iu = up[m];
                if (id >= 0 && ne < HEOM_PC_NE) {
                    eoff[ne * T + threadIdx.x] = (int)(hb + (size_t)id * nn + idx);
                    ecf[ne * T + threadIdx.x] = heom_dn_coef(d, par, m, (double)st[m], v);
                    ++ne;
                }
                if (iu >= 0 && v.x != v.y && ne < HEOM_PC_NE) {
                    eoff[ne * T + threadIdx.x] = (int)(hb + (size_t)iu * nn + idx);
                    ecf[ne * T + threadIdx.x] = cscale(v.x - v.y, d.pref_up);
                    ++ne;
                }
            }
        }
        for (; ne < HEOM_PC_NE; ++ne) {
            eoff[ne * T + threadIdx.x] = (int)own;
            ecf[ne * T + threadIdx.x] = cmake(0, 0);
        }
    }
    __syncthreads();
    unsigned bar_target = 0;
    cplx ycur = rreg;                      // the element's value in the current stage vector (= its own last output)
    // tier-0 observables / trajectory of a finished step (a.rho of every CTA is visible once that step's last barrier
    // has been waited for)
    auto outputs = [&](int step) {
        const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        const int nwarps = (gridDim.x * blockDim.x) >> 5;
        if (p.obs)
            for (int w = warp; w < a.B * p.E; w += nwarps) {
                const int bb = w / p.E, e = w - bb * p.E;
                const cplx* r = a.rho + (size_t)bb * d.nhe * nn;
                cplx v = cmake(0, 0);
                for (int l = lane; l < nn; l += 32) cfma(v, __ldg(p.eT + (size_t)e * nn + l), r[l]);
                for (int off = 16; off > 0; off >>= 1) {
                    v.x += __shfl_down_sync(0xffffffffu, v.x, off);
                    v.y += __shfl_down_sync(0xffffffffu, v.y, off);
                }
                if (lane == 0) p.obs[((size_t)step * a.B + bb) * p.E + e] = v;
            }
        if (p.traj && ((step + 1) % p.traj_every) == 0)
            for (long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x; l < (long long)a.B * nn;
                 l += (long long)gridDim.x * blockDim.x) {
                const long long bb = l / nn;
                p.traj[((size_t)(step / p.traj_every) * a.B + bb) * nn + (l - bb * nn)] =
                    a.rho[(size_t)bb * d.nhe * nn + (l - bb * nn)];
            }
    };
    for (int step = 0; step < p.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const cplx* yin = (stage & 1) ? p.y1 : p.y0;
            cplx* yout = (stage & 1) ? p.y0 : p.y1;
            // ---- the part of the right-hand side that needs only this CTA's own ADOs: -i[H, Y_a] and the damping.  It runs
            // BEFORE the wait on the previous stage's barrier, i.e. while the other CTAs' arrivals propagate.
            if (g < a.apc) ys[threadIdx.x] = ycur;
            __syncthreads();
            cplx k = heom_sys_any(Hs, n, ys + (size_t)g * nn, i, j);
            k.x = fma(-damp, ycur.x, k.x);
            k.y = fma(-damp, ycur.y, k.y);
            if (step > 0 || stage > 0) {
                heom_barrier_wait(p, bar_target, p.epoch + 4u * step + stage);
                if (stage == 0) outputs(step - 1);
            }
            // ---- neighbour ADOs (other CTAs, other GPUs)
            cplx nb[HEOM_PC_NE];
#pragma unroll
            for (int e = 0; e < HEOM_PC_NE; ++e) nb[e] = yin[eoff[e * T + threadIdx.x]];
#pragma unroll
            for (int e = 0; e < HEOM_PC_NE; ++e) cfma(k, ecf[e * T + threadIdx.x], nb[e]);
            const cplx yn = heom_rk_update(stage, k, rreg, areg, a.dt);
            if (act) {
                yout[own] = yn;
                const unsigned m = (step == p.nsteps - 1 && stage == 3) ? 0xffu : pmask;
                for (int q = 0; q < p.world - 1; ++q)
                    if ((m >> q) & 1u) ((stage & 1) ? p.y0p[q] : p.y1p[q])[own] = yn;
                if (stage == 3) a.rho[own] = rreg;
            }
            ycur = yn;
            bar_target += gridDim.x;
            heom_barrier_arrive(p);
        }
    }
    if (p.nsteps > 0) {
        heom_barrier_wait(p, bar_target, p.epoch + 4u * p.nsteps);
        outputs(p.nsteps - 1);
    }
}

#include "heom_flow.cuh"

// ---- packed-neighbour stage kernel (diagonal coupling operators) -----------------------------------------------------
// heom_stage_kernel walks the index tables for every element in every stage: ~850 thread instructions per element-stage
// of the FMO hierarchy against ~100 FP64 operations (profiles/r01_heom_fmo_batch64_stage_v2.txt: issue 49 %, FP64 19 %,
// l1tex 72 %).  Here the walk is done ONCE per plan: every element gets its <= 8 neighbour entries as packed 32-bit
// words (entry index within a hierarchy | n_k << 26, as in heom_flow_cached_kernel), 32 bytes per element and shared by
// all hierarchies of a batch, and the finished coefficients for n_k = 0 .. max come from a small table
// [npar][nn][8][NK] (31 KB for the FMO hierarchy: L1 resident).  A stage of one element is then: 2 packed-word loads,
// 8 independent neighbour loads, -i[H, .] through shared memory, 8 table look-ups + complex FMAs, the RK4 update.  A CTA
// keeps the packed words in registers and walks HB hierarchies of the batch with them.  Same arithmetic, same order as
// heom_stage_tile / heom_persist_cached_kernel (bit-identical results).
static_assert(HEOM_PK_NE == HEOM_FLOW_NE && HEOM_PK_IDXBITS == HEOM_FLOW_IDXBITS, "one packed-word format");
struct HeomFastArgs {
    HeomDev d;
    int B, stage, HB, NK, apc;       // stage -1: plain right-hand side into ynext
    long long row_lo, row_hi;
    cplx* rho; const cplx* yin; cplx* ynext; cplx* acc;
    double dt;
    const uint4* pk;                 // [nhe * nn][2]
    const cplx* ctab;                // [npar][nn][HEOM_FLOW_NE][NK]
};

__global__ void __launch_bounds__(256)
heom_pack_kernel(HeomDev d, uint4* __restrict__ pk) {
    const long long total = d.nhe * d.nn;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ado = e / d.nn;
        const int idx = (int)(e - ado * d.nn);
        unsigned w[HEOM_FLOW_NE];
#pragma unroll
        for (int s = 0; s < HEOM_FLOW_NE; ++s) w[s] = (unsigned)e;           // own entry, n_k slot 0: coefficient 0
        const int* st = d.states + ado * d.nmodes;
        const int* dn = d.dn + ado * d.nmodes;
        const int* up = d.up + ado * d.nmodes;
        int q = 0;
        for (int t = d.em_start[idx]; t < d.em_start[idx + 1] && q < HEOM_FLOW_NE / 2; ++t, ++q) {
            const int m = d.em_mode[t];
            const int id = dn[m], iu = up[m];
            if (id >= 0) w[2 * q] = ((unsigned)id * (unsigned)d.nn + (unsigned)idx) | ((unsigned)st[m] << HEOM_FLOW_IDXBITS);
            if (iu >= 0) w[2 * q + 1] = ((unsigned)iu * (unsigned)d.nn + (unsigned)idx) | (1u << HEOM_FLOW_IDXBITS);
        }
        pk[2 * e] = make_uint4(w[0], w[1], w[2], w[3]);
        pk[2 * e + 1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}
// ctab[par][slot][nk][idx] (matrix element fastest: the lanes of a warp read neighbouring entries; with the element
// index slowest every lane hit its own 128-byte line and the look-ups alone saturated l1tex --
// profiles/r02_heom_fmo_batch64_stage_fast_v1.txt: l1tex 94.7 %).  Down slots 2q hold n_k (q_i pref_dn c_k - q_j pref_dn
// conj c_k) formed exactly as heom_dn_coef forms it, up slots 2q + 1 hold (q_i - q_j) pref_up at "n_k" = 1; the rest 0.
__global__ void __launch_bounds__(256)
heom_ctab_kernel(HeomDev d, int NK, cplx* __restrict__ ctab) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= d.npar * d.nn) return;
    const int par = l / d.nn, idx = l - par * d.nn;
    for (int s = 0; s < HEOM_PK_NE; ++s)
        for (int nk = 0; nk < NK; ++nk) ctab[heom_ctab_index(par, s, nk, idx, NK, d.nn)] = cmake(0, 0);
    int q = 0;
    for (int t = d.em_start[idx]; t < d.em_start[idx + 1] && q < HEOM_PK_NE / 2; ++t, ++q) {
        const int m = d.em_mode[t];
        const double2 v = d.em_v[t];
        for (int nk = 1; nk < NK; ++nk) ctab[heom_ctab_index(par, 2 * q, nk, idx, NK, d.nn)] = heom_dn_coef(d, par, m, (double)nk, v);
        if (NK > 1) ctab[heom_ctab_index(par, 2 * q + 1, 1, idx, NK, d.nn)] = cscale(v.x - v.y, d.pref_up);
    }
}

// -i [H, Y]_{ij} for a REAL Hamiltonian (Hr: n x n doubles): half the shared-memory bytes and half the FP64 work of the
// complex form on the H side; same values (the dropped terms are products with Im H = 0)
template <int NN_>
__device__ __forceinline__ cplx heom_sys_real_t(const double* Hr, const cplx* ya, int i, int j) {
    cplx s = cmake(0, 0);
    const double* hrow = Hr + i * NN_;
    const cplx* yrow = ya + i * NN_;
#pragma unroll
    for (int m = 0; m < NN_; ++m) {
        const double h1 = hrow[m];
        const cplx y1 = ya[m * NN_ + j];
        s.x = fma(h1, y1.x, s.x);
        s.y = fma(h1, y1.y, s.y);
        const cplx a = yrow[m];
        const double h2 = Hr[m * NN_ + j];
        s.x = fma(-a.x, h2, s.x);
        s.y = fma(-a.y, h2, s.y);
    }
    return cmake(s.y, -s.x);          // -i * s
}

// HR: the system Hamiltonian is real (needs NN_ != 0).  CB (shared bath parameters only): the coefficient of a slot is
// n_k x base[slot][element] with the 8 x nn base coefficients (the n_k = 1 column of the table) in shared memory --
// conflict-free reads, one conversion and two multiplications per neighbour -- instead of a look-up of the finished
// coefficient in global memory / L1.  n_k x base is how heom_dn_coef forms the value, so the results are the same bits.
template <int NN_, bool HR, bool CB>
__global__ void __launch_bounds__(256, 3)
heom_stage_fast_kernel(HeomFastArgs a) {
    extern __shared__ double2 smem[];
    const HeomDev& d = a.d;
    const int n = NN_ ? NN_ : d.n, nn = n * n, T = blockDim.x;
    cplx* Hs = smem;                 // [nn] (HR: the first nn doubles hold Re H)
    double* Hr = reinterpret_cast<double*>(smem);
    cplx* ys0 = Hs + nn;             // [2][apc * nn]: double-buffered over the hierarchies of the CTA (one barrier each)
    cplx* base = ys0 + (size_t)2 * a.apc * nn;     // [NE][nn] (CB)
    for (int l = threadIdx.x; l < nn; l += T) {
        if (HR) Hr[l] = d.H[l].x; else Hs[l] = d.H[l];
    }
    if (CB)
        for (int l = threadIdx.x; l < HEOM_FLOW_NE * nn; l += T) {
            const int s = l / nn;
            base[l] = a.NK > 1 ? __ldg(a.ctab + heom_ctab_index(0, s, 1, l - s * nn, a.NK, nn)) : cmake(0, 0);
        }
    const int g = threadIdx.x / nn, idx = threadIdx.x - g * nn;
    const int i = idx / n, j = idx - i * n;
    const long long ado = a.row_lo + (long long)blockIdx.x * a.apc + g;
    const bool act = g < a.apc && ado < a.row_hi;
    const size_t e = (size_t)(act ? ado : a.row_lo) * nn + (act ? idx : 0);      // element index inside one hierarchy
    unsigned w[HEOM_FLOW_NE];
    {
        const uint4 p0 = __ldg(a.pk + 2 * e), p1 = __ldg(a.pk + 2 * e + 1);
        w[0] = p0.x; w[1] = p0.y; w[2] = p0.z; w[3] = p0.w; w[4] = p1.x; w[5] = p1.y; w[6] = p1.z; w[7] = p1.w;
    }
    const size_t hstride = (size_t)d.nhe * nn;
    const double damp1 = (act && d.damp) ? __ldg(d.damp + ado) : 0.0;      // shared bath parameters: one damping rate per ADO
    const int b0 = blockIdx.y * a.HB, b1 = min(a.B, b0 + a.HB);
    for (int b = b0; b < b1; ++b) {
        const int par = d.npar > 1 ? b : 0;
        const cplx* y = a.yin + (size_t)b * hstride;
        const size_t o = (size_t)b * hstride + e;
        // own value and the 8 neighbour values: independent loads, all in flight together
        const cplx yv = y[e];
        cplx nb[HEOM_FLOW_NE];
#pragma unroll
        for (int s = 0; s < HEOM_FLOW_NE; ++s) nb[s] = y[w[s] & ((1u << HEOM_FLOW_IDXBITS) - 1u)];
        // (no second barrier: the buffer written now was last read two hierarchies ago, and every thread has passed
        //  the barrier in between only after finishing those reads)
        cplx* ys = ys0 + (size_t)((b - b0) & 1) * a.apc * nn;
        if (g < a.apc) ys[threadIdx.x] = yv;
        __syncthreads();
        if (act) {
            cplx k = HR ? heom_sys_real_t<(NN_ ? NN_ : 2)>(Hr, ys + (size_t)g * nn, i, j)
                   : NN_ ? heom_sys_t<(NN_ ? NN_ : 2)>(Hs, ys + (size_t)g * nn, i, j) : heom_sys(Hs, n, ys + (size_t)g * nn, i, j);
            const double damp = d.damp ? damp1 : heom_damp(d, par, ado);
            k.x = fma(-damp, yv.x, k.x);
            k.y = fma(-damp, yv.y, k.y);
#pragma unroll
            for (int s = 0; s < HEOM_FLOW_NE; ++s) {
                if (CB) cfma(k, cscale((double)(w[s] >> HEOM_FLOW_IDXBITS), base[s * nn + idx]), nb[s]);
                else cfma(k, __ldg(a.ctab + heom_ctab_index(par, s, (int)(w[s] >> HEOM_FLOW_IDXBITS), idx, a.NK, nn)), nb[s]);
            }
            if (a.stage < 0) {
                a.ynext[o] = k;
            } else {
                cplx r = a.rho[o];
                cplx ac = (a.stage == 0) ? cmake(0, 0) : a.acc[o];
                const cplx yn = heom_rk_update(a.stage, k, r, ac, a.dt);
                if (a.stage < 3) a.acc[o] = ac; else a.rho[o] = r;
                a.ynext[o] = yn;
            }
        }
    }
}

// on-chip kernel: one CTA per hierarchy, all nsteps fused.  smem: Hs[nn], y0,y1 [nhe*nn] (+ L,R if dense Q)
struct HeomChipArgs {
    HeomDev d;
    int B, nsteps, traj_every, E, T;
    cplx* ado; const cplx* eT; cplx* obs; cplx* traj;
    double dt;
};

// tier-0 observables / trajectory from a shared-memory hierarchy (ADO 0 = reduced density matrix)
__device__ __forceinline__ void heom_chip_outputs(const HeomChipArgs& a, const cplx* y0, int b, int step) {
    const int nn = a.d.nn;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (a.obs)
        for (int e = warp; e < a.E; e += nw) {
            cplx v = cmake(0, 0);
            for (int l = lane; l < nn; l += 32) cfma(v, __ldg(a.eT + (size_t)e * nn + l), y0[l]);
            for (int off = 16; off > 0; off >>= 1) {
                v.x += __shfl_down_sync(0xffffffffu, v.x, off);
                v.y += __shfl_down_sync(0xffffffffu, v.y, off);
            }
            if (lane == 0) a.obs[((size_t)step * a.B + b) * a.E + e] = v;
        }
    if (a.traj && ((step + 1) % a.traj_every) == 0) {
        cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * nn;
        for (int l = threadIdx.x; l < nn; l += blockDim.x) dst[l] = y0[l];
    }
}

// diagonal Q, one element per thread, at most NE/2 modes per element: neighbour offsets and
// coefficients are computed ONCE and stay in registers for the whole launch
// NS > 0: system dimension known at compile time (row i / column j of H in registers, -i[H, .] unrolled)
template <int NE, int NS>
__global__ void __launch_bounds__(1024, 1)
heom_onchip_cached(HeomChipArgs a) {
    extern __shared__ double2 smem[];
    const HeomDev& d = a.d;
    const int nn = d.nn, n = d.n;
    const int total = (int)d.nhe * nn;
    const int b = blockIdx.x;
    const int par = d.npar > 1 ? b : 0;
    cplx* Hs = smem;
    cplx* y0 = Hs + nn;
    cplx* y1 = y0 + total;
    cplx* gado = a.ado + (size_t)b * total;
    const int l = threadIdx.x;
    const bool ok = l < total;
    const int ado = ok ? l / nn : 0, idx = ok ? l - ado * nn : 0;
    const int i = idx / n, j = idx - i * n;
    for (int t = threadIdx.x; t < nn; t += blockDim.x) Hs[t] = d.H[t];
    cplx rho = ok ? gado[l] : cmake(0, 0), acc = cmake(0, 0);
    if (ok) y0[l] = rho;
    const double damp = ok ? heom_damp(d, par, ado) : 0.0;
    int eoff[NE];
    cplx ecf[NE];
    {
        const int t0 = __ldg(d.em_start + idx), t1 = __ldg(d.em_start + idx + 1);
#pragma unroll
        for (int s = 0; s < NE / 2; ++s) {
            eoff[2 * s] = eoff[2 * s + 1] = 0;
            ecf[2 * s] = ecf[2 * s + 1] = cmake(0, 0);
            if (ok && t0 + s < t1) {
                const int m = __ldg(d.em_mode + t0 + s);
                const double2 v = __ldg(d.em_v + t0 + s);
                const int id = __ldg(d.dn + (size_t)ado * d.nmodes + m);
                if (id >= 0) {
                    eoff[2 * s] = id * nn + idx;
                    ecf[2 * s] = heom_dn_coef(d, par, m, (double)__ldg(d.states + (size_t)ado * d.nmodes + m), v);
                }
                const int iu = __ldg(d.up + (size_t)ado * d.nmodes + m);
                if (iu >= 0) {
                    eoff[2 * s + 1] = iu * nn + idx;
                    ecf[2 * s + 1] = cscale(v.x - v.y, d.pref_up);
                }
            }
        }
    }
    __syncthreads();
    constexpr int NSS = NS > 0 ? NS : 1;
    cplx hrow[NSS], hcol[NSS];
#pragma unroll
    for (int m = 0; m < NSS; ++m) {
        hrow[m] = NS > 0 ? Hs[i * NS + m] : cmake(0, 0);
        hcol[m] = NS > 0 ? Hs[m * NS + j] : cmake(0, 0);
    }
    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const cplx* yin = (stage & 1) ? y1 : y0;
            cplx* yout = (stage & 1) ? y0 : y1;
            const cplx yv = yin[ok ? l : 0];
            cplx k;
            if (NS > 0) {
                const cplx* ya = yin + (size_t)ado * nn;
                cplx sacc = cmake(0, 0);
#pragma unroll
                for (int m = 0; m < NSS; ++m) {
                    cfma(sacc, hrow[m], ya[m * NS + j]);
                    const cplx t = cmul(ya[i * NS + m], hcol[m]);
                    sacc.x -= t.x; sacc.y -= t.y;
                }
                k = cmake(sacc.y, -sacc.x);
            } else {
                k = heom_sys(Hs, n, yin + (size_t)ado * nn, i, j);
            }
            k.x = fma(-damp, yv.x, k.x);
            k.y = fma(-damp, yv.y, k.y);
#pragma unroll
            for (int e = 0; e < NE; ++e) cfma(k, ecf[e], yin[eoff[e]]);
            const cplx yn = heom_rk_update(stage, k, rho, acc, a.dt);
            if (ok) yout[l] = yn;
            __syncthreads();
        }
        heom_chip_outputs(a, y0, b, step);
    }
    if (ok) gado[l] = rho;
}

// Small systems (n = NS known at compile time, diagonal Q, at most NM modes): ONE THREAD PER ADO,
// several hierarchies per CTA.  The thread keeps its ADO (NS x NS), rho, the RK4 accumulator, the
// neighbour offsets and the per-mode coefficients in registers; -i[H, .] needs no shared memory
// at all, and the stage vector is stored structure-of-arrays ([element][ADO]) so that the gather of
// a neighbour ADO by consecutive threads is conflict-free.  Shared-memory traffic per element-stage
// drops from ~10 to (1 + 2 NM + 1) accesses, index arithmetic from per-element to per-ADO.
template <int NS, int NM>
__global__ void __launch_bounds__(576, 1)
heom_onchip_ado_kernel(HeomChipArgs a, int HP) {
    extern __shared__ double2 smem[];
    constexpr int NN = NS * NS;
    const HeomDev& d = a.d;
    const int nhe = (int)d.nhe;
    const int hl = threadIdx.x / nhe;                  // hierarchy slot in this CTA
    const int ado = threadIdx.x - hl * nhe;
    const int b = blockIdx.x * HP + hl;
    const bool ok = hl < HP && b < a.B;
    const int par = d.npar > 1 ? (ok ? b : 0) : 0;
    const int span = HP * nhe;                         // ADOs per buffer row
    cplx* y0 = smem;                                   // [NN][span]
    cplx* y1 = y0 + (size_t)NN * span;
    const int me = hl < HP ? hl * nhe + ado : 0;       // this ADO's column (idle threads: any valid one)
    cplx* gado = a.ado + ((size_t)(ok ? b : 0) * nhe + ado) * NN;
    cplx H[NN];
#pragma unroll
    for (int e = 0; e < NN; ++e) H[e] = __ldg(d.H + e);
    cplx rho[NN], acc[NN];
#pragma unroll
    for (int e = 0; e < NN; ++e) {
        rho[e] = ok ? gado[e] : cmake(0, 0);
        acc[e] = cmake(0, 0);
        if (hl < HP) y0[e * span + me] = rho[e];
    }
    const double damp = ok ? heom_damp(d, par, ado) : 0.0;
    int dno[NM], upo[NM];
    cplx cl[NM], cr[NM];
    double qd[NM][NS];
    {
        // diagonal of the coupling operator of each mode, recovered from the per-element lists
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            dno[m] = upo[m] = me;
            cl[m] = cr[m] = cmake(0, 0);
#pragma unroll
            for (int i = 0; i < NS; ++i) qd[m][i] = 0.0;
        }
        for (int i = 0; i < NS; ++i) {
            const int idx = i * NS + i;
            for (int t = d.em_start[idx]; t < d.em_start[idx + 1]; ++t) {
                const int m = d.em_mode[t];
#pragma unroll
                for (int mm = 0; mm < NM; ++mm)
                    if (mm == m) {
#pragma unroll
                        for (int ii = 0; ii < NS; ++ii)
                            if (ii == i) qd[mm][ii] = d.em_v[t].x;
                    }
            }
        }
        if (ok) {
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                if (m < d.nmodes) {
                    const int id = d.dn[(size_t)ado * d.nmodes + m], iu = d.up[(size_t)ado * d.nmodes + m];
                    const double nk = (double)d.states[(size_t)ado * d.nmodes + m];
                    if (id >= 0) {
                        dno[m] = hl * nhe + id;
                        cl[m] = cscale(nk, d.cdn[(size_t)par * d.nmodes + m]);
                        cr[m] = cscale(nk, d.cdnR[(size_t)par * d.nmodes + m]);
                    }
                    if (iu >= 0) upo[m] = hl * nhe + iu; else upo[m] = -1 - me;     // flag: no up neighbour
                }
            }
        }
    }
    __syncthreads();
    const cplx pu = d.pref_up;
    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const cplx* yin = (stage & 1) ? y1 : y0;
            cplx* yout = (stage & 1) ? y0 : y1;
            cplx y[NN], k[NN];
#pragma unroll
            for (int e = 0; e < NN; ++e) y[e] = yin[e * span + me];
            // -i [H, y] - damp y, all in registers
#pragma unroll
            for (int i = 0; i < NS; ++i)
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    cplx sacc = cmake(0, 0);
#pragma unroll
                    for (int m = 0; m < NS; ++m) {
                        cfma(sacc, H[i * NS + m], y[m * NS + j]);
                        const cplx t = cmul(y[i * NS + m], H[m * NS + j]);
                        sacc.x -= t.x; sacc.y -= t.y;
                    }
                    k[i * NS + j] = cmake(fma(-damp, y[i * NS + j].x, sacc.y), fma(-damp, y[i * NS + j].y, -sacc.x));
                }
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const bool has_up = upo[m] >= 0;
                const int uo = has_up ? upo[m] : me;
#pragma unroll
                for (int i = 0; i < NS; ++i)
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        const int e = i * NS + j;
                        const cplx yd = yin[e * span + dno[m]];
                        const cplx yu = yin[e * span + uo];
                        const double qi = qd[m][i], qj = qd[m][j];
                        const cplx cdn = cmake(qi * cl[m].x - qj * cr[m].x, qi * cl[m].y - qj * cr[m].y);
                        cfma(k[e], cdn, yd);
                        if (has_up) cfma(k[e], cscale(qi - qj, pu), yu);
                    }
            }
#pragma unroll
            for (int e = 0; e < NN; ++e) {
                const cplx yn = heom_rk_update(stage, k[e], rho[e], acc[e], a.dt);
                if (hl < HP) yout[e * span + me] = yn;
            }
            __syncthreads();
        }
        if (ok && ado == 0) {          // tier 0 = reduced density matrix, in this thread's registers
            if (a.obs)
                for (int eo = 0; eo < a.E; ++eo) {
                    cplx v = cmake(0, 0);
#pragma unroll
                    for (int e = 0; e < NN; ++e) cfma(v, __ldg(a.eT + (size_t)eo * NN + e), rho[e]);
                    a.obs[((size_t)step * a.B + b) * a.E + eo] = v;
                }
            if (a.traj && ((step + 1) % a.traj_every) == 0) {
                cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * NN;
#pragma unroll
                for (int e = 0; e < NN; ++e) dst[e] = rho[e];
            }
        }
    }
    if (ok) {
#pragma unroll
        for (int e = 0; e < NN; ++e) gado[e] = rho[e];
    }
}

template <int EPT>
__global__ void __launch_bounds__(1024, 1)
heom_onchip_kernel(HeomChipArgs a) {
    extern __shared__ double2 smem[];
    const HeomDev& d = a.d;
    const int nn = d.nn, n = d.n, T = a.T;
    const int total = (int)d.nhe * nn;
    const int b = blockIdx.x;
    const int par = d.npar > 1 ? b : 0;
    cplx* Hs = smem;
    cplx* y0 = Hs + nn;
    cplx* y1 = y0 + total;
    cplx* Ls = y1 + total;
    cplx* Rs = Ls + total;
    cplx* gado = a.ado + (size_t)b * total;
    cplx rho[EPT], acc[EPT];
    double damp[EPT];
    bool ok[EPT];
    for (int t = threadIdx.x; t < nn; t += T) Hs[t] = d.H[t];
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
                cplx s = heom_sys(Hs, n, yin + (size_t)ado * nn, i, j);
                const cplx yv = yin[l];
                s.x = fma(-damp[e], yv.x, s.x);
                s.y = fma(-damp[e], yv.y, s.y);
                if (d.diagq) heom_bath_diag(d, par, d.states + (size_t)ado * d.nmodes, d.dn + (size_t)ado * d.nmodes,
                                            d.up + (size_t)ado * d.nmodes, idx, yin, s);
                k[e] = s;
            }
            if (!d.diagq) {
                for (int q = 0; q < d.nq; ++q) {
                    cplx L[EPT], R[EPT];
#pragma unroll
                    for (int e = 0; e < EPT; ++e) {
                        L[e] = cmake(0, 0); R[e] = cmake(0, 0);
                        if (!ok[e]) continue;
                        const int l = threadIdx.x + e * T;
                        const int ado = l / nn, idx = l - ado * nn;
                        heom_combos(d, par, d.states + (size_t)ado * d.nmodes, d.dn + (size_t)ado * d.nmodes,
                                    d.up + (size_t)ado * d.nmodes, q, idx, yin, L[e], R[e]);
                    }
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
                yout[threadIdx.x + e * T] = heom_rk_update(stage, k[e], rho[e], acc[e], a.dt);
            }
            __syncthreads();
        }
        heom_chip_outputs(a, y0, b, step);      // y0 holds the new hierarchy; tier 0 is ADO 0
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
    bool h_real = false;                // Im H == 0 everywhere: the packed-neighbour stage kernel takes the real -i[H, .] form
    cplx pref_up;
    DevBuf dH, dQ, dqstart, dqmodes, dems, demm, demv, ddamp, dcdn, dcdnR, dnu, dstates, ddn, dup;
    int max_modes_per_elem = 0;
    int max_nk = 0;                     // largest occupation number in the index table
    DevBuf s_y, s_acc, dbar;
    DevBuf s_T;                         // dataflow path: two tagged stage vectors (32 B per element each)
    DevBuf d_pk, d_ctab;                // packed-neighbour stage kernel: per-element neighbour words, coefficient table
    int fast_state = 0;                 // 0: not prepared, 1: ready, -1: not applicable
    unsigned long long flow_tag = 1;    // next unused stage tag (monotonic over the launches of the plan)
    long long launches = 0;
    long long smem_optin = 0;
    int sm_count = 148;
    HeomDev dev() const {
        HeomDev d;
        d.n = n; d.nn = n * n; d.nmodes = nmodes; d.nq = nq; d.npar = npar; d.nhe = nhe;
        d.H = dH.as<cplx>(); d.Q = dQ.as<cplx>(); d.diagq = diagq ? 1 : 0;
        d.qstart = dqstart.as<int>(); d.qmodes = dqmodes.as<int>();
        d.em_start = dems.as<int>(); d.em_mode = demm.as<int>(); d.em_v = demv.as<double2>();
        d.damp = (npar == 1) ? ddamp.as<double>() : nullptr;
        d.cdn = dcdn.as<cplx>(); d.cdnR = dcdnR.as<cplx>(); d.nu = dnu.as<double>();
        d.pref_up = pref_up;
        d.states = dstates.as<int>(); d.dn = ddn.as<int>(); d.up = dup.as<int>();
        return d;
    }
};

extern "C" {

static int heom_fast_prepare(limeb200_heom_t p);
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
    p->h_real = true;
    for (int l = 0; l < nn; ++l)
        if (h_H[2 * l + 1] != 0.0) p->h_real = false;
    LB_CUDA(p->dQ.upload(h_Q, (size_t)nq * nn * 16));
    {   // per matrix element: the modes whose (diagonal) coupling operator touches row i or column j
        std::vector<int> ems(nn + 1, 0), emm;
        std::vector<double> emv;
        int mx = 0;
        for (int idx = 0; idx < nn; ++idx) {
            const int i = idx / n, j = idx % n;
            for (int k = 0; k < nmodes; ++k) {
                const double vi = qd[qmap[k] * n + i], vj = qd[qmap[k] * n + j];
                if (vi != 0.0 || vj != 0.0) { emm.push_back(k); emv.push_back(vi); emv.push_back(vj); }
            }
            ems[idx + 1] = (int)emm.size();
            mx = std::max(mx, ems[idx + 1] - ems[idx]);
        }
        p->max_modes_per_elem = mx;
        if (emm.empty()) { emm.push_back(0); emv.push_back(0.0); emv.push_back(0.0); }
        LB_CUDA(p->dems.upload(ems.data(), ems.size() * 4));
        LB_CUDA(p->demm.upload(emm.data(), emm.size() * 4));
        LB_CUDA(p->demv.upload(emv.data(), emv.size() * 8));
    }
    if (npar == 1) {
        std::vector<double> damp((size_t)nhe, 0.0);
        for (long long a = 0; a < nhe; ++a) {
            double sd = 0.0;          // same fma order as the device loop
            for (int k = 0; k < nmodes; ++k) sd = std::fma((double)states[a * nmodes + k], h_nu[k], sd);
            damp[a] = sd;
        }
        LB_CUDA(p->ddamp.upload(damp.data(), damp.size() * 8));
    }
    LB_CUDA(p->dqstart.upload(qstart.data(), qstart.size() * 4));
    LB_CUDA(p->dqmodes.upload(qmodes.data(), qmodes.size() * 4));
    LB_CUDA(p->dcdn.upload(cdn.data(), cdn.size() * 16));
    LB_CUDA(p->dcdnR.upload(cdnR.data(), cdnR.size() * 16));
    LB_CUDA(p->dnu.upload(h_nu, (size_t)npar * nmodes * 8));
    for (long long l = 0; l < nhe * nmodes; ++l) p->max_nk = std::max(p->max_nk, states[l]);
    LB_CUDA(p->dstates.upload(states, (size_t)nhe * nmodes * 4));
    LB_CUDA(p->ddn.upload(dn, (size_t)nhe * nmodes * 4));
    LB_CUDA(p->dup.upload(up, (size_t)nhe * nmodes * 4));
    {
        const int r = heom_fast_prepare(p.get());
        if (r != LB_OK) return r;
    }
    *plan = p.release();
    return LB_OK;
}

int limeb200_heom_destroy(limeb200_heom_t p) {
    if (p) { cudaSetDevice(p->device); delete p; }
    return LB_OK;
}
int limeb200_heom_set_path(limeb200_heom_t p, int path) {
    LB_REQUIRE(p && path >= 0 && path <= 4, "bad arguments");
    p->path_req = path;
    return LB_OK;
}
int limeb200_heom_get_path(limeb200_heom_t p) { return p ? p->path : LB_ERR_ARG; }
long long limeb200_heom_last_launches(limeb200_heom_t p) { return p ? p->launches : -1; }

// b0: first hierarchy of the batch slice [b0, b0 + B) this launch works on (all batch-major pointers and the
// per-hierarchy bath parameters are offset by it)
// How a stage gathers the bath terms.  0: walk the index tables (heom_bath_diag), 1: generic tile code with the packed
// gather, 2: heom_stage_fast_kernel with the finished-coefficient table in global memory, 3: the same with n_k x base
// coefficients from shared memory.  Measured on B200 (profiles/r02_heom_stage_modes_pass3.jsonl, ADO-steps/s):
//   batch of 64 FMO hierarchies (3060 ADOs each):  0: 1.01e8   1: 1.30e8   2: 1.40e8   3: 1.42e8
//   one 38 760-ADO hierarchy:                       0: 8.50e7   1: 1.10e8   2: 9.48e7   3: 8.77e7
// so batches (where a CTA reuses its packed words over 4 hierarchies) take mode 3 and single hierarchies mode 1 (61
// registers, 4 CTAs per SM: the gather of a 30 MB stage vector is latency bound and wants the occupancy).  Per-hierarchy
// bath parameters keep the table walk (their coefficient tables were not measured).  LIMEB200_HEOM_STAGE_MODE overrides.
static int heom_stage_mode(limeb200_heom_t p, int B) {
    if (p->fast_state != 1 || getenv("LIMEB200_HEOM_NO_FAST_STAGE")) return 0;
    const char* e = getenv("LIMEB200_HEOM_STAGE_MODE");
    if (e && *e >= '0' && *e <= '3' && !e[1]) return *e - '0';
    if (p->npar > 1) return 0;
    return B >= 4 ? 3 : 1;
}
// tables of the packed-neighbour stage kernel; called once when the plan is created (never inside a stream capture)
static int heom_fast_prepare(limeb200_heom_t p) {
    const int nn = p->n * p->n;
    const long long total = p->nhe * nn;
    p->fast_state = (p->diagq && 2 * p->max_modes_per_elem <= HEOM_FLOW_NE && total < (1LL << HEOM_FLOW_IDXBITS) &&
                     p->max_nk < 64 && nn <= 256) ? 1 : -1;
    const int NK = p->max_nk + 1;
    if ((size_t)p->npar * nn * HEOM_FLOW_NE * NK * 16 > ((size_t)256 << 20)) p->fast_state = -1;     // per-hierarchy parameter sets: keep the table small
    if (p->fast_state != 1) return LB_OK;
    LB_CUDA(p->d_pk.alloc((size_t)total * 32));
    LB_CUDA(p->d_ctab.alloc((size_t)p->npar * nn * HEOM_FLOW_NE * NK * 16));
    heom_pack_kernel<<<(unsigned)std::min<long long>(ceil_div(total, 256LL), 148 * 16), 256>>>(p->dev(), p->d_pk.as<uint4>());
    heom_ctab_kernel<<<ceil_div(p->npar * nn, 256), 256>>>(p->dev(), NK, p->d_ctab.as<cplx>());
    LB_CUDA(cudaGetLastError());
    LB_CUDA(cudaDeviceSynchronize());
    return LB_OK;
}

static int heom_launch_stage(limeb200_heom_t p, int stage, cplx* rho, const cplx* yin, cplx* ynext, cplx* acc,
                             int B, double dt, cudaStream_t st, int b0 = 0) {
    const long long nown = p->row_hi - p->row_lo;
    if (nown <= 0) return LB_OK;
    const int nn = p->n * p->n;
    HeomStageArgs a;
    memset(&a, 0, sizeof(a));          // (tagged-halo fields must be null here)
    a.npeer = 0; a.peer_mask = nullptr; a.send_all = 1;
    a.d = p->dev();
    if (b0 > 0) {
        const size_t off = (size_t)b0 * p->nhe * nn;
        if (rho) rho += off;
        if (acc) acc += off;
        yin += off; ynext += off;
        if (p->npar > 1) {
            a.d.cdn += (size_t)b0 * p->nmodes; a.d.cdnR += (size_t)b0 * p->nmodes; a.d.nu += (size_t)b0 * p->nmodes;
        }
    }
    a.B = B; a.stage = stage; a.row_lo = p->row_lo; a.row_hi = p->row_hi;
    a.rho = rho; a.yin = yin; a.ynext = ynext; a.acc = acc; a.dt = dt;
    a.apc = std::max(1, 256 / nn);
    // packed-neighbour kernel (diagonal coupling operators; tables built by heom_fast_prepare at plan creation)
    const int mode = heom_stage_mode(p, B);
    if (mode >= 1) {
        a.pk = p->d_pk.as<uint4>(); a.NK = p->max_nk + 1;
        a.ctab = p->d_ctab.as<cplx>() + (p->npar > 1 ? (size_t)b0 * nn * HEOM_PK_NE * a.NK : 0);
    }
    if (mode >= 2) {
        HeomFastArgs f;
        f.d = a.d; f.B = B; f.stage = stage; f.NK = p->max_nk + 1; f.apc = a.apc;
        f.HB = std::max(std::min(B, 4), ceil_div(B, 65535));
        f.row_lo = p->row_lo; f.row_hi = p->row_hi;
        f.rho = rho; f.yin = yin; f.ynext = ynext; f.acc = acc; f.dt = dt;
        f.pk = p->d_pk.as<uint4>();
        f.ctab = p->d_ctab.as<cplx>() + (p->npar > 1 ? (size_t)b0 * nn * HEOM_FLOW_NE * f.NK : 0);
        const dim3 fgrid((unsigned)ceil_div(nown, (long long)f.apc), (unsigned)ceil_div(B, f.HB));
        const int threads = ceil_div(f.apc * nn, 32) * 32;
        const size_t fsmem = (size_t)(1 + 2 * f.apc + HEOM_FLOW_NE) * nn * 16;
        // (compiled for 3 resident CTAs per SM, 80 registers; 2 / 4 CTAs measured within 2 % of it)
        const bool hr = p->h_real && !getenv("LIMEB200_HEOM_COMPLEX_H");
        const bool cb = p->npar == 1 && mode == 3;
#define LB_FAST(NN) (hr ? (cb ? heom_stage_fast_kernel<NN, true, true> : heom_stage_fast_kernel<NN, true, false>) \
                        : (cb ? heom_stage_fast_kernel<NN, false, true> : heom_stage_fast_kernel<NN, false, false>))
        void (*fk)(HeomFastArgs) = p->n == 7 ? LB_FAST(7) : p->n == 3 ? LB_FAST(3) : p->n == 2 ? LB_FAST(2)
                                 : (cb ? heom_stage_fast_kernel<0, false, true> : heom_stage_fast_kernel<0, false, false>);
#undef LB_FAST
        fk<<<fgrid, threads, fsmem, st>>>(f);
        p->launches++;
        return LB_OK;
    }
    size_t smem = (size_t)(1 + 3 * a.apc) * nn * 16;
    dim3 grid((unsigned)ceil_div(nown * B, (long long)a.apc));
    if (smem > 48 * 1024)
        LB_CUDA(cudaFuncSetAttribute(heom_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    heom_stage_kernel<<<grid, ceil_div(a.apc * nn, 32) * 32, smem, st>>>(a);
    p->launches++;
    return LB_OK;
}

// launch geometry of the persistent kernels for B hierarchies over the plan's owned ADO range
struct HeomPersistCfg {
    void (*kern)(HeomPersistArgs) = nullptr;
    int apc = 1, threads = 32, grid = 1;
    size_t smem = 0;
    bool one_tile_per_cta = false;
};
static int heom_persist_config(limeb200_heom_t p, int B, HeomPersistCfg& c, bool allow_cached = true) {
    const int nn = p->n * p->n;
    const long long total = p->nhe * nn;
    const long long nown = p->row_hi - p->row_lo;
    if (nown <= 0) return LB_ERR_UNSUPPORTED;
    int coop = 0;
    LB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device));
    if (!coop) return LB_ERR_UNSUPPORTED;
    // balanced tiles: one tile per CTA for the whole run when the items fit in (1 or 2 CTAs per SM) x
    // at most (1024 or 576) threads; larger problems loop over tiles of up to 1024 threads
    const long long nitems = nown * B;
    c.kern = heom_persist_kernel<1024, 1>;
    int per_sm = 1;
    {
        long long apc1 = ceil_div(nitems, (long long)p->sm_count);
        long long apc2 = ceil_div(nitems, 2LL * p->sm_count);
        if (apc1 * nn <= 1024) c.apc = (int)std::max<long long>(1, apc1);
        else if (apc2 * nn <= 576) { c.apc = (int)apc2; c.kern = heom_persist_kernel<576, 2>; per_sm = 2; }
        else c.apc = std::max(1, 1024 / nn);
    }
    c.threads = ceil_div(c.apc * nn, 32) * 32;
    c.smem = (size_t)(1 + 3 * c.apc) * nn * 16 + (size_t)3 * c.apc * p->nmodes * 4;
    c.one_tile_per_cta = ceil_div(nitems, (long long)c.apc) <= (long long)per_sm * p->sm_count;
    const size_t smem_c = (size_t)(1 + c.apc) * nn * 16 + (size_t)HEOM_PC_NE * c.threads * 20;
    if (allow_cached && !getenv("LIMEB200_HEOM_NO_CACHED") && p->diagq && c.one_tile_per_cta && 2 * p->max_modes_per_elem <= HEOM_PC_NE &&
        (long long)B * total < (1LL << 31) && smem_c * per_sm + 2048 <= (size_t)p->smem_optin) {
        c.kern = per_sm == 2 ? heom_persist_cached_kernel<576, 2> : heom_persist_cached_kernel<1024, 1>;
        c.smem = smem_c;
    }
    LB_CUDA(cudaFuncSetAttribute(c.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    int occ = 0;
    LB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c.kern, c.threads, c.smem));
    per_sm = std::min(per_sm, occ);
    if (per_sm < 1) return LB_ERR_UNSUPPORTED;
    const long long ntiles = ceil_div(nitems, (long long)c.apc);
    c.grid = (int)std::min<long long>(ntiles, (long long)per_sm * p->sm_count);
    return LB_OK;
}

// launch one of the persistent kernels over the plan's owned ADO range (pa: y0/y1, outputs and the
// sharding fields filled by the caller; y0 must already hold the state)
static int heom_launch_persist(limeb200_heom_t p, HeomPersistArgs& pa, cplx* rho, int B, double dt, int nsteps,
                               cudaStream_t st, bool require_one_tile_per_cta = false) {
    HeomPersistCfg c;
    int r = heom_persist_config(p, B, c, !pa.hybrid);      // the tagged-halo mode lives in the generic tile code
    if (r != LB_OK) return r;
    // larger problems gain nothing from persistence (the per-stage launches are already long) and
    // run better with the smaller CTAs of the stage-wise kernel
    if (require_one_tile_per_cta && !c.one_tile_per_cta) return LB_ERR_UNSUPPORTED;
    pa.s.d = p->dev();
    pa.s.B = B; pa.s.stage = 0; pa.s.row_lo = p->row_lo; pa.s.row_hi = p->row_hi;
    pa.s.rho = rho; pa.s.acc = p->s_acc.as<cplx>(); pa.s.yin = nullptr; pa.s.ynext = nullptr;
    pa.s.dt = dt;
    pa.s.npeer = 0; pa.s.peer_mask = nullptr; pa.s.send_all = 1;
    // The generic persistent kernel keeps the table walk: with its 64-register budget (1024 threads per CTA) the packed
    // gather spills and was measured 19 % SLOWER on the 38 760-ADO hierarchy over 2 GPUs (18.6 vs 15.1 ms per 50 steps,
    // profiles/r02_bench_default_2gpu_packed_persist.json vs r02_bench_default_2gpu.json).  LIMEB200_HEOM_STAGE_MODE=1
    // switches it on explicitly (tests).
    {
        const char* e = getenv("LIMEB200_HEOM_STAGE_MODE");
        if (p->fast_state == 1 && e && e[0] == '1' && !e[1] && !pa.hybrid && !getenv("LIMEB200_HEOM_NO_FAST_STAGE")) {
            pa.s.pk = p->d_pk.as<uint4>(); pa.s.ctab = p->d_ctab.as<cplx>(); pa.s.NK = p->max_nk + 1;
        }
    }
    pa.s.apc = c.apc;
    pa.nsteps = nsteps;
    if (!p->dbar.p) LB_CUDA(p->dbar.alloc(512));
    LB_CUDA(cudaMemsetAsync(p->dbar.p, 0, 512, st));
    pa.barrier = p->dbar.as<unsigned>();
    void* kargs[] = {&pa};
    LB_CUDA(cudaLaunchCooperativeKernel((void*)c.kern, dim3(c.grid), dim3(c.threads), kargs, c.smem, st));
    p->launches += 1;
    return LB_OK;
}

// ---- dataflow-synchronised persistent kernel (heom_flow.cuh)
static bool heom_flow_supported(limeb200_heom_t p) {
    const long long total = p->nhe * p->n * p->n;
    return p->diagq && p->npar == 1 && 2 * p->max_modes_per_elem <= HEOM_FLOW_NE && total < (1LL << 30) && p->max_nk < 64 &&
           p->n * p->n <= 1024 && p->row_hi > p->row_lo;
}
struct HeomFlowCfg { void (*kern)(HeomFlowArgs) = nullptr; int apc = 1, threads = 32, grid = 1; size_t smem = 0; bool cached = false; int ept = 1; };
static int heom_flow_config(limeb200_heom_t p, HeomFlowCfg& c) {
    const int nn = p->n * p->n;
    const long long nown = p->row_hi - p->row_lo;
    int coop = 0;
    LB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device));
    if (!coop || !heom_flow_supported(p)) return LB_ERR_UNSUPPORTED;
    // (a) register-resident variant, one or two elements per thread for the whole run: the owned ADOs fit one tile per
    //     CTA on one (<= 1024 threads, x EPT) or two (<= 576 threads) CTAs per SM
    if (p->nhe * nn < (1LL << HEOM_FLOW_IDXBITS) && !getenv("LIMEB200_HEOM_FLOW_TILED")) {
        const long long apc1 = ceil_div(nown, (long long)p->sm_count), apc2 = ceil_div(nown, 2LL * p->sm_count);
        int apc = 0, per_sm = 1, ept = 1;
        if (apc1 * nn <= 576) { apc = (int)apc1; }
        else if (apc2 * nn <= 576) { apc = (int)apc2; per_sm = 2; }
        else if (apc1 * nn <= 1024) { apc = (int)apc1; }
        else if (apc1 * nn <= 2048) { apc = (int)apc1; ept = 2; }
        if (apc > 0) {
            const int threads = ceil_div((int)ceil_div(apc * nn, ept), 32) * 32;
            const int cls = ept == 2 ? 3 : threads > 576 ? 2 : per_sm == 2 ? 1 : 0;
#define LB_FLOWC(NN) (cls == 3 ? heom_flow_cached_kernel<NN, 1024, 1, 2> : cls == 2 ? heom_flow_cached_kernel<NN, 1024, 1, 1> \
                      : cls == 1 ? heom_flow_cached_kernel<NN, 576, 2, 1> : heom_flow_cached_kernel<NN, 576, 1, 1>)
            c.kern = p->n == 7 ? LB_FLOWC(7) : p->n == 3 ? LB_FLOWC(3) : p->n == 2 ? LB_FLOWC(2) : LB_FLOWC(0);
#undef LB_FLOWC
            c.apc = apc; c.threads = threads; c.cached = true; c.ept = ept;
            c.grid = (int)ceil_div(nown, (long long)apc);
            c.smem = (size_t)(2 + apc) * nn * 16 + (size_t)nn * HEOM_FLOW_NE * (p->max_nk + 1) * 16 +
                     (size_t)ept * HEOM_FLOW_NE * threads * 4 + (ept == 1 ? (size_t)HEOM_FLOW_NE * threads * 16 : 0);
            LB_CUDA(cudaFuncSetAttribute(c.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
            int occ = 0;
            LB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c.kern, c.threads, c.smem));
            if (occ >= 1 && c.grid <= occ * p->sm_count) return LB_OK;
            c.cached = false;
        }
    }
    // (b) tiled variant: one CTA per SM, contiguous balanced blocks of ADOs, each walked in equal tiles of <= 1024 threads;
    //     a block that fits 576 threads takes the variant compiled for 576 threads (more registers, no spills)
    long long per = ceil_div(nown, (long long)p->sm_count);
    const bool small = per * nn <= 576;
    const int apc_max = std::max(1, (small ? 576 : 1024) / nn);
    const long long ntile = ceil_div(per, (long long)apc_max);
    c.apc = (int)ceil_div(per, ntile);
    c.threads = ceil_div(c.apc * nn, 32) * 32;
    if (small)
        c.kern = p->n == 7 ? heom_flow_kernel<7, 576> : p->n == 3 ? heom_flow_kernel<3, 576>
               : p->n == 2 ? heom_flow_kernel<2, 576> : heom_flow_kernel<0, 576>;
    else
        c.kern = p->n == 7 ? heom_flow_kernel<7, 1024> : p->n == 3 ? heom_flow_kernel<3, 1024>
               : p->n == 2 ? heom_flow_kernel<2, 1024> : heom_flow_kernel<0, 1024>;
    c.grid = (int)ceil_div(nown, per);
    c.smem = (size_t)(2 + c.apc) * nn * 16;
    LB_CUDA(cudaFuncSetAttribute(c.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    int occ = 0;
    LB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c.kern, c.threads, c.smem));
    if (occ < 1 || c.grid > occ * p->sm_count) return LB_ERR_UNSUPPORTED;
    return LB_OK;
}
// fa: T / Tp / npeer / peer_mask / tag0 / outputs filled by the caller; T[0] must already hold the tagged state
static int heom_flow_launch(limeb200_heom_t p, HeomFlowArgs& fa, cplx* rho, double dt, int nsteps, cudaStream_t st) {
    HeomFlowCfg c;
    int r = heom_flow_config(p, c);
    if (r != LB_OK) return r;
    const long long total = p->nhe * p->n * p->n;
    if (p->s_acc.bytes < (size_t)total * 16) LB_CUDA(p->s_acc.alloc((size_t)total * 16));
    if (!p->dbar.p) LB_CUDA(p->dbar.alloc(512));
    LB_CUDA(cudaMemsetAsync(p->dbar.p, 0, 512, st));
    fa.d = p->dev();
    fa.row_lo = p->row_lo; fa.row_hi = p->row_hi;
    fa.apc = c.apc; fa.nsteps = nsteps; fa.dt = dt; fa.maxn = p->max_nk;
    fa.rho = rho; fa.acc = p->s_acc.as<cplx>();
    fa.err = p->dbar.as<unsigned>() + 2;
    void* kargs[] = {&fa};
    LB_CUDA(cudaLaunchCooperativeKernel((void*)c.kern, dim3(c.grid), dim3(c.threads), kargs, c.smem, st));
    p->launches += 1;
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
    const size_t smem_chip = ((size_t)nbuf * total + nn) * 16;
    bool chip_ok = total <= 8 * 1024 && smem_chip + 1024 <= (size_t)p->smem_optin;
    int path = p->path_req;
    if (path == 0) path = (chip_ok && (B >= p->sm_count / 2 || total <= 4096)) ? 1 : 3;
    LB_REQUIRE(path != 1 || chip_ok, "hierarchy (%lld elements) does not fit the on-chip path", total);
    p->path = path;
    if (path == 1) {
        HeomChipArgs a;
        a.d = p->dev();
        a.B = B; a.nsteps = nsteps; a.traj_every = traj_every; a.E = E;
        a.ado = (cplx*)d_ado; a.eT = (const cplx*)d_eT; a.obs = E > 0 ? (cplx*)d_obs : nullptr; a.traj = (cplx*)d_traj;
        a.dt = dt;
        a.T = 0;
        // (measured on config 3: below ~2 hierarchies per SM the thread-per-element kernel has 4x the parallelism and is
        //  1.8x faster -- B = 1: 3.3e7 vs 1.8e7 ADO-steps/s; from B = 512 up the thread-per-ADO kernel wins)
        if (p->diagq && p->n == 2 && p->nmodes <= 4 && p->nhe <= 576 && B >= 2 * p->sm_count &&
            !getenv("LIMEB200_HEOM_NO_ADO_KERNEL")) {
            // one thread per ADO, HP hierarchies per CTA
            int HP = std::max(1, std::min(576 / (int)p->nhe, (int)((100 * 1024) / ((size_t)2 * p->nhe * nn * 16))));
            HP = std::min(HP, std::max(1, ceil_div(B, 2 * p->sm_count)));
            const size_t smem = (size_t)2 * nn * HP * p->nhe * 16;
            const int threads = ceil_div(HP * (int)p->nhe, 32) * 32;
            void (*k2)(HeomChipArgs, int) = p->nmodes <= 2 ? heom_onchip_ado_kernel<2, 2> : heom_onchip_ado_kernel<2, 4>;
            LB_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k2<<<ceil_div(B, HP), threads, smem, st>>>(a, HP);
            LB_CUDA(cudaGetLastError());
            p->launches++;
            return LB_OK;
        }
        int T = (int)std::min<long long>(1024, ceil_div(total, 32LL) * 32);
        int ept = (int)ceil_div(total, (long long)T);
        int EPT = ept <= 1 ? 1 : ept <= 2 ? 2 : ept <= 4 ? 4 : 8;
        a.T = T;
        void (*kern)(HeomChipArgs) = EPT == 1 ? heom_onchip_kernel<1> : EPT == 2 ? heom_onchip_kernel<2>
                                   : EPT == 4 ? heom_onchip_kernel<4> : heom_onchip_kernel<8>;
        if (p->diagq && EPT == 1 && p->max_modes_per_elem <= 4) {     // register-cached neighbour lists
            if (p->n == 2) kern = p->max_modes_per_elem <= 2 ? heom_onchip_cached<4, 2> : heom_onchip_cached<8, 2>;
            else kern = p->max_modes_per_elem <= 2 ? heom_onchip_cached<4, 0> : heom_onchip_cached<8, 0>;
        }
        LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chip));
        kern<<<B, T, smem_chip, st>>>(a);
        LB_CUDA(cudaGetLastError());
        p->launches++;
        return LB_OK;
    }
    // ---- scratch shared by the persistent and the stage-wise paths
    // (capacities are tracked in bytes: the sharded entry point allocates s_acc only)
    if (p->s_y.bytes < (size_t)2 * B * total * 16) LB_CUDA(p->s_y.alloc((size_t)2 * B * total * 16));
    if (p->s_acc.bytes < (size_t)B * total * 16) LB_CUDA(p->s_acc.alloc((size_t)B * total * 16));
    // (measured on one GPU, config 4: 6.7e7 ADO-steps/s against 9.4e7 of the barrier kernel -- tagged entries double the
    //  L2 bytes of the neighbour gather, which costs more than the grid barrier it saves; the dataflow kernel pays off
    //  where the barrier crosses NVLink, so on one GPU it is opt-in: path 4 or LIMEB200_HEOM_FLOW)
    if ((path == 4 || (path == 3 && p->path_req == 0 && getenv("LIMEB200_HEOM_FLOW"))) && B == 1) {
        // dataflow-synchronised persistent kernel: no barrier, every value carries its stage tag
        HeomFlowCfg c;
        int rc = heom_flow_config(p, c);
        if (rc == LB_OK) {
            if (p->s_T.bytes < (size_t)4 * total * 16) {
                LB_CUDA(p->s_T.alloc((size_t)4 * total * 16));
                LB_CUDA(cudaMemsetAsync(p->s_T.p, 0, (size_t)4 * total * 16, st));
            }
            HeomFlowArgs fa;
            memset(&fa, 0, sizeof(fa));
            fa.T[0] = p->s_T.as<ulonglong2>(); fa.T[1] = p->s_T.as<ulonglong2>() + (size_t)2 * total;
            fa.tag0 = p->flow_tag;
            p->flow_tag += 4ull * nsteps + 4;
            fa.E = E; fa.traj_every = traj_every;
            fa.eT = (const cplx*)d_eT; fa.obs = E > 0 ? (cplx*)d_obs : nullptr; fa.traj = (cplx*)d_traj;
            heom_flow_pack_kernel<<<std::min<long long>(ceil_div(total, 256LL), 148 * 8), 256, 0, st>>>(
                (const cplx*)d_ado, total, fa.tag0, fa.T[0]);
            rc = heom_flow_launch(p, fa, (cplx*)d_ado, dt, nsteps, st);
            if (rc == LB_OK) { p->path = 4; return LB_OK; }
        }
        if (rc != LB_ERR_UNSUPPORTED) return rc;
        LB_REQUIRE(p->path_req != 4, "the dataflow path needs diagonal coupling operators, one hierarchy and a cooperative launch");
        path = 3; p->path = 3;
    }
    LB_REQUIRE(path != 4, "the dataflow path takes one hierarchy (B = 1)");
    if (path == 3) {
        // persistent cooperative kernel: all steps in one launch, one grid barrier per stage
        HeomPersistArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.world = 1;
        pa.y0 = p->s_y.as<cplx>(); pa.y1 = p->s_y.as<cplx>() + (size_t)B * total;
        pa.eT = (const cplx*)d_eT; pa.obs = E > 0 ? (cplx*)d_obs : nullptr; pa.traj = (cplx*)d_traj;
        pa.E = E; pa.traj_every = traj_every;
        LB_CUDA(cudaMemcpyAsync(pa.y0, d_ado, (size_t)B * total * 16, cudaMemcpyDeviceToDevice, st));
        int r = heom_launch_persist(p, pa, (cplx*)d_ado, B, dt, nsteps, st, p->path_req == 0);
        if (r == LB_OK) return LB_OK;
        if (r != LB_ERR_UNSUPPORTED) return r;
        LB_REQUIRE(p->path_req != 3, "cooperative launch is not available on this device");
        p->path = 2;
    }
    // ---- stage-wise path
    cplx* rho = (cplx*)d_ado;
    cplx* y[2] = {p->s_y.as<cplx>(), p->s_y.as<cplx>() + (size_t)B * total};
    cplx* acc = p->s_acc.as<cplx>();
    LB_CUDA(cudaMemcpyAsync(y[0], rho, (size_t)B * total * 16, cudaMemcpyDeviceToDevice, st));
    // Optional L2 blocking over the batch (LIMEB200_HEOM_L2_CHUNK = slice size, 0 = sized to L2): a slice of `chunk`
    // hierarchies is walked through the whole time loop before the next one starts, so that the four vectors a stage
    // touches stay in L2.  Measured with heom_stage_kernel on the batch of 64 FMO hierarchies it LOSES 7 % (9.33e7 vs
    // 1.00e8 ADO-steps/s): that kernel is issue / latency bound, not DRAM bound, and smaller launches cost more than the
    // saved traffic buys -- so it is opt-in.
    int chunk = B;
    if (B > 1 && nsteps > 1 && getenv("LIMEB200_HEOM_L2_CHUNK")) {
        int l2 = 0;
        LB_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, p->device));
        const double ws = 4.0 * (double)total * 16.0;
        const long long fit = (long long)(0.6 * (double)l2 / ws);
        const long long min_items = 8LL * p->sm_count * std::max(1, 256 / nn);      // keep >= 8 CTAs per SM per launch
        if (fit >= 1 && fit < B && fit * p->nhe >= min_items) chunk = (int)fit;      // LIMEB200_HEOM_L2_CHUNK=0: automatic
        if (atoi(getenv("LIMEB200_HEOM_L2_CHUNK")) >= 1) chunk = std::min(B, atoi(getenv("LIMEB200_HEOM_L2_CHUNK")));
    }
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int cb = std::min(chunk, B - b0);
        for (int step = 0; step < nsteps; ++step) {
            for (int stage = 0; stage < 4; ++stage) {
                int r = heom_launch_stage(p, stage, rho, y[stage & 1], y[(stage + 1) & 1], acc, cb, dt, st, b0);
                if (r != LB_OK) return r;
            }
            const bool save = d_traj && ((step + 1) % traj_every) == 0;
            if (E > 0 || save) {
                heom_tier0_obs<<<cb, 64, 0, st>>>(rho + (size_t)b0 * total, total, nn, (const cplx*)d_eT, E,
                                                 E > 0 ? (cplx*)d_obs + ((size_t)step * B + b0) * E : nullptr,
                                                 save ? (cplx*)d_traj + ((size_t)(step / traj_every) * B + b0) * nn : nullptr);
                p->launches++;
            }
        }
    }
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

// ---- peer memory (CUDA IPC) for the ADO-sharded propagator: one process per GPU, one box
int limeb200_peer_alloc(int device, long long bytes, void** d_ptr, unsigned char* handle64) {
    LB_REQUIRE(d_ptr && handle64 && bytes > 0, "bad arguments");
    LB_CUDA(cudaSetDevice(device));
    void* ptr = nullptr;
    LB_CUDA(cudaMalloc(&ptr, (size_t)bytes));
    LB_CUDA(cudaMemset(ptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) { cudaFree(ptr); LB_CUDA(e); }
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    *d_ptr = ptr;
    return LB_OK;
}
int limeb200_peer_open(int device, const unsigned char* handle64, void** d_ptr) {
    LB_REQUIRE(d_ptr && handle64, "bad arguments");
    LB_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    LB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return LB_OK;
}
int limeb200_peer_close(int device, void* d_ptr) {
    LB_CUDA(cudaSetDevice(device));
    LB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return LB_OK;
}
int limeb200_peer_free(int device, void* d_ptr) {
    LB_CUDA(cudaSetDevice(device));
    LB_CUDA(cudaFree(d_ptr));
    return LB_OK;
}

int limeb200_heom_persist_grid(limeb200_heom_t p, int B) {
    LB_REQUIRE(p && B >= 1, "bad arguments");
    LB_CUDA(cudaSetDevice(p->device));
    HeomPersistCfg c;
    int r = heom_persist_config(p, B, c);
    if (r != LB_OK) { limeb200::set_error("persistent kernel cannot be launched for this plan"); return r; }
    return c.grid;
}

int limeb200_heom_run_sharded(limeb200_heom_t p, int rank, int world, void* const* d_y0, void* const* d_y1,
                              void* const* d_flags, const int* h_grids, double* d_rho,
                              const unsigned char* d_peer_mask, double dt, int nsteps, unsigned epoch, void* stream) {
    LB_REQUIRE(p && d_y0 && d_y1 && d_flags && d_rho && h_grids, "null argument");
    LB_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "world must be 1..8");
    LB_REQUIRE(p->npar == 1, "sharded runs take one hierarchy (no parameter batch)");
    LB_REQUIRE(nsteps >= 0, "bad nsteps");
    LB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    p->launches = 0;
    if (nsteps == 0) return LB_OK;
    const long long total = p->nhe * p->n * p->n;
    if (p->s_acc.bytes < (size_t)total * 16) LB_CUDA(p->s_acc.alloc((size_t)total * 16));
    HeomPersistArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.world = world; pa.rank = rank; pa.epoch = epoch;
    pa.peer_mask = d_peer_mask;
    pa.y0 = (cplx*)d_y0[rank]; pa.y1 = (cplx*)d_y1[rank]; pa.flags = (unsigned*)d_flags[rank];
    int q = 0;
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        pa.y0p[q] = (cplx*)d_y0[r]; pa.y1p[q] = (cplx*)d_y1[r]; pa.flagp[q] = (unsigned*)d_flags[r];
        LB_REQUIRE(h_grids[r] >= 1, "h_grids[%d] must be the peer's limeb200_heom_persist_grid", r);
        pa.peer_grid[q] = (unsigned)h_grids[r];
        ++q;
    }
    int r = heom_launch_persist(p, pa, (cplx*)d_rho, 1, dt, nsteps, st);
    if (r == LB_ERR_UNSUPPORTED) { limeb200::set_error("persistent sharded kernel cannot be launched on this device"); return r; }
    return r;
}
/* dataflow variant of the sharded propagator: tagged stage vectors instead of barriers (heom_flow.cuh) */
int limeb200_heom_flow_supported(limeb200_heom_t p) {
    LB_REQUIRE(p, "null plan");
    LB_CUDA(cudaSetDevice(p->device));
    HeomFlowCfg c;
    if (heom_flow_config(p, c) != LB_OK) return 0;
    // 2: the register-resident one-element-per-thread variant applies -- the latency-bound regime the kernel is built for
    // (measured on 8 GPUs, 38 760 ADOs = 2 elements per thread: 1.87e8 ADO-steps/s against 2.55e8 of the barrier kernel)
    return (c.cached && c.ept == 1) ? 2 : 1;
}
int limeb200_heom_flow_pack(limeb200_heom_t p, const double* d_y, void* d_T0, unsigned long long tag, void* stream) {
    LB_REQUIRE(p && d_y && d_T0, "null argument");
    LB_CUDA(cudaSetDevice(p->device));
    const long long total = p->nhe * p->n * p->n;
    heom_flow_pack_kernel<<<(int)std::min<long long>(ceil_div(total, 256LL), 148 * 8), 256, 0, (cudaStream_t)stream>>>(
        (const cplx*)d_y, total, tag, (ulonglong2*)d_T0);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}
int limeb200_heom_flow_unpack(limeb200_heom_t p, const void* d_T0, unsigned long long tag, double* d_y, void* stream) {
    LB_REQUIRE(p && d_y && d_T0, "null argument");
    LB_CUDA(cudaSetDevice(p->device));
    const long long total = p->nhe * p->n * p->n;
    if (!p->dbar.p) { LB_CUDA(p->dbar.alloc(512)); LB_CUDA(cudaMemsetAsync(p->dbar.p, 0, 512, (cudaStream_t)stream)); }
    heom_flow_unpack_kernel<<<(int)std::min<long long>(ceil_div(total, 256LL), 148 * 8), 256, 0, (cudaStream_t)stream>>>(
        (const ulonglong2*)d_T0, total, tag, (cplx*)d_y, p->dbar.as<unsigned>() + 2);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}
int limeb200_heom_flow_run_sharded(limeb200_heom_t p, int rank, int world, void* const* d_T0, void* const* d_T1,
                                   double* d_rho, const unsigned char* d_peer_mask, double dt, int nsteps,
                                   unsigned long long tag0, void* stream) {
    LB_REQUIRE(p && d_T0 && d_T1 && d_rho, "null argument");
    LB_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "world must be 1..8");
    LB_REQUIRE(nsteps >= 0 && tag0 >= 1, "bad nsteps / tag0");
    LB_CUDA(cudaSetDevice(p->device));
    p->launches = 0;
    if (nsteps == 0) return LB_OK;
    HeomFlowArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.T[0] = (ulonglong2*)d_T0[rank]; fa.T[1] = (ulonglong2*)d_T1[rank];
    int q = 0;
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        fa.Tp[0][q] = (ulonglong2*)d_T0[r]; fa.Tp[1][q] = (ulonglong2*)d_T1[r];
        ++q;
    }
    fa.npeer = world - 1;
    fa.peer_mask = d_peer_mask;
    fa.tag0 = tag0;
    int r = heom_flow_launch(p, fa, (cplx*)d_rho, dt, nsteps, (cudaStream_t)stream);
    if (r == LB_ERR_UNSUPPORTED) limeb200::set_error("the dataflow sharded kernel does not support this plan");
    return r;
}

/* hybrid sharded propagator: barrier kernel inside each GPU, tagged halo between the GPUs (no cross-GPU barrier) */
int limeb200_heom_run_sharded_halo(limeb200_heom_t p, int rank, int world, void* const* d_T0, void* const* d_T1,
                                   double* d_rho, const unsigned char* d_peer_mask, double dt, int nsteps,
                                   unsigned long long tag0, void* stream) {
    LB_REQUIRE(p && d_T0 && d_T1 && d_rho, "null argument");
    LB_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "world must be 1..8");
    LB_REQUIRE(p->npar == 1 && p->diagq, "the tagged-halo propagator takes one hierarchy with diagonal coupling operators");
    LB_REQUIRE(nsteps >= 0 && tag0 >= 1, "bad nsteps / tag0");
    LB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    p->launches = 0;
    if (nsteps == 0) return LB_OK;
    const long long total = p->nhe * p->n * p->n;
    if (p->s_y.bytes < (size_t)2 * total * 16) LB_CUDA(p->s_y.alloc((size_t)2 * total * 16));
    if (p->s_acc.bytes < (size_t)total * 16) LB_CUDA(p->s_acc.alloc((size_t)total * 16));
    HeomPersistArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.world = world; pa.rank = rank; pa.hybrid = 1; pa.tag0 = tag0;
    pa.peer_mask = d_peer_mask;
    pa.y0 = p->s_y.as<cplx>(); pa.y1 = p->s_y.as<cplx>() + (size_t)total;
    pa.T[0] = (ulonglong2*)d_T0[rank]; pa.T[1] = (ulonglong2*)d_T1[rank];
    int q = 0;
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        pa.Tp[0][q] = (ulonglong2*)d_T0[r]; pa.Tp[1][q] = (ulonglong2*)d_T1[r];
        ++q;
    }
    // local rows of the stage input come from y0; d_rho holds the full state on entry
    LB_CUDA(cudaMemcpyAsync(pa.y0, d_rho, (size_t)total * 16, cudaMemcpyDeviceToDevice, st));
    int r = heom_launch_persist(p, pa, (cplx*)d_rho, 1, dt, nsteps, st);
    if (r == LB_ERR_UNSUPPORTED) limeb200::set_error("persistent sharded kernel cannot be launched on this device");
    return r;
}

/* 1 when a bounded spin of the last sharded run timed out (a peer never arrived) */
int limeb200_heom_sharded_error(limeb200_heom_t p, void* stream) {
    LB_REQUIRE(p, "null plan");
    if (!p->dbar.p) return 0;
    unsigned w4[4] = {0, 0, 0, 0};
    LB_CUDA(cudaMemcpyAsync(w4, p->dbar.p, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    LB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return (int)(w4[2] | (w4[3] << 1));       // bit 0: a wait timed out; bit 1: unpack found a stale tag
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
