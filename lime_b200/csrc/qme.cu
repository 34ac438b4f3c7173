// Host side of the quantum-master-equation plan: operator analysis, kernel selection,
// launch loops.  C ABI declared in include/lime_b200.h.
#include "../../include/lime_b200.h"
#include "common.cuh"
#include "qme_dense.cuh"
#include "qme_sparse.cuh"
#include "qme_cluster.cuh"
#include "qme_band.cuh"
#include "qme_tile.cuh"
#include <algorithm>
#include <cstdlib>
#include <memory>

namespace {

struct HostOp {                 // one operator as handed over by the caller
    bool given = false;
    bool is_csr = false;
    int nb = 1;
    std::vector<hcplx> dense;                   // [nb][N*N]
    std::vector<int> indptr, indices;           // csr pattern
    std::vector<hcplx> data;                    // [nb][nnz]
};

struct EllHost {
    int w = 0;
    std::vector<int> col;       // [N][w]
    std::vector<hcplx> val;     // [nb][N][w]
};

}  // namespace

struct limeb200_qme_s {
    int N = 0, device = 0;
    int path_req = 0, path = 0;
    bool finalized = false;
    int nb = 1;
    bool step_vals = false;                     // the operator batch index is the RK4 step (time-dependent generator)
    HostOp G, Gr;                               // left generator, optional right generator (default G^H)
    std::vector<HostOp> X, Z;
    std::vector<std::vector<hcplx>> D, Dr;      // drives, dense: G_k += c D, Gr_k += c Dr (or conj(c) D^H)
    bool drive_conj = true;
    std::vector<hcplx> eops;                    // [E][N*N]
    int E = 0;
    long long launches = 0;

    // ---- device copies, dense path
    DevBuf dG, dGh, dX, dZh, dD, dDh, deT;
    // ---- device copies, sparse path
    DevBuf dGcol, dGval;
    DevBuf dXcol[QME_MAXS], dXval[QME_MAXS], dZcol[QME_MAXS], dZval[QME_MAXS];
    int wG = 0, wX[QME_MAXS] = {0}, wZ[QME_MAXS] = {0};
    int bandwidth = 0;                          // max |row - col| over all sparse operators (permuted basis)
    DevBuf dperm;                               // [N] new -> old (cluster path)
    bool permuted = false;
    DevBuf deptr, deidx, deval;
    // ---- band path (qme_band.cuh)
    DevBuf dbgd, dbgcol, dbgval, dbxcol[QME_BAND_MAXS], dbxval[QME_BAND_MAXS], dbzcol[QME_BAND_MAXS], dbzval[QME_BAND_MAXS];
    int band_noff = 0, band_gt = 0, band_xt = 0, band_C = 0, band_R = 0, band_chain = 0;
    std::vector<int> host_perm;                 // basis order chosen by finalize (new -> old)
    size_t band_smem = 0;
    // ---- register-patch path (qme_tile.cuh)
    DevBuf dtbrow, dtcolold, dtcolpos, dtcolc, dtrowc, dtxs0, dteptr, dterank, dteoff, dteval;
    int tile_NP = 0, tile_P = 0, tile_chunk = 0, tile_C = 0;
    size_t tile_smem = 0;
    // ---- scratch
    DevBuf s_y, s_acc, s_tmp, s_gk;
    int scratch_B = 0;
    int sm_count = 148;
    long long smem_optin = 0;
};

namespace {

void densify(const HostOp& op, int N, std::vector<hcplx>& out) {
    const size_t NN = (size_t)N * N;
    if (!op.is_csr) { out = op.dense; return; }
    out.assign((size_t)op.nb * NN, hcplx(0, 0));
    const size_t nnz = op.indices.size();
    for (int b = 0; b < op.nb; ++b)
        for (int i = 0; i < N; ++i)
            for (int p = op.indptr[i]; p < op.indptr[i + 1]; ++p)
                out[b * NN + (size_t)i * N + op.indices[p]] += op.data[b * nnz + p];
}

// union sparsity pattern over the batch -> ELL (padded with (col=row, val=0)) of the
// basis-permuted operator A'[i'][j'] = A[perm[i']][perm[j']]  (inv[old] = new)
void to_ell(const HostOp& op, int N, EllHost& e, int& bw, const std::vector<int>& perm,
            const std::vector<int>& inv) {
    const size_t NN = (size_t)N * N;
    std::vector<std::vector<std::pair<int, int>>> cols(N);   // per new row: (new col, source slot)
    for (int in = 0; in < N; ++in) {
        const int io = perm[in];
        if (op.is_csr) {
            for (int p = op.indptr[io]; p < op.indptr[io + 1]; ++p) cols[in].push_back({inv[op.indices[p]], p});
        } else {
            for (int jo = 0; jo < N; ++jo) {
                bool nz = false;
                for (int b = 0; b < op.nb && !nz; ++b) nz = op.dense[b * NN + (size_t)io * N + jo] != hcplx(0, 0);
                if (nz) cols[in].push_back({inv[jo], jo});
            }
        }
        std::sort(cols[in].begin(), cols[in].end());
    }
    // distinct columns per row (csr input may hold duplicates, which scipy sums)
    std::vector<std::vector<int>> ucol(N);
    int w = 1;
    for (int i = 0; i < N; ++i) {
        for (auto& c : cols[i])
            if (ucol[i].empty() || ucol[i].back() != c.first) ucol[i].push_back(c.first);
        w = std::max(w, (int)ucol[i].size());
        for (int c : ucol[i]) bw = std::max(bw, std::abs(c - i));
    }
    e.w = w;
    e.col.assign((size_t)N * w, 0);
    e.val.assign((size_t)op.nb * N * w, hcplx(0, 0));
    const size_t nnz = op.indices.size();
    for (int i = 0; i < N; ++i) {
        for (int q = 0; q < w; ++q) e.col[(size_t)i * w + q] = (q < (int)ucol[i].size()) ? ucol[i][q] : i;
        const int io = perm[i];
        for (auto& c : cols[i]) {
            int q = int(std::lower_bound(ucol[i].begin(), ucol[i].end(), c.first) - ucol[i].begin());
            for (int b = 0; b < op.nb; ++b) {
                hcplx v = op.is_csr ? op.data[b * nnz + c.second] : op.dense[b * NN + (size_t)io * N + c.second];
                e.val[((size_t)b * N + i) * w + q] += v;
            }
        }
    }
}

// reverse Cuthill-McKee ordering of the symmetrised union pattern; returns perm[new] = old
std::vector<int> rcm_order(int N, const std::vector<std::vector<int>>& adj) {
    std::vector<int> order;
    order.reserve(N);
    std::vector<char> seen(N, 0);
    auto deg = [&](int v) { return (int)adj[v].size(); };
    while ((int)order.size() < N) {
        int start = -1;
        for (int v = 0; v < N; ++v)
            if (!seen[v] && (start < 0 || deg(v) < deg(start))) start = v;
        // pseudo-peripheral start: repeat BFS from the farthest lowest-degree node
        for (int it = 0; it < 4; ++it) {
            std::vector<int> dist(N, -1), q{start};
            dist[start] = 0;
            for (size_t h = 0; h < q.size(); ++h)
                for (int u : adj[q[h]])
                    if (!seen[u] && dist[u] < 0) { dist[u] = dist[q[h]] + 1; q.push_back(u); }
            int far = q.back(), best = far;
            for (int v : q)
                if (dist[v] == dist[far] && deg(v) < deg(best)) best = v;
            if (best == start) break;
            start = best;
        }
        std::vector<int> q{start};
        seen[start] = 1;
        for (size_t h = 0; h < q.size(); ++h) {
            std::vector<int> nb;
            for (int u : adj[q[h]])
                if (!seen[u]) { seen[u] = 1; nb.push_back(u); }
            std::sort(nb.begin(), nb.end(), [&](int x, int y) { return deg(x) != deg(y) ? deg(x) < deg(y) : x < y; });
            q.insert(q.end(), nb.begin(), nb.end());
        }
        order.insert(order.end(), q.begin(), q.end());
    }
    std::reverse(order.begin(), order.end());
    return order;
}

void add_pattern(const HostOp& op, int N, std::vector<std::vector<int>>& adj) {
    const size_t NN = (size_t)N * N;
    auto link = [&](int i, int j) { if (i != j) { adj[i].push_back(j); adj[j].push_back(i); } };
    if (op.is_csr) {
        for (int i = 0; i < N; ++i)
            for (int p = op.indptr[i]; p < op.indptr[i + 1]; ++p) link(i, op.indices[p]);
    } else {
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                bool nz = false;
                for (int b = 0; b < op.nb && !nz; ++b) nz = op.dense[b * NN + (size_t)i * N + j] != hcplx(0, 0);
                if (nz) link(i, j);
            }
    }
}

int pattern_bandwidth(const std::vector<std::vector<int>>& adj, const std::vector<int>& inv) {
    int bw = 0;
    for (size_t i = 0; i < adj.size(); ++i)
        for (int j : adj[i]) bw = std::max(bw, std::abs(inv[i] - inv[j]));
    return bw;
}

int max_row_nnz(const HostOp& op, int N) {
    EllHost e; int bw = 0;
    std::vector<int> id(N);
    for (int i = 0; i < N; ++i) id[i] = i;
    to_ell(op, N, e, bw, id, id);
    return e.w;
}

void replicate(std::vector<hcplx>& v, size_t per, int nb) {   // [1][per] -> [nb][per]
    std::vector<hcplx> out((size_t)nb * per);
    for (int b = 0; b < nb; ++b) std::copy(v.begin(), v.begin() + per, out.begin() + (size_t)b * per);
    v.swap(out);
}

void adjoint(const hcplx* in, hcplx* out, int N) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) out[(size_t)j * N + i] = std::conj(in[(size_t)i * N + j]);
}

// Cuthill-McKee from every start node of every connected component; keeps the ordering with
// the smallest bandwidth (N is at most a few hundred on the paths that use it)
std::vector<int> best_cm_order(int N, const std::vector<std::vector<int>>& adj) {
    std::vector<int> comp(N, -1), order;
    order.reserve(N);
    int nc = 0;
    for (int s = 0; s < N; ++s) {
        if (comp[s] >= 0) continue;
        std::vector<int> members{s};
        comp[s] = nc;
        for (size_t hd = 0; hd < members.size(); ++hd)
            for (int u : adj[members[hd]])
                if (comp[u] < 0) { comp[u] = nc; members.push_back(u); }
        std::vector<int> best;
        int best_bw = N + 1;
        std::vector<int> pos(N, -1);
        for (int start : members) {
            std::vector<int> q{start};
            std::vector<char> seen(N, 0);
            seen[start] = 1;
            for (size_t hd = 0; hd < q.size(); ++hd) {
                std::vector<int> nbv;
                for (int u : adj[q[hd]])
                    if (!seen[u]) { seen[u] = 1; nbv.push_back(u); }
                std::sort(nbv.begin(), nbv.end(), [&](int x, int y) {
                    return adj[x].size() != adj[y].size() ? adj[x].size() < adj[y].size() : x < y; });
                q.insert(q.end(), nbv.begin(), nbv.end());
            }
            for (size_t i = 0; i < q.size(); ++i) pos[q[i]] = (int)i;
            int bw = 0;
            for (int v : q)
                for (int u : adj[v]) bw = std::max(bw, std::abs(pos[v] - pos[u]));
            if (bw < best_bw) { best_bw = bw; best = q; }
        }
        order.insert(order.end(), best.begin(), best.end());
        ++nc;
    }
    return order;
}

struct BandHost {
    int noff = 0, gt = 1, xt = 1, bw = 0;
    std::vector<hcplx> gd, gval;           // [nb][N], [nb][NOFF][N]
    std::vector<int> gcol;                 // [NOFF][N]
    std::vector<int> xcol[QME_BAND_MAXS], zcol[QME_BAND_MAXS];
    std::vector<hcplx> xval[QME_BAND_MAXS], zval[QME_BAND_MAXS];
};


// If the off-diagonal graph of G is a disjoint union of P <= 4 simple paths of equal length, return the
// interleaved ordering (new index = position * P + path) whose FULL pattern (G and all X_s, Z_s) has the
// smallest bandwidth over the 2^P path orientations.  In that ordering G couples row i only to rows i -/+ P.
bool chain_order(int N, const HostOp& G, const std::vector<std::vector<int>>& adj_all, std::vector<int>& order,
                 int& RS, int& bw_out) {
    std::vector<std::vector<int>> adj(N);
    add_pattern(G, N, adj);
    for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    for (int v = 0; v < N; ++v)
        if (adj[v].size() > 2) return false;
    std::vector<char> seen(N, 0);
    std::vector<std::vector<int>> paths;
    for (int v = 0; v < N; ++v) {
        if (seen[v] || adj[v].size() > 1) continue;           // start at path ends (degree 0 or 1)
        std::vector<int> path{v};
        seen[v] = 1;
        int cur = v;
        for (;;) {
            int nxt = -1;
            for (int u : adj[cur])
                if (!seen[u]) nxt = u;
            if (nxt < 0) break;
            seen[nxt] = 1;
            path.push_back(nxt);
            cur = nxt;
        }
        paths.push_back(path);
    }
    for (int v = 0; v < N; ++v)
        if (!seen[v]) return false;                            // a cycle
    const int P = (int)paths.size();
    if (P < 1 || P > 4) return false;
    const size_t L = paths[0].size();
    for (auto& pth : paths)
        if (pth.size() != L) return false;
    int best = N + 1;
    std::vector<int> cand(N), cinv(N);
    for (int mask = 0; mask < (1 << P); ++mask) {
        for (int k = 0; k < P; ++k)
            for (size_t pos = 0; pos < L; ++pos)
                cand[pos * P + k] = paths[k][(mask >> k) & 1 ? L - 1 - pos : pos];
        for (int i = 0; i < N; ++i) cinv[cand[i]] = i;
        const int bw = pattern_bandwidth(adj_all, cinv);
        if (bw < best) { best = bw; order = cand; }
    }
    RS = P;
    bw_out = best;
    return true;
}

// split the (permuted) operators into the tables of qme_band.cuh; false when the structure is
// outside what that kernel handles (more than 4 off-diagonal entries per row of G, more than
// one entry per row of an X_s / Z_s, more than 2 sandwich terms)
bool build_band_host(const HostOp& G, const std::vector<HostOp>& X, const std::vector<HostOp>& Z, int N, int nb,
                     const std::vector<int>& perm, const std::vector<int>& inv, BandHost& o, int chain_rs = 0) {
    const int S = (int)X.size();
    if (S > QME_BAND_MAXS || N > 128) return false;
    EllHost eg;
    to_ell(G, N, eg, o.bw, perm, inv);
    if (G.nb < nb) replicate(eg.val, (size_t)N * eg.w, nb);
    int noff = 0;
    for (int i = 0; i < N; ++i) {
        int c = 0;
        for (int q = 0; q < eg.w; ++q) c += eg.col[(size_t)i * eg.w + q] != i;
        noff = std::max(noff, c);
    }
    if (noff > 4) return false;
    o.noff = noff <= 2 ? 2 : 4;
    o.gd.assign((size_t)nb * N, hcplx(0, 0));
    o.gcol.assign((size_t)o.noff * N, 0);
    o.gval.assign((size_t)nb * o.noff * N, hcplx(0, 0));
    for (int i = 0; i < N; ++i) {
        int slot = 0;
        for (int q = 0; q < o.noff; ++q) o.gcol[(size_t)q * N + i] = i;
        for (int q = 0; q < eg.w; ++q) {
            const int c = eg.col[(size_t)i * eg.w + q];
            for (int b = 0; b < nb; ++b) {
                const hcplx v = eg.val[((size_t)b * N + i) * eg.w + q];
                if (c == i) o.gd[(size_t)b * N + i] += v;
                else {
                    o.gval[((size_t)b * o.noff + slot) * N + i] = v;
                    if (v.real() != 0.0) o.gt = 0;
                }
            }
            if (c != i) { o.gcol[(size_t)slot * N + i] = c; ++slot; }
        }
        if (chain_rs > 0) {
            // directional slots: slot 0 = row i - RS, slot 1 = row i + RS (absent neighbours: value 0)
            int cols[2] = {o.gcol[i], o.gcol[(size_t)N + i]};
            std::vector<hcplx> v0(nb), v1(nb);
            for (int b = 0; b < nb; ++b) { v0[b] = o.gval[((size_t)b * 2 + 0) * N + i]; v1[b] = o.gval[((size_t)b * 2 + 1) * N + i]; }
            for (int q = 0; q < 2; ++q) {
                o.gcol[(size_t)q * N + i] = i;
                for (int b = 0; b < nb; ++b) o.gval[((size_t)b * 2 + q) * N + i] = hcplx(0, 0);
            }
            for (int q = 0; q < 2; ++q) {
                if (cols[q] == i) continue;
                const int dst = cols[q] == i - chain_rs ? 0 : (cols[q] == i + chain_rs ? 1 : -1);
                if (dst < 0) return false;
                o.gcol[(size_t)dst * N + i] = cols[q];
                for (int b = 0; b < nb; ++b) o.gval[((size_t)b * 2 + dst) * N + i] = q == 0 ? v0[b] : v1[b];
            }
        }
    }
    if (chain_rs > 0 && (o.noff != 2 || !o.gt)) return false;
    for (int s = 0; s < S; ++s) {
        const HostOp* ops[2] = {&X[s], &Z[s]};
        std::vector<int>* cols[2] = {&o.xcol[s], &o.zcol[s]};
        std::vector<hcplx>* vals[2] = {&o.xval[s], &o.zval[s]};
        for (int t = 0; t < 2; ++t) {
            EllHost e;
            to_ell(*ops[t], N, e, o.bw, perm, inv);
            if (e.w != 1) return false;
            if (ops[t]->nb < nb) replicate(e.val, (size_t)N, nb);
            *cols[t] = e.col;
            *vals[t] = e.val;
            for (const hcplx& v : e.val)
                if (v.imag() != 0.0) o.xt = 0;
        }
    }
    return true;
}


// ---- register-patch path (qme_tile.cuh): structure analysis and tables -------------------------------------------
struct TileHost {
    int NP = 0, P = 0, L = 0, Lp = 0, C = 0, chunk = 0;
    size_t smem = 0;
    std::vector<int> order;                 // chain order (unpadded) q -> old index, for get_info
    std::vector<int> brow, colold, colpos, xs0, eptr, erank, eoff;
    std::vector<double> colc, rowc;
    std::vector<hcplx> eval;
};

// Does the problem have the structure qme_tile_kernel needs?  G: off-diagonal graph = P simple paths of equal
// length L (tridiagonal in chain order) with purely imaginary off-diagonal values; every X_s, Z_s: one real entry
// per row, and in chain order the source rows of TR consecutive rows are consecutive rows that the owning CTA holds
// (own or halo rows).  Paths are padded to a multiple of C*TR rows with empty rows.
bool build_tile_host(const HostOp& G, const std::vector<HostOp>& X, const std::vector<HostOp>& Z, int N, int nb, int E,
                     const std::vector<hcplx>& eops, long long smem_optin, TileHost& o) {
    constexpr int TR = 4;
    const int S = (int)X.size();
    if (S > QME_TILE_MAXS || N < 4 || N > 128) return false;
    std::vector<int> id(N);
    for (int i = 0; i < N; ++i) id[i] = i;
    EllHost eg;
    int bw = 0;
    to_ell(G, N, eg, bw, id, id);
    if (G.nb < nb) replicate(eg.val, (size_t)N * eg.w, nb);
    // dense views of the (few) coefficients: gdiag[b][i], goff[b][i][j] through a per-row lookup
    auto gval = [&](int b, int i, int j) -> hcplx {
        hcplx v(0, 0);
        for (int q = 0; q < eg.w; ++q)
            if (eg.col[(size_t)i * eg.w + q] == j) v += eg.val[((size_t)b * N + i) * eg.w + q];
        return v;
    };
    std::vector<std::vector<int>> adj(N);
    for (int i = 0; i < N; ++i)
        for (int q = 0; q < eg.w; ++q) {
            const int j = eg.col[(size_t)i * eg.w + q];
            if (j == i) continue;
            bool nz = false;
            for (int b = 0; b < nb; ++b) {
                const hcplx v = eg.val[((size_t)b * N + i) * eg.w + q];
                if (v != hcplx(0, 0)) nz = true;
                if (v.real() != 0.0) return false;            // off-diagonal G must be purely imaginary
            }
            if (nz) { adj[i].push_back(j); adj[j].push_back(i); }
        }
    for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    for (int v = 0; v < N; ++v)
        if (adj[v].size() > 2) return false;
    std::vector<char> seen(N, 0);
    std::vector<std::vector<int>> paths;
    for (int v = 0; v < N; ++v) {
        if (seen[v] || adj[v].size() > 1) continue;
        std::vector<int> path{v};
        seen[v] = 1;
        for (int cur = v;;) {
            int nxt = -1;
            for (int u : adj[cur])
                if (!seen[u]) nxt = u;
            if (nxt < 0) break;
            seen[nxt] = 1;
            path.push_back(nxt);
            cur = nxt;
        }
        paths.push_back(path);
    }
    for (int v = 0; v < N; ++v)
        if (!seen[v]) return false;                            // a cycle
    const int P = (int)paths.size();
    if (P < 1 || P > 4) return false;
    const int L = (int)paths[0].size();
    for (auto& pth : paths)
        if ((int)pth.size() != L) return false;
    // sandwich operators: one real entry per row
    std::vector<EllHost> ex(S), ez(S);
    for (int s = 0; s < S; ++s) {
        int bws = 0;
        to_ell(X[s], N, ex[s], bws, id, id);
        to_ell(Z[s], N, ez[s], bws, id, id);
        if (ex[s].w != 1 || ez[s].w != 1) return false;
        if (X[s].nb < nb) replicate(ex[s].val, (size_t)N, nb);
        if (Z[s].nb < nb) replicate(ez[s].val, (size_t)N, nb);
        for (const hcplx& v : ex[s].val) if (v.imag() != 0.0) return false;
        for (const hcplx& v : ez[s].val) if (v.imag() != 0.0) return false;
    }
    auto has_entry = [&](const EllHost& e, int i) {
        for (int b = 0; b < nb; ++b)
            if (e.val[(size_t)b * N + i] != hcplx(0, 0)) return true;
        return false;
    };
    // geometry: smallest cluster whose CTAs fit (<= 16 warps, shared memory)
    for (int C = 1; C <= 8; C *= 2) {
        const int Lp = ceil_div(L, C * TR) * C * TR;
        const int chunk = Lp / C;
        const int Np = P * Lp;
        if (Np > 128) continue;
        const int NP = Np <= 64 ? 64 : 128;
        const int warps = (P * chunk / TR) * (NP / 64);
        if (warps > 16) continue;
        const size_t smem = qme_tile_smem(NP, P, chunk, C, E);
        if (smem > (size_t)smem_optin) continue;
        const int NBR = P * (chunk + 2), R = P * chunk;
        // orientation of the paths
        for (int mask = 0; mask < (1 << P); ++mask) {
            std::vector<int> qold((size_t)Np, -1), inv(N, -1);      // padded chain index q = p*Lp + n
            for (int pth = 0; pth < P; ++pth)
                for (int n = 0; n < L; ++n) {
                    const int old = paths[pth][(mask >> pth) & 1 ? L - 1 - n : n];
                    qold[(size_t)pth * Lp + n] = old;
                    inv[old] = pth * Lp + n;
                }
            // buffer row of chain row q in CTA c (own or halo), -1 if the CTA does not hold it
            auto bufrow = [&](int c, int q) {
                const int pth = q / Lp, t = q % Lp - c * chunk + 1;
                return (t >= 0 && t <= chunk + 1) ? pth * (chunk + 2) + t : -1;
            };
            std::vector<int> xs0((size_t)C * (R / TR) * QME_TILE_MAXS, 0);
            bool ok = true;
            for (int c = 0; c < C && ok; ++c)
                for (int g = 0; g < R / TR && ok; ++g) {
                    const int own0 = g * TR, pth = own0 / chunk, t0 = 1 + own0 % chunk;
                    const int q0 = pth * Lp + c * chunk + t0 - 1;
                    for (int s = 0; s < S && ok; ++s) {
                        int base = pth * (chunk + 2) + t0;                 // default: the patch's own rows (coefficient 0)
                        bool have = false;
                        for (int r = 0; r < TR && ok; ++r) {
                            const int old = qold[q0 + r];
                            if (old < 0 || !has_entry(ex[s], old)) continue;
                            const int br = bufrow(c, inv[ex[s].col[old]]);
                            if (br < 0) { ok = false; break; }
                            if (have && br - r != base) ok = false;
                            base = br - r; have = true;
                        }
                        if (!ok) break;
                        if (base < 0 || base + TR - 1 >= NBR) { ok = false; break; }
                        // rows read from a halo row must belong to a patch that pushes in that direction (see qme_tile.cuh)
                        for (int r = 0; r < TR; ++r) {
                            const int t = (base + r) % (chunk + 2);
                            if (t == 0 && t0 != 1 && have) ok = false;
                            if (t == chunk + 1 && t0 + TR - 1 != chunk && have) ok = false;
                        }
                        xs0[((size_t)c * (R / TR) + g) * QME_TILE_MAXS + s] = base;
                    }
                }
            if (!ok) continue;
            // ---- tables
            o.NP = NP; o.P = P; o.L = L; o.Lp = Lp; o.C = C; o.chunk = chunk; o.smem = smem;
            o.xs0 = xs0;
            o.order.clear();
            for (int q = 0; q < Np; ++q)
                if (qold[q] >= 0) o.order.push_back(qold[q]);
            auto posof = [&](int q) { return 64 * (q / 64) + 32 * (q % 2) + (q % 64) / 2; };
            auto same_path = [&](int q1, int q2) {
                return q1 >= 0 && q2 >= 0 && q1 < Np && q2 < Np && q1 / Lp == q2 / Lp && qold[q1] >= 0 && qold[q2] >= 0;
            };
            o.brow.assign((size_t)C * NBR, -1);
            for (int c = 0; c < C; ++c)
                for (int pth = 0; pth < P; ++pth)
                    for (int t = 0; t <= chunk + 1; ++t) {
                        const int n = c * chunk + t - 1;
                        if (n >= 0 && n < Lp) o.brow[(size_t)c * NBR + pth * (chunk + 2) + t] = qold[(size_t)pth * Lp + n];
                    }
            o.colold.assign(NP, -1);
            o.colpos.assign((size_t)NP * 4, 0);
            o.colc.assign((size_t)nb * NP * QME_TILE_COLC, 0.0);
            for (int pos = 0; pos < NP; ++pos) {
                const int q = 64 * (pos / 64) + 2 * (pos % 32) + (pos % 64) / 32;
                for (int k = 0; k < 4; ++k) o.colpos[(size_t)pos * 4 + k] = pos * 16;
                if (q >= Np || qold[q] < 0) continue;
                const int old = qold[q];
                o.colold[pos] = old;
                const bool hl = same_path(q, q - 1), hr = same_path(q, q + 1);
                if (hl) o.colpos[(size_t)pos * 4 + 0] = posof(q - 1) * 16;
                if (hr) o.colpos[(size_t)pos * 4 + 1] = posof(q + 1) * 16;
                for (int s = 0; s < S; ++s)
                    if (has_entry(ez[s], old)) o.colpos[(size_t)pos * 4 + 2 + s] = posof(inv[ez[s].col[old]]) * 16;
                for (int b = 0; b < nb; ++b) {
                    double* cc = &o.colc[((size_t)b * NP + pos) * QME_TILE_COLC];
                    const hcplx gd = gval(b, old, old);
                    cc[0] = gd.real(); cc[1] = -gd.imag();                         // conj(G_jj)
                    cc[2] = hl ? -gval(b, old, qold[q - 1]).imag() : 0.0;         // conj(G[j][j-1]) = i cL
                    cc[3] = hr ? -gval(b, old, qold[q + 1]).imag() : 0.0;
                    for (int s = 0; s < S; ++s) cc[4 + s] = ez[s].val[(size_t)b * N + old].real();
                }
            }
            o.rowc.assign((size_t)nb * C * R * QME_TILE_ROWC, 0.0);
            for (int b = 0; b < nb; ++b)
                for (int c = 0; c < C; ++c)
                    for (int ow = 0; ow < R; ++ow) {
                        const int pth = ow / chunk, q = pth * Lp + c * chunk + ow % chunk;
                        const int old = qold[q];
                        if (old < 0) continue;
                        double* rcp = &o.rowc[(((size_t)b * C + c) * R + ow) * QME_TILE_ROWC];
                        const hcplx gd = gval(b, old, old);
                        rcp[0] = gd.real(); rcp[1] = gd.imag();
                        rcp[2] = same_path(q, q - 1) ? gval(b, old, qold[q - 1]).imag() : 0.0;
                        rcp[3] = same_path(q, q + 1) ? gval(b, old, qold[q + 1]).imag() : 0.0;
                        for (int s = 0; s < S; ++s) rcp[4 + s] = ex[s].val[(size_t)b * N + old].real();
                    }
            // observables: Tr(e rho) = sum_ij e[j][i] rho[i][j]
            const size_t NN = (size_t)N * N;
            o.eptr.assign(E + 1, 0);
            o.erank.clear(); o.eoff.clear(); o.eval.clear();
            for (int e = 0; e < E; ++e) {
                for (int qi = 0; qi < Np; ++qi)
                    for (int qj = 0; qj < Np; ++qj) {
                        if (qold[qi] < 0 || qold[qj] < 0) continue;
                        const hcplx v = eops[e * NN + (size_t)qold[qj] * N + qold[qi]];
                        if (v == hcplx(0, 0)) continue;
                        const int n = qi % Lp, c = n / chunk;
                        o.erank.push_back(c);
                        o.eoff.push_back(((qi / Lp) * (chunk + 2) + n - c * chunk + 1) * NP + posof(qj));
                        o.eval.push_back(v);
                    }
                o.eptr[e + 1] = (int)o.erank.size();
            }
            if (o.erank.empty()) { o.erank.push_back(-1); o.eoff.push_back(0); o.eval.push_back(hcplx(0, 0)); }
            return true;
        }
    }
    return false;
}

int set_op_dense(limeb200_qme_t p, HostOp& op, const double* h, int nb) {
    LB_REQUIRE(p && h, "null argument");
    LB_REQUIRE(!p->finalized, "plan already finalized");
    LB_REQUIRE(nb >= 1, "nb must be >= 1");
    const size_t NN = (size_t)p->N * p->N;
    op.given = true; op.is_csr = false; op.nb = nb;
    const hcplx* c = reinterpret_cast<const hcplx*>(h);
    op.dense.assign(c, c + (size_t)nb * NN);
    return LB_OK;
}

int set_op_csr(limeb200_qme_t p, HostOp& op, const int* indptr, const int* indices, const double* data,
               int nnz, int nb) {
    LB_REQUIRE(p && indptr && (nnz == 0 || (indices && data)), "null argument");
    LB_REQUIRE(!p->finalized, "plan already finalized");
    LB_REQUIRE(nb >= 1 && nnz >= 0, "bad nb/nnz");
    LB_REQUIRE(indptr[0] == 0 && indptr[p->N] == nnz, "csr indptr inconsistent with nnz");
    for (int i = 0; i < nnz; ++i) LB_REQUIRE(indices[i] >= 0 && indices[i] < p->N, "csr column out of range");
    op.given = true; op.is_csr = true; op.nb = nb;
    op.indptr.assign(indptr, indptr + p->N + 1);
    op.indices.assign(indices, indices + nnz);
    const hcplx* c = reinterpret_cast<const hcplx*>(data);
    op.data.assign(c, c + (size_t)nb * nnz);
    return LB_OK;
}

}  // namespace

extern "C" {

int limeb200_qme_create(limeb200_qme_t* plan, int N, int device) {
    LB_REQUIRE(plan, "null plan pointer");
    LB_REQUIRE(N >= 1 && N <= 8192, "N out of range (1..8192)");
    if (device == -1) {
        // analysis-only plan (host): kernel selection, basis ordering and cluster geometry for a B200 (148 SMs,
        // 227 KB shared memory per block) can be inspected with limeb200_qme_get_info; it cannot run
        auto* p = new limeb200_qme_s();
        p->N = N; p->device = -1; p->sm_count = 148; p->smem_optin = 232448;
        *plan = p;
        return LB_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        limeb200::set_error("no CUDA device available (%s): liblime_b200 has no CPU fallback",
                            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return LB_ERR_CUDA;
    }
    LB_REQUIRE(device >= 0 && device < ndev, "device %d out of range", device);
    LB_CUDA(cudaSetDevice(device));
    auto* p = new limeb200_qme_s();
    p->N = N; p->device = device;
    cudaDeviceProp prop;
    LB_CUDA(cudaGetDeviceProperties(&prop, device));
    p->sm_count = prop.multiProcessorCount;
    p->smem_optin = (long long)prop.sharedMemPerBlockOptin;
    *plan = p;
    return LB_OK;
}

int limeb200_qme_destroy(limeb200_qme_t plan) {
    if (plan) { if (plan->device >= 0) cudaSetDevice(plan->device); delete plan; }
    return LB_OK;
}

int limeb200_qme_set_generator_dense(limeb200_qme_t p, const double* h_G, int nb) {
    LB_REQUIRE(p, "null plan");
    return set_op_dense(p, p->G, h_G, nb);
}
int limeb200_qme_set_generator_csr(limeb200_qme_t p, const int* indptr, const int* indices,
                                   const double* h_data, int nnz, int nb) {
    LB_REQUIRE(p, "null plan");
    return set_op_csr(p, p->G, indptr, indices, h_data, nnz, nb);
}
int limeb200_qme_add_sandwich_dense(limeb200_qme_t p, const double* h_X, const double* h_Z, int nb) {
    LB_REQUIRE(p, "null plan");
    p->X.emplace_back(); p->Z.emplace_back();
    int r = set_op_dense(p, p->X.back(), h_X, nb);
    if (r == LB_OK) r = set_op_dense(p, p->Z.back(), h_Z, nb);
    if (r != LB_OK) { p->X.pop_back(); p->Z.pop_back(); }
    return r;
}
int limeb200_qme_add_sandwich_csr(limeb200_qme_t p,
                                  const int* x_indptr, const int* x_indices, const double* h_xdata, int x_nnz,
                                  const int* z_indptr, const int* z_indices, const double* h_zdata, int z_nnz,
                                  int nb) {
    LB_REQUIRE(p, "null plan");
    p->X.emplace_back(); p->Z.emplace_back();
    int r = set_op_csr(p, p->X.back(), x_indptr, x_indices, h_xdata, x_nnz, nb);
    if (r == LB_OK) r = set_op_csr(p, p->Z.back(), z_indptr, z_indices, h_zdata, z_nnz, nb);
    if (r != LB_OK) { p->X.pop_back(); p->Z.pop_back(); }
    return r;
}
int limeb200_qme_set_right_generator_dense(limeb200_qme_t p, const double* h_Gr, int nb) {
    LB_REQUIRE(p, "null plan");
    return set_op_dense(p, p->Gr, h_Gr, nb);
}
int limeb200_qme_add_drive_dense(limeb200_qme_t p, const double* h_D, const double* h_Dr) {
    LB_REQUIRE(p && h_D, "null argument");
    LB_REQUIRE(!p->finalized, "plan already finalized");
    LB_REQUIRE(p->D.empty() || (h_Dr != nullptr) == !p->drive_conj, "all drives must agree on passing h_Dr");
    const size_t NN = (size_t)p->N * p->N;
    const hcplx* c = reinterpret_cast<const hcplx*>(h_D);
    p->D.emplace_back(c, c + NN);
    if (h_Dr) {
        const hcplx* r = reinterpret_cast<const hcplx*>(h_Dr);
        p->Dr.emplace_back(r, r + NN);
        p->drive_conj = false;
    } else {
        std::vector<hcplx> dh(NN);
        adjoint(c, dh.data(), p->N);
        p->Dr.push_back(dh);
        p->drive_conj = true;
    }
    return LB_OK;
}
int limeb200_qme_set_observables(limeb200_qme_t p, const double* h_e, int E) {
    LB_REQUIRE(p && (E == 0 || h_e), "null argument");
    LB_REQUIRE(!p->finalized, "plan already finalized");
    LB_REQUIRE(E >= 0 && E <= 64, "E out of range (0..64)");
    const hcplx* c = reinterpret_cast<const hcplx*>(h_e);
    p->eops.assign(c, c + (size_t)E * p->N * p->N);
    p->E = E;
    return LB_OK;
}
int limeb200_qme_set_step_values(limeb200_qme_t p, int on) {
    LB_REQUIRE(p, "null plan");
    LB_REQUIRE(!p->finalized, "plan already finalized");
    p->step_vals = on != 0;
    return LB_OK;
}
int limeb200_qme_set_path(limeb200_qme_t p, int path) {
    LB_REQUIRE(p, "null plan");
    LB_REQUIRE(!p->finalized, "plan already finalized");
    LB_REQUIRE(path >= 0 && path <= 6, "path must be 0..6");
    p->path_req = path;
    return LB_OK;
}
int limeb200_qme_get_path(limeb200_qme_t p) { return p ? p->path : LB_ERR_ARG; }
int limeb200_qme_get_info(limeb200_qme_t p, int* info9, int* perm) {
    LB_REQUIRE(p && p->finalized && info9, "plan not finalized");
    info9[0] = p->path; info9[1] = p->permuted ? 1 : 0; info9[2] = p->bandwidth;
    info9[3] = p->band_noff; info9[4] = p->band_gt; info9[5] = p->band_xt;
    info9[6] = p->band_C; info9[7] = p->band_R; info9[8] = p->band_chain;
    if (p->path == 6) { info9[6] = p->tile_C; info9[7] = p->tile_P * p->tile_chunk; info9[8] = p->tile_P; }
    if (perm)
        for (int i = 0; i < p->N; ++i) perm[i] = i < (int)p->host_perm.size() ? p->host_perm[i] : i;
    return LB_OK;
}
long long limeb200_qme_last_launches(limeb200_qme_t p) { return p ? p->launches : -1; }

int limeb200_qme_finalize(limeb200_qme_t p) {
    LB_REQUIRE(p, "null plan");
    LB_REQUIRE(!p->finalized, "plan already finalized");
    LB_REQUIRE(p->G.given, "generator not set");
    if (p->device >= 0) LB_CUDA(cudaSetDevice(p->device));
    const int N = p->N, S = (int)p->X.size(), nd = (int)p->D.size();
    const size_t NN = (size_t)N * N;
    // operator batch
    int nb = p->G.nb;
    for (int s = 0; s < S; ++s) nb = std::max(nb, std::max(p->X[s].nb, p->Z[s].nb));
    auto okb = [&](const HostOp& o) { return o.nb == 1 || o.nb == nb; };
    LB_REQUIRE(okb(p->G), "generator batch %d incompatible with %d", p->G.nb, nb);
    for (int s = 0; s < S; ++s)
        LB_REQUIRE(okb(p->X[s]) && okb(p->Z[s]), "sandwich %d batch incompatible with %d", s, nb);
    p->nb = nb;

    // ---- choose path
    int wg = max_row_nnz(p->G, N), wprod = 1;
    for (int s = 0; s < S; ++s) wprod = std::max(wprod, max_row_nnz(p->X[s], N) * max_row_nnz(p->Z[s], N));
    const bool sparse_ok = (nd == 0) && !p->Gr.given && S <= QME_MAXS && wg <= 16 && wprod <= 16;
    const bool dense_mem_ok = (double)nb * NN * 16.0 * (2 + 2 * S) < 8e9;
    int path = p->path_req;
    if (p->step_vals) {
        // per-step generator values: the global-scratch sparse kernel is the one that re-reads the values every step
        LB_REQUIRE(sparse_ok, "per-step operator values need sparse operands without dense drives");
        path = 3;
    }
    if (path == 0) {
        if (N >= 24 && sparse_ok && wg * 4 <= N) path = 5;          // structured mid/large N
        else if (N <= 64 && dense_mem_ok) path = 1;
        else if (dense_mem_ok) path = 2;
        else path = 3;
    }
    // the cluster paths work in a bandwidth-reducing basis order
    std::vector<int> perm(N), inv(N);
    for (int i = 0; i < N; ++i) perm[i] = inv[i] = i;
    BandHost band;
    TileHost tile;
    if (path == 6 || (path == 5 && p->path_req == 0 && N >= 64 && !getenv("LIMEB200_NO_TILE"))) {
        // register-patch kernel for chain-structured operators; anything else goes to the band kernel
        const bool ok6 = sparse_ok && build_tile_host(p->G, p->X, p->Z, N, nb, p->E, p->eops, p->smem_optin, tile);
        LB_REQUIRE(ok6 || p->path_req != 6, "register-patch path does not fit this problem (N=%d)", N);
        path = ok6 ? 6 : 5;
    }
    if (path == 4 || path == 5) {
        bool fits = sparse_ok;
        if (fits) {
            std::vector<std::vector<int>> adj(N);
            add_pattern(p->G, N, adj);
            for (int s = 0; s < S; ++s) { add_pattern(p->X[s], N, adj); add_pattern(p->Z[s], N, adj); }
            for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
            int bw_best = pattern_bandwidth(adj, inv);
            std::vector<std::vector<int>> cands;
            cands.push_back(rcm_order(N, adj));
            if (N <= 512) {
                cands.push_back(best_cm_order(N, adj));
                cands.push_back(cands.back());
                std::reverse(cands.back().begin(), cands.back().end());
            }
            for (auto& cand : cands) {
                std::vector<int> cinv(N);
                for (int i = 0; i < N; ++i) cinv[cand[i]] = i;
                int bw = pattern_bandwidth(adj, cinv);
                if (bw < bw_best) { bw_best = bw; perm = cand; inv = cinv; p->permuted = true; }
            }
            p->band_chain = 0;
            if (path == 5 && getenv("LIMEB200_BAND_CHAIN")) {     // opt-in: measured SLOWER (1.40e6 vs 2.0-2.2e6 rho-steps/s on config 2, spills at 128 registers)
                // chain structure: interleaved path ordering + sliding-window kernel
                std::vector<int> corder, cinv(N);
                int rs = 0, cbw = 0;
                if (chain_order(N, p->G, adj, corder, rs, cbw)) {
                    for (int i = 0; i < N; ++i) cinv[corder[i]] = i;
                    BandHost cb;
                    int C = 0, R = 0;
                    size_t sm = 0;
                    if (build_band_host(p->G, p->X, p->Z, N, nb, corder, cinv, cb, rs) && cb.xt &&
                        qme_band_geometry(N, p->E, cb.bw, cb.noff, S, p->smem_optin, &C, &R, &sm) &&
                        C * R == N && R % (4 * rs) == 0 && N % 64 == 0) {
                        band = cb; perm = corder; inv = cinv; p->permuted = true;
                        p->band_C = C; p->band_R = R; p->band_smem = sm; p->band_chain = rs;
                    }
                }
            }
            if (path == 5 && !p->band_chain) {
                bool ok5 = build_band_host(p->G, p->X, p->Z, N, nb, perm, inv, band) &&
                           qme_band_geometry(N, p->E, band.bw, band.noff, S, p->smem_optin, &p->band_C, &p->band_R, &p->band_smem);
                if (!ok5) {
                    LB_REQUIRE(p->path_req != 5, "band path does not fit this problem (N=%d)", N);
                    path = 4;
                }
            }
            if (path == 4) {
                int wX[QME_MAXS], wZ[QME_MAXS];
                for (int s = 0; s < S; ++s) { wX[s] = max_row_nnz(p->X[s], N); wZ[s] = max_row_nnz(p->Z[s], N); }
                int C, R; size_t sm;
                fits = qme_cluster_geometry(N, S, wg, wX, wZ, p->E, bw_best, p->smem_optin, &C, &R, &sm);
            }
        }
        if (!fits) {
            LB_REQUIRE(p->path_req != 4 && p->path_req != 5, "sparse cluster path does not fit this problem (N=%d)", N);
            path = 3;
            for (int i = 0; i < N; ++i) perm[i] = inv[i] = i;
            p->permuted = false;
        }
    }
    if (path == 1) LB_REQUIRE(N <= 64, "dense on-chip path needs N <= 64 (N=%d)", N);
    if (path >= 3) {
        LB_REQUIRE(nd == 0, "sparse paths do not support drive operators");
        LB_REQUIRE(!p->Gr.given, "sparse paths need the right generator to be G^H");
        LB_REQUIRE(S <= QME_MAXS, "sparse paths support at most %d sandwich terms", QME_MAXS);
    }
    if (path == 1 || path == 2) LB_REQUIRE(dense_mem_ok, "dense operator batch too large (nb=%d, N=%d)", nb, N);
    p->path = path;
    p->host_perm = perm;
    if (path == 6) {
        p->host_perm = tile.order;
        p->tile_NP = tile.NP; p->tile_P = tile.P; p->tile_chunk = tile.chunk; p->tile_C = tile.C;
        p->tile_smem = std::max(tile.smem, (size_t)120 * 1024);     // > half of the SM's shared memory: one CTA (and its tensor memory) per SM
    }
    if (p->device < 0) {             // analysis-only plan: nothing to upload
        if (path >= 3) {
            int bw = 0;
            EllHost e;
            to_ell(p->G, N, e, bw, perm, inv);
            for (int s2 = 0; s2 < S; ++s2) { to_ell(p->X[s2], N, e, bw, perm, inv); to_ell(p->Z[s2], N, e, bw, perm, inv); }
            p->bandwidth = bw;
        }
        if (path == 5) { p->band_noff = band.noff; p->band_gt = band.gt; p->band_xt = band.xt; }
        p->finalized = true;
        return LB_OK;
    }

    if (path == 1 || path == 2) {
        std::vector<hcplx> g, gh, x, zh;
        densify(p->G, N, g);
        if (p->G.nb < nb) replicate(g, NN, nb);
        if (p->Gr.given) {
            LB_REQUIRE(p->Gr.nb == 1 || p->Gr.nb == nb, "right generator batch incompatible with %d", nb);
            densify(p->Gr, N, gh);
            if (p->Gr.nb < nb) replicate(gh, NN, nb);
        } else {
            gh.resize(g.size());
            for (int b = 0; b < nb; ++b) adjoint(&g[b * NN], &gh[b * NN], N);
        }
        x.resize((size_t)nb * S * NN);
        zh.resize((size_t)nb * S * NN);
        for (int s = 0; s < S; ++s) {
            std::vector<hcplx> xs, zs;
            densify(p->X[s], N, xs);
            densify(p->Z[s], N, zs);
            if (p->X[s].nb < nb) replicate(xs, NN, nb);
            if (p->Z[s].nb < nb) replicate(zs, NN, nb);
            for (int b = 0; b < nb; ++b) {
                std::copy(xs.begin() + b * NN, xs.begin() + (b + 1) * NN, x.begin() + ((size_t)b * S + s) * NN);
                adjoint(&zs[b * NN], &zh[((size_t)b * S + s) * NN], N);
            }
        }
        LB_CUDA(p->dG.upload(g.data(), g.size() * 16));
        LB_CUDA(p->dGh.upload(gh.data(), gh.size() * 16));
        LB_CUDA(p->dX.upload(x.data(), x.size() * 16));
        LB_CUDA(p->dZh.upload(zh.data(), zh.size() * 16));
        if (nd > 0) {
            LB_REQUIRE(nb == 1, "drive operators need a shared (nb = 1) generator");
            std::vector<hcplx> d((size_t)nd * NN), dh((size_t)nd * NN);
            for (int i = 0; i < nd; ++i) {
                std::copy(p->D[i].begin(), p->D[i].end(), d.begin() + i * NN);
                std::copy(p->Dr[i].begin(), p->Dr[i].end(), dh.begin() + i * NN);
            }
            LB_CUDA(p->dD.upload(d.data(), d.size() * 16));
            LB_CUDA(p->dDh.upload(dh.data(), dh.size() * 16));
        }
        std::vector<hcplx> et((size_t)p->E * NN);
        for (int e = 0; e < p->E; ++e)
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) et[e * NN + (size_t)i * N + j] = p->eops[e * NN + (size_t)j * N + i];
        LB_CUDA(p->deT.upload(et.data(), et.size() * 16));
    } else {
        int bw = 0;
        auto up = [&](const HostOp& op, DevBuf& dcol, DevBuf& dval, int& w) -> cudaError_t {
            EllHost e;
            to_ell(op, N, e, bw, perm, inv);
            if (op.nb < nb) replicate(e.val, (size_t)N * e.w, nb);
            w = e.w;
            cudaError_t r = dcol.upload(e.col.data(), e.col.size() * sizeof(int));
            if (r != cudaSuccess) return r;
            return dval.upload(e.val.data(), e.val.size() * 16);
        };
        LB_CUDA(up(p->G, p->dGcol, p->dGval, p->wG));
        for (int s = 0; s < S; ++s) {
            LB_CUDA(up(p->X[s], p->dXcol[s], p->dXval[s], p->wX[s]));
            LB_CUDA(up(p->Z[s], p->dZcol[s], p->dZval[s], p->wZ[s]));
        }
        p->bandwidth = bw;
        if (path == 5) {
            p->band_noff = band.noff; p->band_gt = band.gt; p->band_xt = band.xt;
            LB_CUDA(p->dbgd.upload(band.gd.data(), band.gd.size() * 16));
            LB_CUDA(p->dbgcol.upload(band.gcol.data(), band.gcol.size() * sizeof(int)));
            LB_CUDA(p->dbgval.upload(band.gval.data(), band.gval.size() * 16));
            for (int s = 0; s < S; ++s) {
                LB_CUDA(p->dbxcol[s].upload(band.xcol[s].data(), band.xcol[s].size() * sizeof(int)));
                LB_CUDA(p->dbxval[s].upload(band.xval[s].data(), band.xval[s].size() * 16));
                LB_CUDA(p->dbzcol[s].upload(band.zcol[s].data(), band.zcol[s].size() * sizeof(int)));
                LB_CUDA(p->dbzval[s].upload(band.zval[s].data(), band.zval[s].size() * 16));
            }
        }
        if (path == 6) {
            LB_CUDA(p->dtbrow.upload(tile.brow.data(), tile.brow.size() * sizeof(int)));
            LB_CUDA(p->dtcolold.upload(tile.colold.data(), tile.colold.size() * sizeof(int)));
            LB_CUDA(p->dtcolpos.upload(tile.colpos.data(), tile.colpos.size() * sizeof(int)));
            LB_CUDA(p->dtcolc.upload(tile.colc.data(), tile.colc.size() * sizeof(double)));
            LB_CUDA(p->dtrowc.upload(tile.rowc.data(), tile.rowc.size() * sizeof(double)));
            LB_CUDA(p->dtxs0.upload(tile.xs0.data(), tile.xs0.size() * sizeof(int)));
            LB_CUDA(p->dteptr.upload(tile.eptr.data(), tile.eptr.size() * sizeof(int)));
            LB_CUDA(p->dterank.upload(tile.erank.data(), tile.erank.size() * sizeof(int)));
            LB_CUDA(p->dteoff.upload(tile.eoff.data(), tile.eoff.size() * sizeof(int)));
            LB_CUDA(p->dteval.upload(tile.eval.data(), tile.eval.size() * 16));
        }
        if (p->permuted) LB_CUDA(p->dperm.upload(perm.data(), N * sizeof(int)));
        // observables as COO over rho's (permuted) linear index:
        // Tr(e rho) = sum_{ij} e[j][i] rho[i][j] = sum_{i'j'} e[perm j'][perm i'] rho'[i'][j']
        std::vector<int> eptr(p->E + 1, 0), eidx;
        std::vector<hcplx> eval;
        for (int e = 0; e < p->E; ++e) {
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) {
                    hcplx v = p->eops[e * NN + (size_t)perm[j] * N + perm[i]];
                    if (v != hcplx(0, 0)) { eidx.push_back(i * N + j); eval.push_back(v); }
                }
            eptr[e + 1] = (int)eidx.size();
        }
        if (eidx.empty()) { eidx.push_back(0); eval.push_back(hcplx(0, 0)); }
        LB_CUDA(p->deptr.upload(eptr.data(), eptr.size() * sizeof(int)));
        LB_CUDA(p->deidx.upload(eidx.data(), eidx.size() * sizeof(int)));
        LB_CUDA(p->deval.upload(eval.data(), eval.size() * 16));
    }
    p->finalized = true;
    return LB_OK;
}

}  // extern "C"

namespace {

__global__ void qme_build_gk(const cplx* G, const cplx* Gh, const cplx* D, const cplx* Dh,
                             const cplx* coef, int nd, int NN, int conj_r, cplx* Gk, cplx* Gkh) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NN) return;
    cplx v = G[i], vh = Gh[i];
    for (int d = 0; d < nd; ++d) {
        cplx c = coef[d];
        cfma(v, c, D[(size_t)d * NN + i]);
        cfma(vh, conj_r ? cconj(c) : c, Dh[(size_t)d * NN + i]);
    }
    Gk[i] = v; Gkh[i] = vh;
}

// dst[b][i'][j'] = src[b][perm i'][perm j'] (gather != 0) or dst[b][perm i'][perm j'] = src[b][i'][j']
__global__ void qme_permute(const cplx* __restrict__ src, cplx* __restrict__ dst, const int* __restrict__ perm,
                            int N, int gather) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    const int i = idx / N, j = idx - i * N;
    const size_t base = (size_t)blockIdx.y * N * N;
    const size_t o = (size_t)perm[i] * N + perm[j];
    if (gather) dst[base + idx] = src[base + o];
    else dst[base + o] = src[base + idx];
}

void fill_ell_args(limeb200_qme_t p, QmeEllArgs& a, int B) {
    a.N = p->N; a.S = (int)p->X.size(); a.E = p->E; a.B = B; a.nb = p->nb; a.step_vals = p->step_vals ? 1 : 0;
    a.G.w = p->wG; a.G.col = p->dGcol.as<int>(); a.G.val = p->dGval.as<cplx>();
    for (int s = 0; s < a.S; ++s) {
        a.X[s].w = p->wX[s]; a.X[s].col = p->dXcol[s].as<int>(); a.X[s].val = p->dXval[s].as<cplx>();
        a.Z[s].w = p->wZ[s]; a.Z[s].col = p->dZcol[s].as<int>(); a.Z[s].val = p->dZval[s].as<cplx>();
    }
    a.eptr = p->deptr.as<int>(); a.eidx = p->deidx.as<int>(); a.eval = p->deval.as<cplx>();
}

// N >= 96: FP64 tensor-core (DMMA) kernel, or the 64x64 register-tiled DFMA kernel when
// LIMEB200_DENSE_NO_DMMA is set; 32x32 tiles below
void launch_dense_stage(const QmeStageArgs& a, cudaStream_t st) {
    const bool no_dmma = getenv("LIMEB200_DENSE_NO_DMMA") != nullptr;     // read per launch so that tests can switch it
    if (a.N >= 96 && !no_dmma) {
        dim3 grid(ceil_div(a.N, 64), ceil_div(a.N, 64), a.B);
        qme_dense_stage_dmma<<<grid, 128, 0, st>>>(a);
    } else if (a.N >= 96) {
        dim3 grid(ceil_div(a.N, 64), ceil_div(a.N, 64), a.B);
        qme_dense_stage64<<<grid, 256, 0, st>>>(a);
    } else {
        dim3 grid(ceil_div(a.N, 32), ceil_div(a.N, 32), a.B);
        qme_dense_stage<<<grid, 256, 0, st>>>(a);
    }
}

int run_dense_onchip(limeb200_qme_t p, cplx* rho, int B, double dt, int nsteps, const cplx* coef,
                     cplx* obs, cplx* traj, int traj_every, cudaStream_t st, bool* handled) {
    const int N = p->N, NN = N * N, S = (int)p->X.size(), nd = (int)p->D.size();
    *handled = false;
    QmeDenseArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.S = S; a.E = p->E; a.nd = nd; a.B = B; a.nsteps = nsteps;
    a.traj_every = traj_every > 0 ? traj_every : 1;
    a.nb = p->nb;
    a.G = p->dG.as<cplx>(); a.Gh = p->dGh.as<cplx>(); a.X = p->dX.as<cplx>(); a.Zh = p->dZh.as<cplx>();
    a.D = p->dD.as<cplx>(); a.Dh = p->dDh.as<cplx>(); a.eT = p->deT.as<cplx>();
    a.coef = coef; a.rho = rho; a.obs = obs; a.traj = traj; a.dt = dt;
    a.drive_conj = p->drive_conj ? 1 : 0;
    int ept = 1, tps;
    if (NN <= 4) tps = 4;
    else if (NN <= 16) tps = 16;
    else {
        if (NN > 2048) ept = 4; else if (NN > 1024) ept = 2;
        tps = ceil_div(ceil_div(NN, ept), 32) * 32;
    }
    int slots = std::max(1, 256 / tps);
    slots = std::min(slots, std::max(1, B));
    // fill the machine: prefer >= 2 CTAs per SM worth of blocks when the batch is small
    while (slots > 1 && ceil_div(B, slots) < 2 * p->sm_count) slots >>= 1;
    const size_t ops_bytes = (size_t)(2 + 2 * S) * NN * 16;
    auto smem_for = [&](int sl, bool ops) { return (size_t)sl * (2 * NN + 32) * 16 + (ops ? ops_bytes : 0); };
    bool ops = (p->nb == 1) && smem_for(slots, true) <= (size_t)p->smem_optin;
    if (nd > 0 && !ops) return LB_OK;            // caller falls back to per-step launches
    if (smem_for(slots, ops) > (size_t)p->smem_optin) return LB_OK;
    a.slots = slots; a.tps = tps; a.ops_in_smem = ops ? 1 : 0;
    const size_t smem = smem_for(slots, ops);
    dim3 grid(ceil_div(B, slots)), block(slots * tps);
    void (*kern)(QmeDenseArgs) = ept == 1 ? qme_dense_onchip<1> : (ept == 2 ? qme_dense_onchip<2> : qme_dense_onchip<4>);
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(a);
    LB_CUDA(cudaGetLastError());
    p->launches += 1;
    *handled = true;
    return LB_OK;
}

int ensure_scratch(limeb200_qme_t p, int B, int ny, bool need_tmp) {
    const size_t NN = (size_t)p->N * p->N;
    const int S = (int)p->X.size();
    if (B > p->scratch_B || !p->s_y.p) {
        LB_CUDA(p->s_y.alloc((size_t)ny * B * NN * 16));
        LB_CUDA(p->s_acc.alloc((size_t)B * NN * 16));
        if (need_tmp && S > 0) LB_CUDA(p->s_tmp.alloc((size_t)S * B * NN * 16));
        p->scratch_B = B;
    }
    return LB_OK;
}

int run_dense_stage(limeb200_qme_t p, cplx* rho, int B, double dt, int nsteps, const cplx* coef,
                    cplx* obs, cplx* traj, int traj_every, cudaStream_t st) {
    const int N = p->N, S = (int)p->X.size(), nd = (int)p->D.size();
    const size_t NN = (size_t)N * N;
    int r = ensure_scratch(p, B, 2, true);
    if (r != LB_OK) return r;
    if (nd > 0 && !p->s_gk.p) LB_CUDA(p->s_gk.alloc(2 * NN * 16));
    cplx* y[2] = {p->s_y.as<cplx>(), p->s_y.as<cplx>() + (size_t)B * NN};
    cplx* acc = p->s_acc.as<cplx>();
    cplx* tmp = p->s_tmp.as<cplx>();
    const long long sOp = (p->nb > 1) ? (long long)NN : 0;
    const cplx* G = p->dG.as<cplx>();
    const cplx* Gh = p->dGh.as<cplx>();
    dim3 grid(ceil_div(N, 32), ceil_div(N, 32), B);
    for (int step = 0; step < nsteps; ++step) {
        if (nd > 0) {
            cplx* gk = p->s_gk.as<cplx>();
            qme_build_gk<<<ceil_div((int)NN, 256), 256, 0, st>>>(p->dG.as<cplx>(), p->dGh.as<cplx>(), p->dD.as<cplx>(),
                                                                p->dDh.as<cplx>(), coef + (size_t)step * nd, nd,
                                                                (int)NN, p->drive_conj ? 1 : 0, gk, gk + NN);
            p->launches++;
            G = gk; Gh = gk + NN;
        }
        for (int stage = 0; stage < 4; ++stage) {
            const cplx* yin = (stage == 0) ? rho : y[(stage - 1) & 1];
            cplx* yout = y[stage & 1];
            for (int s = 0; s < S; ++s) {           // tmp_s = yin Z_s^H
                QmeStageArgs a;
                memset(&a, 0, sizeof(a));
                a.N = N; a.B = B; a.nprod = 1;
                a.A[0] = yin; a.sA[0] = (long long)NN;
                a.Bm[0] = p->dZh.as<cplx>() + (size_t)s * NN; a.sB[0] = (p->nb > 1) ? (long long)S * NN : 0;
                a.mode = 0; a.out = tmp + (size_t)s * B * NN; a.sOut = (long long)NN; a.dt = dt;
                launch_dense_stage(a, st);
                p->launches++;
            }
            QmeStageArgs a;
            memset(&a, 0, sizeof(a));
            a.N = N; a.B = B; a.nprod = 2 + S;
            LB_REQUIRE(a.nprod <= QME_MAXPROD, "too many sandwich terms for the dense stage kernel");
            a.A[0] = G; a.sA[0] = sOp; a.Bm[0] = yin; a.sB[0] = (long long)NN;
            a.A[1] = yin; a.sA[1] = (long long)NN; a.Bm[1] = Gh; a.sB[1] = sOp;
            for (int s = 0; s < S; ++s) {
                a.A[2 + s] = p->dX.as<cplx>() + (size_t)s * NN; a.sA[2 + s] = (p->nb > 1) ? (long long)S * NN : 0;
                a.Bm[2 + s] = tmp + (size_t)s * B * NN; a.sB[2 + s] = (long long)NN;
            }
            a.mode = stage + 1; a.rho = rho; a.acc = acc; a.ynext = yout; a.dt = dt;
            launch_dense_stage(a, st);
            p->launches++;
        }
        if (obs && p->E > 0) {
            qme_trace_obs<<<dim3(p->E, B), 256, 0, st>>>(p->deT.as<cplx>(), rho, obs + (size_t)step * B * p->E,
                                                         (int)NN, p->E, (long long)p->E);
            p->launches++;
        }
        if (traj && ((step + 1) % traj_every) == 0) {
            LB_CUDA(cudaMemcpyAsync(traj + (size_t)(step / traj_every) * B * NN, rho, (size_t)B * NN * 16,
                                    cudaMemcpyDeviceToDevice, st));
        }
    }
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

}  // namespace

extern "C" {

int limeb200_qme_run(limeb200_qme_t p, double* d_rho, int B, double dt, int nsteps,
                     const double* d_coef, double* d_obs, double* d_traj, int traj_every,
                     void* stream) {
    LB_REQUIRE(p && p->finalized, "plan not finalized");
    LB_REQUIRE(p->device >= 0, "analysis-only plan (device -1) cannot run: there is no CPU fallback");
    LB_REQUIRE(d_rho && B >= 1 && nsteps >= 0, "bad arguments");
    LB_REQUIRE(p->step_vals || p->nb == 1 || p->nb == B, "operator batch %d != B %d", p->nb, B);
    LB_REQUIRE(!p->step_vals || nsteps <= p->nb, "per-step operator values cover %d steps, %d requested", p->nb, nsteps);
    LB_REQUIRE(p->D.empty() || d_coef, "drive operators present but d_coef is NULL");
    LB_REQUIRE(!d_traj || traj_every >= 1, "traj_every must be >= 1");
    LB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    p->launches = 0;
    if (nsteps == 0) return LB_OK;
    cplx* rho = (cplx*)d_rho;
    const cplx* coef = (const cplx*)d_coef;
    cplx* obs = (p->E > 0) ? (cplx*)d_obs : nullptr;
    cplx* traj = (cplx*)d_traj;
    if (!d_traj) traj_every = 1;
    if (p->path == 1) {
        bool handled = false;
        int r = run_dense_onchip(p, rho, B, dt, nsteps, coef, obs, traj, traj_every, st, &handled);
        if (r != LB_OK) return r;
        if (handled) return LB_OK;
        return run_dense_stage(p, rho, B, dt, nsteps, coef, obs, traj, traj_every, st);
    }
    if (p->path == 2) return run_dense_stage(p, rho, B, dt, nsteps, coef, obs, traj, traj_every, st);
    QmeEllArgs a;
    memset(&a, 0, sizeof(a));
    fill_ell_args(p, a, B);
    a.nsteps = nsteps; a.traj_every = traj_every; a.rho = rho; a.obs = obs; a.traj = traj; a.dt = dt;
    if (p->path == 6) {
        QmeTileArgs t;
        memset(&t, 0, sizeof(t));
        t.N = p->N; t.E = p->E; t.B = B; t.nsteps = nsteps; t.traj_every = traj_every; t.nb = p->nb;
        t.C = p->tile_C; t.P = p->tile_P; t.chunk = p->tile_chunk;
        t.brow = p->dtbrow.as<int>(); t.colold = p->dtcolold.as<int>(); t.colpos = p->dtcolpos.as<int>();
        t.colc = p->dtcolc.as<double>(); t.rowc = p->dtrowc.as<double>(); t.xs0 = p->dtxs0.as<int>();
        t.eptr = p->dteptr.as<int>(); t.erank = p->dterank.as<int>(); t.eoff = p->dteoff.as<int>();
        t.eval = p->dteval.as<cplx>();
        t.rho = rho; t.obs = obs; t.traj = traj; t.dt = dt;
        int r = qme_tile_launch(t, p->tile_NP, (int)p->X.size(), p->tile_smem, st);
        if (r != LB_OK) return r;
        p->launches += 1;
        return LB_OK;
    }
    if (p->path == 5) {
        QmeBandArgs g;
        memset(&g, 0, sizeof(g));
        g.N = p->N; g.E = p->E; g.B = B; g.nsteps = nsteps; g.traj_every = traj_every; g.nb = p->nb;
        g.R = p->band_R; g.h = p->bandwidth; g.C = p->band_C;
        g.gd = p->dbgd.as<cplx>(); g.gcol = p->dbgcol.as<int>(); g.gval = p->dbgval.as<cplx>();
        const int S = (int)p->X.size();
        for (int s = 0; s < S; ++s) {
            g.xcol[s] = p->dbxcol[s].as<int>(); g.xval[s] = p->dbxval[s].as<cplx>();
            g.zcol[s] = p->dbzcol[s].as<int>(); g.zval[s] = p->dbzval[s].as<cplx>();
        }
        g.perm = p->permuted ? p->dperm.as<int>() : nullptr;
        g.eptr = p->deptr.as<int>(); g.eidx = p->deidx.as<int>(); g.eval = p->deval.as<cplx>();
        g.rho = rho; g.obs = obs; g.traj = traj; g.dt = dt;
        { const char* dbg = getenv("LIMEB200_DEBUG_FLAGS"); g.debug_flags = dbg ? atoi(dbg) : 0; }
        g.chain_rs = p->band_chain;
        int r;
        if (p->band_chain) r = qme_band_launch_chain_tc2(g, S, p->band_smem, st);
        else
        r = p->band_noff == 2 ? qme_band_launch_tc2_n2(g, S, p->band_gt, p->band_xt, p->band_smem, st)
                              : qme_band_launch_tc2_n4(g, S, p->band_gt, p->band_xt, p->band_smem, st);
        if (r != LB_OK) return r;
        p->launches += 1;
        return LB_OK;
    }
    if (p->path == 4) {
        int r = qme_cluster_launch(a, p->bandwidth, p->permuted ? p->dperm.as<int>() : nullptr, p->smem_optin, st);
        if (r == LB_OK) { p->launches += 1; return LB_OK; }
        if (r != LB_ERR_UNSUPPORTED) return r;
        // geometry does not fit the cluster kernel: use the global-scratch kernel
    }
    int r = ensure_scratch(p, B, 2, false);
    if (r != LB_OK) return r;
    a.ybuf = p->s_y.as<cplx>(); a.accbuf = p->s_acc.as<cplx>();
    const int NN = p->N * p->N;
    int threads = std::min(1024, ceil_div(NN, 32) * 32);
    qme_ell_global<<<B, threads, 0, st>>>(a);
    LB_CUDA(cudaGetLastError());
    p->launches += 1;
    return LB_OK;
}

int limeb200_qme_rhs(limeb200_qme_t p, const double* d_in, double* d_out, int B, void* stream) {
    LB_REQUIRE(p && p->finalized, "plan not finalized");
    LB_REQUIRE(p->device >= 0, "analysis-only plan (device -1) cannot run: there is no CPU fallback");
    LB_REQUIRE(d_in && d_out && B >= 1, "bad arguments");
    LB_REQUIRE(p->nb == 1 || p->nb == B, "operator batch %d != B %d", p->nb, B);
    LB_REQUIRE(p->D.empty(), "rhs with drive operators is not supported");
    LB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int N = p->N, S = (int)p->X.size();
    const size_t NN = (size_t)N * N;
    p->launches = 0;
    if (p->path == 1 || p->path == 2) {
        int r = ensure_scratch(p, B, 2, true);
        if (r != LB_OK) return r;
        cplx* tmp = p->s_tmp.as<cplx>();
        const cplx* yin = (const cplx*)d_in;
        dim3 grid(ceil_div(N, 32), ceil_div(N, 32), B);
        const long long sOp = (p->nb > 1) ? (long long)NN : 0;
        for (int s = 0; s < S; ++s) {
            QmeStageArgs a;
            memset(&a, 0, sizeof(a));
            a.N = N; a.B = B; a.nprod = 1;
            a.A[0] = yin; a.sA[0] = (long long)NN;
            a.Bm[0] = p->dZh.as<cplx>() + (size_t)s * NN; a.sB[0] = (p->nb > 1) ? (long long)S * NN : 0;
            a.mode = 0; a.out = tmp + (size_t)s * B * NN; a.sOut = (long long)NN;
            launch_dense_stage(a, st);
            p->launches++;
        }
        QmeStageArgs a;
        memset(&a, 0, sizeof(a));
        a.N = N; a.B = B; a.nprod = 2 + S;
        a.A[0] = p->dG.as<cplx>(); a.sA[0] = sOp; a.Bm[0] = yin; a.sB[0] = (long long)NN;
        a.A[1] = yin; a.sA[1] = (long long)NN; a.Bm[1] = p->dGh.as<cplx>(); a.sB[1] = sOp;
        for (int s = 0; s < S; ++s) {
            a.A[2 + s] = p->dX.as<cplx>() + (size_t)s * NN; a.sA[2 + s] = (p->nb > 1) ? (long long)S * NN : 0;
            a.Bm[2 + s] = tmp + (size_t)s * B * NN; a.sB[2 + s] = (long long)NN;
        }
        a.mode = 0; a.out = (cplx*)d_out; a.sOut = (long long)NN;
        launch_dense_stage(a, st);
        p->launches++;
    } else {
        QmeEllArgs a;
        memset(&a, 0, sizeof(a));
        fill_ell_args(p, a, B);
        dim3 grid(ceil_div((int)NN, 256), B);
        if (p->permuted) {          // operators live in the permuted basis
            int r = ensure_scratch(p, B, 2, false);
            if (r != LB_OK) return r;
            cplx* t0 = p->s_y.as<cplx>();
            cplx* t1 = t0 + (size_t)B * NN;
            qme_permute<<<grid, 256, 0, st>>>((const cplx*)d_in, t0, p->dperm.as<int>(), N, 1);
            qme_ell_rhs<<<grid, 256, 0, st>>>(a, t0, t1);
            qme_permute<<<grid, 256, 0, st>>>(t1, (cplx*)d_out, p->dperm.as<int>(), N, 0);
            p->launches += 3;
        } else {
            qme_ell_rhs<<<grid, 256, 0, st>>>(a, (const cplx*)d_in, (cplx*)d_out);
            p->launches++;
        }
    }
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

/* plain batched complex GEMM on the FP64 tensor cores: C[b] = A[b] * B[b], row-major, A [M][K], B [K][N], C [M][N];
 * strides in complex elements (0 = operand shared by the batch).  Used by the Liouvillian eigen-decomposition
 * correlation functions (lime/superoperator.py:703-754: tmp1.T @ coeff @ tmp2). */
int limeb200_zgemm(const double* d_A, const double* d_B, double* d_C, int M, int N, int K, int batch,
                   long long sA, long long sB, long long sC, void* stream) {
    LB_REQUIRE(d_A && d_B && d_C, "null argument");
    LB_REQUIRE(M >= 1 && N >= 1 && K >= 1 && batch >= 1 && batch <= 65535, "bad sizes");
    QmeStageArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.B = batch; a.nprod = 1; a.Mr = M; a.Nc = N; a.Kd = K;
    a.A[0] = (const cplx*)d_A; a.sA[0] = sA;
    a.Bm[0] = (const cplx*)d_B; a.sB[0] = sB;
    a.mode = 0; a.out = (cplx*)d_C; a.sOut = sC;
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64), batch);
    qme_dense_stage_dmma<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

}  // extern "C"
