"""
Sum-over-states nonlinear response functions with lime's signatures (lime/signal/sos.py),
evaluated on the device in factorised form.

Each pathway of sos.py is a sum over state triples (b,c,d) of four dipole elements times two
frequency-domain Green's functions G(w) = 1/(w - dE + i Gamma) (one per scanned axis) and one
propagator at the fixed delay.  Grouping by the index the first-axis factor depends on,

    S[r][c] = sum_q A[q][r] * B[q][c],      B[q][c] = sum_d W[q][d] G(w_c; pole_{q,d})

so the O(states^3) weights W are assembled on the host (numpy, microseconds) and the O(grid)
work -- factor tables and the rank-R outer product, one complex store per grid point -- is two
CUDA kernels (limeb200_sos_factor / limeb200_sos_outer).  Results differ from lime's only by
floating-point re-association (~1e-15 relative).

Conventions kept from lime: np.meshgrid(omega1, omega3) uses 'xy' indexing, so the returned
array is indexed [omega3, omega1] (lime/signal/sos.py:379-388); DQC_R1 evaluates G_ba at the
second axis (lime/signal/sos.py:949).
[ext]: the *_batch variants take a vector of delays and return [T, n3, n1] in one launch.
"""
import numpy as np
import torch

from .. import engine
from .. import _dev
from ..units import au2mev
from ..phys import lorentzian  # noqa: F401  (re-exported like lime)


def _z(w):
    return _dev.to_dev(np.asarray(w, dtype=np.float64).reshape(-1), np.float64)


def _simple_factor(z, e, g):
    """A[q][n] = 1/(z_n - e_q + i g_q) -> device [1,R,n]"""
    e = np.asarray(e, dtype=float).reshape(-1)
    g = np.asarray(g, dtype=float).reshape(-1)
    R = len(e)
    W = np.ones((1, R, 1), dtype=complex)
    p = np.stack([e, g], axis=-1).reshape(R, 1, 2)
    return engine.sos_factor(z, W, p)


def _U(E, gamma, x, y, tau, extra=None):
    """-i exp(-i (E_x - E_y) tau - Gamma_xy tau) for index arrays x, y (broadcast) and delays tau[T]"""
    tau = np.asarray(tau, dtype=float).reshape(-1, *([1] * np.ndim(np.broadcast_arrays(x, y)[0])))
    dE = E[x] - E[y]
    G = (gamma[x] + gamma[y]) / 2.0
    if extra is not None:
        G = G + extra
    return -1j * np.exp(-1j * dE * tau - G * tau)


def _finish(out, single):
    r = _dev.d2h_owned(out) if out.numel() >= (1 << 16) else out.cpu().numpy()
    return r[0] if single else r


def _fp(*arrays):
    """content fingerprint of small host inputs (energies, dipoles, index lists, grids) for the plan cache"""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        b = np.ascontiguousarray(np.asarray(a if not isinstance(a, range) else list(a)))
        h.update(str((b.shape, b.dtype.str)).encode())
        h.update(b.view(np.uint8).reshape(-1))
    return h.hexdigest()


_PE_CACHE = {}
_PE_CACHE_MAX = 4


def _check_grid(n1, n3):
    # lime adds (n3,n1) arrays into zeros((n1,n3)): only broadcast-compatible shapes work
    np.broadcast_shapes((n1, n3), (n3, n1))


# ---------------------------------------------------------------------------------------
# photon echo: GSB, SE, ESA  (frequency-frequency at waiting time tau2)
# ---------------------------------------------------------------------------------------
def _pe_terms(evals, dip, tau2, g_idx, e_idx, f_idx, gamma, parts):
    """weights and poles of the requested pathways, grouped by b (the omega1 index).
    returns W[T,R,D], poles[R,D,2], (eA[R], gA[R])"""
    E = np.asarray(evals, dtype=float)
    dip = np.asarray(dip)
    gamma = np.asarray(gamma, dtype=float)
    a = 0
    eb = np.array(list(e_idx), dtype=int)
    taus = np.atleast_1d(np.asarray(tau2, dtype=float))
    T, R = len(taus), len(eb)
    Ws, Ps = [], []
    if 'GSB' in parts:
        c = 0
        d = eb
        w = (dip[a, eb][:, None] * dip[eb, c][:, None]) * (dip[c, d] * dip[d, a])[None, :]       # [b,d]
        Ws.append(np.broadcast_to(w[None].astype(complex), (T, R, len(d))))
        pe = np.broadcast_to((E[d] - E[c])[None, :], (R, len(d)))
        pg = np.broadcast_to(((gamma[d] + gamma[c]) / 2.0)[None, :], (R, len(d)))
        Ps.append(np.stack([pe, pg], axis=-1))
    if 'SE' in parts:
        cc = eb
        dd = np.array(list(g_idx), dtype=int)
        U = _U(E, gamma, cc[None, :], eb[:, None], taus)                                        # [T,b,c]
        w = (dip[a, eb][:, None, None] * dip[cc, a][None, :, None]
             * dip[dd[None, :], cc[:, None]][None, :, :] * dip[eb[:, None], dd[None, :]][:, None, :])  # [b,c,d]
        Ws.append((U[:, :, :, None] * w[None]).reshape(T, R, -1))
        pe = np.broadcast_to((E[cc][:, None] - E[dd][None, :])[None], (R, len(cc), len(dd))).reshape(R, -1)
        pg = np.broadcast_to(((gamma[cc][:, None] + gamma[dd][None, :]) / 2.0)[None], (R, len(cc), len(dd))).reshape(R, -1)
        Ps.append(np.stack([pe, pg], axis=-1))
    if 'ESA' in parts:
        cc = eb
        dd = np.array(list(f_idx), dtype=int)
        U = _U(E, gamma, cc[None, :], eb[:, None], taus)                                        # [T,b,c]
        m = dip[cc, a][:, None] * dip[dd[None, :], cc[:, None]]                                  # [c,d]
        inner = np.einsum('tbc,cd->tbd', U, m)
        w = -(dip[eb, a][:, None] * dip[eb[:, None], dd[None, :]])[None] * inner                # [T,b,d]
        Ws.append(w)
        pe = E[dd][None, :] - E[eb][:, None]
        pg = (gamma[dd][None, :] + gamma[eb][:, None]) / 2.0
        Ps.append(np.stack([pe, pg], axis=-1))
    W = np.concatenate(Ws, axis=2)
    P = np.concatenate(Ps, axis=1)
    eA = E[a] - E[eb]
    gA = (gamma[a] + gamma[eb]) / 2.0
    return W, P, (eA, gA)


def _pe_eval(evals, dip, omega1, omega3, tau2, g_idx, e_idx, f_idx, gamma, parts, return_device=False):
    single = np.ndim(tau2) == 0
    n1, n3 = len(omega1), len(omega3)
    _check_grid(n1, n3)
    if len(list(e_idx)) == 0:
        T = 1 if single else len(tau2)
        z = np.zeros((T, n3, n1), dtype=complex)
        return z[0] if single else z
    # the O(states^3) weights, poles and grids are a pure function of the inputs: the uploaded plan (PhotonEchoGrid) is
    # cached on their CONTENT, so that repeated evaluations on one system (scans over anything else, re-plots) cost the
    # O(grid) device work and the transfer of the result only
    key = (_fp(evals, dip, gamma, list(g_idx), list(e_idx), list(f_idx), omega1, omega3, tau2), tuple(parts),
           torch.cuda.current_device() if torch.cuda.is_available() else None)
    grid = _PE_CACHE.pop(key, None)
    if grid is None:
        grid = PhotonEchoGrid(evals, dip, omega1, omega3, tau2, g_idx, e_idx, f_idx, gamma, parts=parts)
        while len(_PE_CACHE) >= _PE_CACHE_MAX:
            _PE_CACHE.pop(next(iter(_PE_CACHE)))
    _PE_CACHE[key] = grid
    out = grid.run()
    if return_device:
        return out.clone()                         # the plan's output buffer is overwritten by the next evaluation
    return _finish(out, single)


class PhotonEchoGrid:
    """[ext] device-resident photon-echo evaluation for repeated use on one system: the O(states^3) weights
    (host set-up, `_pe_terms`) and the grids are uploaded once; `run()` is the O(grid) part only -- two factor
    kernels and the rank-R outer product -- and returns the [T, n3, n1] signal on the device."""

    def __init__(self, evals, dip, omega1, omega3, tau2, g_idx, e_idx, f_idx, gamma, parts=('GSB', 'SE', 'ESA')):
        W, P, (eA, gA) = _pe_terms(evals, dip, np.atleast_1d(tau2), g_idx, e_idx, f_idx, gamma, parts)
        self.T, R = W.shape[0], W.shape[1]
        self.z1, self.z3 = _z(omega1), _z(omega3)
        self.dW = _dev.to_dev(np.ascontiguousarray(W))
        self.dP = _dev.to_dev(np.asarray(P, dtype=np.float64).reshape(R, W.shape[2], 2), np.float64)
        self.dWa = _dev.to_dev(np.ones((1, R, 1), dtype=complex))
        self.dPa = _dev.to_dev(np.stack([np.asarray(eA, dtype=float), np.asarray(gA, dtype=float)], axis=-1)
                               .reshape(R, 1, 2), np.float64)
        self.out = None
        self._graph = None

    def _launch(self):
        A = engine.sos_factor_dev(self.z1, self.dWa, self.dPa)       # [1,R,n1]  G_ab(omega1)
        Bf = engine.sos_factor_dev(self.z3, self.dW, self.dP)        # [T,R,n3]
        self.out = engine.sos_outer(Bf, A, self.T, out=self.out)     # [T,n3,n1]
        return A, Bf

    def run(self, use_graph=True):
        """the three launches are replayed from a CUDA graph (at 256 x 256 x 64 the kernels take ~50 us, less than three
        individual launches from Python); the graph is captured in the FIRST call, right after a plain warm-up pass, so
        that every later call is a replay"""
        if not use_graph:
            self._launch()
        elif self._graph is None:
            self._keep = self._launch()                  # plain launches: warm-up (module load, allocator)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._keep = self._launch()
            self._graph = g
            g.replay()
        else:
            self._graph.replay()
        return self.out


def GSB(evals, dip, omega1, omega3, tau2, g_idx, e_idx, gamma):
    """ground-state bleach; lime/signal/sos.py:478-528"""
    return _pe_eval(evals, dip, omega1, omega3, tau2, g_idx, e_idx, [], gamma, ('GSB',))


def SE(evals, dip, omega1, omega3, tau2, g_idx, e_idx, gamma):
    """stimulated emission; lime/signal/sos.py:576-635"""
    return _pe_eval(evals, dip, omega1, omega3, tau2, g_idx, e_idx, [], gamma, ('SE',))


def ESA(evals, dip, omega1, omega3, tau2, g_idx, e_idx, f_idx, gamma):
    """excited-state absorption (sign -1); lime/signal/sos.py:348-407"""
    return _pe_eval(evals, dip, omega1, omega3, tau2, g_idx, e_idx, f_idx, gamma, ('ESA',))


def _photon_echo(evals, edip, omega1, omega3, t2, g_idx, e_idx, f_idx, gamma, return_device=False):
    """GSB + SE + ESA at waiting time t2 (scalar, or [ext] a vector -> [T,n3,n1]);
    lime/signal/sos.py:695-729.  [ext] return_device=True leaves the [T,n3,n1] result on the GPU."""
    return _pe_eval(evals, edip, omega1, omega3, t2, g_idx, e_idx, f_idx, gamma, ('GSB', 'SE', 'ESA'),
                    return_device=return_device)


def photon_echo(mol, pump, probe, t2=0., g_idx=[0], e_idx=None, f_idx=None, fname='signal',
                plt_signal=False, pol=None):
    """lime/signal/sos.py:812-902 (np.savez side effect kept; plotting is not part of the port)"""
    E = mol.eigvals()
    dip = mol.edip_rms
    gamma = mol.gamma
    if gamma is None:
        raise ValueError('Please set the decay constants gamma first.')
    N = mol.nstates
    if e_idx is None:
        e_idx = range(N)
    if f_idx is None:
        f_idx = range(N)
    S = _photon_echo(E, dip, omega1=-np.asarray(pump), omega3=probe, t2=t2, g_idx=g_idx, e_idx=e_idx,
                     f_idx=f_idx, gamma=gamma)
    if fname is not None:
        np.savez(fname, pump, probe, S)
    return S


# ---- (omega1, omega2) maps at detection time t3, with pure dephasing --------------------
def _t3_eval(E, dip, omega1, omega2, t3, g_idx, e_idx, f_idx, gamma, dephasing, which):
    E = np.asarray(E, dtype=float)
    dip = np.asarray(dip)
    gamma = np.asarray(gamma, dtype=float)
    N = len(E)
    gD = np.ones((N, N)) * dephasing
    np.fill_diagonal(gD, 0)
    a = 0
    eb = np.array(list(e_idx), dtype=int)
    cc = eb
    single = np.ndim(t3) == 0
    taus = np.atleast_1d(np.asarray(t3, dtype=float))
    n1, n2 = len(omega1), len(omega2)
    if which == 'SE':
        dd = np.array(list(g_idx), dtype=int)
        Gt = _U(E, gamma, cc[:, None], dd[None, :], taus, extra=gD[cc[:, None], dd[None, :]])   # [T,c,d]
        w = (dip[a, eb][:, None, None] * dip[cc, a][None, :, None]
             * dip[dd[None, :], cc[:, None]][None] * dip[eb[:, None], dd[None, :]][:, None, :])  # [b,c,d]
        W = np.einsum('bcd,tcd->tbc', w, Gt)
    else:
        dd = np.array(list(f_idx), dtype=int)
        Gt = _U(E, gamma, dd[None, :], eb[:, None], taus, extra=gD[dd[None, :], eb[:, None]])   # [T,b,d]
        w = (dip[eb, a][:, None, None] * dip[cc, a][None, :, None]
             * dip[dd[None, :], cc[:, None]][None] * dip[eb[:, None], dd[None, :]][:, None, :])  # [b,c,d]
        W = -np.einsum('bcd,tbd->tbc', w, Gt)
    pe = E[cc][None, :] - E[eb][:, None]
    pg = (gamma[cc][None, :] + gamma[eb][:, None]) / 2.0 + gD[cc[None, :], eb[:, None]]
    P = np.stack([pe, pg], axis=-1)
    A = _simple_factor(_z(omega1), E[a] - E[eb], (gamma[a] + gamma[eb]) / 2.0 + gD[a, eb])
    Bf = engine.sos_factor(_z(omega2), W, P)
    out = engine.sos_outer(Bf, A, W.shape[0])          # [T, n2, n1]
    return _finish(out, single)


def _SE(E, dip, omega1, omega2, t3, g_idx, e_idx, gamma, dephasing=10 / au2mev):
    """lime/signal/sos.py:638-692"""
    return _t3_eval(E, dip, omega1, omega2, t3, g_idx, e_idx, [], gamma, dephasing, 'SE')


def _ESA(evals, dip, omega1, omega2, t3, g_idx, e_idx, f_idx, gamma, dephasing=10 / au2mev):
    """lime/signal/sos.py:410-476"""
    return _t3_eval(evals, dip, omega1, omega2, t3, g_idx, e_idx, f_idx, gamma, dephasing, 'ESA')


def photon_echo_t3(mol, omega1, omega2, t3, g_idx=[0], e_idx=None, f_idx=None,
                   fname='2DES', plt_signal=False, separate=False):
    """lime/signal/sos.py:731-810"""
    E = mol.eigvals()
    edip = mol.edip_rms
    gamma = mol.gamma
    dephasing = mol.dephasing
    if gamma is None:
        raise ValueError('Please set the decay constants gamma first.')
    N = mol.nstates
    if e_idx is None:
        e_idx = range(1, N)
    if f_idx is None:
        f_idx = range(1, N)
    se = _SE(E, edip, -np.asarray(omega1), omega2, t3, g_idx, e_idx, gamma, dephasing=dephasing)
    esa = _ESA(E, edip, -np.asarray(omega1), omega2, t3, g_idx, e_idx, f_idx, gamma, dephasing=dephasing)
    S = se + esa
    if separate:
        if fname is not None:
            np.savez(fname, omega1, omega2, se, esa)
        return se, esa
    if fname is not None:
        np.savez(fname, omega1, omega2, S)
    return S


# ---------------------------------------------------------------------------------------
# double-quantum coherence
# ---------------------------------------------------------------------------------------
def _ones_factor(n):
    return torch.ones((1, 1, n), dtype=torch.complex128, device=_dev.device())


def DQC_R1(evals, dip, omega1=None, omega2=[], omega3=None, tau1=None, tau3=None,
           g_idx=[0], e_idx=None, f_idx=None, gamma=None):
    """gg -> eg -> fg -> fe' -> e'e'; lime/signal/sos.py:904-992"""
    E = np.asarray(evals, dtype=float)
    dip = np.asarray(dip)
    gamma = np.asarray(gamma, dtype=float)
    a = 0
    eb = np.array(list(e_idx), dtype=int)
    fc = np.array(list(f_idx), dtype=int)
    ed = eb
    if omega3 is None and tau3 is not None:
        single = np.ndim(tau3) == 0
        taus = np.atleast_1d(np.asarray(tau3, dtype=float))
        U = _U(E, gamma, fc[:, None], ed[None, :], taus)                             # [T,c,d]
        m = np.einsum('tcd,d,dc->tc', U, dip[ed, a], dip[ed[:, None], fc[None, :]])   # sum_d
        Wbc = dip[eb, a][None, :, None] * dip[fc[None, :], eb[:, None]][None] * m[:, None, :]   # [T,b,c]
        T = len(taus)
        W = (-Wbc).reshape(T, 1, -1)
        nb, nc = len(eb), len(fc)
        p1 = np.stack([np.broadcast_to((E[eb] - E[a])[:, None], (nb, nc)),
                       np.broadcast_to(((gamma[eb] + gamma[a]) / 2.0)[:, None], (nb, nc))], axis=-1).reshape(1, -1, 2)
        p2 = np.stack([np.broadcast_to((E[fc] - E[a])[None, :], (nb, nc)),
                       np.broadcast_to(((gamma[fc] + gamma[a]) / 2.0)[None, :], (nb, nc))], axis=-1).reshape(1, -1, 2)
        Bf = engine.sos_factor(_z(omega2), W, p1, p2)          # [T,1,n2]  (both G's on omega2, :949)
        A = _ones_factor(len(omega1))
        return _finish(engine.sos_outer(A, Bf, T), single)     # [T, n1, n2]
    elif omega1 is None and tau1 is not None:
        single = np.ndim(tau1) == 0
        taus = np.atleast_1d(np.asarray(tau1, dtype=float))
        U = _U(E, gamma, eb, a, taus)                                                 # [T,b]
        s = np.einsum('tb,b,cb->tc', U, dip[eb, a], dip[fc[:, None], eb[None, :]])    # sum_b
        W = -(dip[ed, a][None, :] * dip[ed[None, :], fc[:, None]])[None] * s[:, :, None]        # [T,c,d]
        pe = E[fc][:, None] - E[ed][None, :]
        pg = (gamma[fc][:, None] + gamma[ed][None, :]) / 2.0
        Bf = engine.sos_factor(_z(omega3), W, np.stack([pe, pg], axis=-1))
        A = _simple_factor(_z(omega2), E[fc] - E[a], (gamma[fc] + gamma[a]) / 2.0)
        return _finish(engine.sos_outer(A, Bf, len(taus)), single)                    # [T, n2, n3]
    # lime falls through with `signal` undefined (UnboundLocalError); make the misuse explicit
    raise Exception('Input Error! Please specify either omega1, tau3 or omega3, tau1.')


def DQC_R2(evals, dip, omega1=None, omega2=[], omega3=None, tau1=None, tau3=None,
           g_idx=[0], e_idx=None, f_idx=None, gamma=None):
    """gg -> eg -> fg -> eg -> gg; lime/signal/sos.py:994-1099"""
    E = np.asarray(evals, dtype=float)
    dip = np.asarray(dip)
    gamma = np.asarray(gamma, dtype=float)
    a = 0
    eb = np.array(list(e_idx), dtype=int)
    fc = np.array(list(f_idx), dtype=int)
    ed = eb
    if omega3 is None and tau3 is not None:
        single = np.ndim(tau3) == 0
        taus = np.atleast_1d(np.asarray(tau3, dtype=float))
        U = _U(E, gamma, ed, a, taus)                                                 # [T,d]
        s = np.einsum('td,dc,d->tc', U, dip[ed[:, None], fc[None, :]], dip[a, ed])    # sum_d
        W = (dip[eb, a][:, None] * dip[fc[None, :], eb[:, None]])[None] * s[:, None, :]         # [T,b,c]
        pe = np.broadcast_to((E[fc] - E[a])[None, :], (len(eb), len(fc)))
        pg = np.broadcast_to(((gamma[fc] + gamma[a]) / 2.0)[None, :], (len(eb), len(fc)))
        Bf = engine.sos_factor(_z(omega2), W, np.stack([pe, pg], axis=-1))
        A = _simple_factor(_z(omega1), E[eb] - E[a], (gamma[eb] + gamma[a]) / 2.0)
        return _finish(engine.sos_outer(A, Bf, len(taus)), single)                    # [T, n1, n2]
    elif omega1 is None and tau1 is not None:
        single = np.ndim(tau1) == 0
        taus = np.atleast_1d(np.asarray(tau1, dtype=float))
        U = 1j * _U(E, gamma, eb, a, taus)              # no -i prefactor here (lime/signal/sos.py:1056)
        s = np.einsum('tb,b,cb->tc', U, dip[eb, a], dip[fc[:, None], eb[None, :]])
        W = (dip[ed[None, :], fc[:, None]] * dip[a, ed][None, :])[None] * s[:, :, None]          # [T,c,d]
        pe = np.broadcast_to((E[ed] - E[a])[None, :], (len(fc), len(ed)))
        pg = np.broadcast_to(((gamma[ed] + gamma[a]) / 2.0)[None, :], (len(fc), len(ed)))
        Bf = engine.sos_factor(_z(omega3), W, np.stack([pe, pg], axis=-1))
        A = _simple_factor(_z(omega2), E[fc] - E[a], (gamma[fc] + gamma[a]) / 2.0)
        return _finish(engine.sos_outer(A, Bf, len(taus)), single)                    # [T, n2, n3]
    raise Exception('Input Error! Please specify either omega1, tau3 or omega3, tau1.')


# ---------------------------------------------------------------------------------------
# two-photon absorption
# ---------------------------------------------------------------------------------------
def _tpa2d(E, dip, omegaps, omega1s, e_idx, f_idx, gamma, time_order):
    from .._lib import lib, check
    dev = _dev.device()
    E = np.asarray(E, dtype=float)
    N = len(E)
    dE = _dev.to_dev(E, np.float64, dev)
    dd = _dev.to_dev(np.asarray(dip, dtype=float).reshape(N, N), np.float64, dev)
    dg = _dev.to_dev(np.asarray(gamma, dtype=float), np.float64, dev)
    ei = _dev.to_dev(np.array(list(e_idx), dtype=np.int32), np.int32, dev)
    fi = _dev.to_dev(np.array(list(f_idx), dtype=np.int32), np.int32, dev)
    wp, w1 = _z(omegaps), _z(omega1s)
    out = torch.empty((wp.shape[0], w1.shape[0]), dtype=torch.float64, device=dev)
    check(lib().limeb200_sos_tpa2d(_dev.ptr(dE), _dev.ptr(dd), _dev.ptr(dg), N, _dev.ptr(ei), ei.shape[0],
                                   _dev.ptr(fi), fi.shape[0], _dev.ptr(wp), wp.shape[0], _dev.ptr(w1), w1.shape[0],
                                   1 if time_order else 0, _dev.ptr(out), _dev.stream_ptr(dev)))
    return out.cpu().numpy()


def TPA2D(E, dip, omegaps, omega1s, g_idx, e_idx, f_idx, gamma):
    """lime/signal/sos.py:230-256"""
    return _tpa2d(E, dip, omegaps, omega1s, e_idx, f_idx, gamma, False)


def TPA2D_time_order(E, dip, omegaps, omega1s, g_idx, e_idx, f_idx, gamma):
    """lime/signal/sos.py:258-283"""
    return _tpa2d(E, dip, omegaps, omega1s, e_idx, f_idx, gamma, True)


def TPA(E, dip, omegap, g_idx, e_idx, f_idx, gamma, degenerate=True):
    """lime/signal/sos.py:199-228 (degenerate pump: omega1 = omega2 = omegap/2)"""
    if not degenerate:
        raise UnboundLocalError("omega1 is not defined for degenerate=False (as in lime)")
    return float(_tpa2d(E, dip, [omegap], [omegap * 0.5], e_idx, f_idx, gamma, False)[0, 0])


# ---------------------------------------------------------------------------------------
# entangled two-photon absorption: double time integrals
# ---------------------------------------------------------------------------------------
def _etpa(omegaps, Es, edip, jta, t1, t2, g_idx, e_idx, f_idx):
    """ETPA signal from the joint temporal amplitude, lime/signal/sos.py:1171-1223:

        signal[j] = sum_{f,e} mu_eg mu_fe  sum_{T1,T2} theta(T2 - T1) e^{i d2 T2 + i d1 T1} (jta + jta^T)[T2, T1],
        d2 = E_f - E_e - w_j/2,  d1 = E_e - E_g - w_j/2,  theta(0) = 1/2

    (lime's two terms have the same detunings because omega1 = omega2 = omegap/2; they differ in jta vs jta.T).
    lime evaluates the full n2 x n1 double sum for every (pump frequency, e, f) in Python loops.  The sum factorises:
    V = (theta o (jta + jta^T)) . U1 with U1[b, (j,e)] = e^{i d1 t1_b} is ONE complex GEMM on the FP64 tensor cores
    (limeb200_zgemm), and the remaining contraction over T2 with e^{i d2 t2_a} is limeb200_etpa_reduce.
    As in lime, `g_idx` must select ONE ground state (its `Es[g]` / `edip[e, g]` only broadcast for a single index) and
    jta must be square (it is transposed)."""
    import ctypes as C
    from .._lib import lib, check
    omegaps = np.asarray(omegaps, dtype=float).reshape(-1)
    Es = np.asarray(Es, dtype=float)
    edip = np.asarray(edip)
    t1 = np.asarray(t1, dtype=float)
    t2 = np.asarray(t2, dtype=float)
    jta = np.asarray(jta)
    g = np.atleast_1d(np.asarray(g_idx, dtype=int))
    if g.size != 1:
        raise ValueError('operands could not be broadcast together: _etpa needs a single ground state (lime indexes Es[g_idx])')
    g = int(g[0])
    e_idx = np.array(list(e_idx), dtype=int)
    f_idx = np.array(list(f_idx), dtype=int)
    T1, T2 = np.meshgrid(t1, t2)
    theta = np.heaviside(T2 - T1, 0.5)
    if jta.shape != theta.shape or jta.shape != jta.T.shape:
        raise ValueError('operands could not be broadcast together with shapes %s %s' % (theta.shape, jta.shape))
    nw, ne, nf = len(omegaps), len(e_idx), len(f_idx)
    signal = np.zeros(nw, dtype=complex)
    if ne == 0 or nf == 0 or nw == 0:
        return signal
    M = np.ascontiguousarray(theta * (jta + jta.T), dtype=np.complex128)                 # [n2, n1]
    d1 = (Es[e_idx][None, :] - Es[g]) - 0.5 * omegaps[:, None]                         # [nw, ne]
    U1 = np.ascontiguousarray(np.exp(1j * t1[:, None] * d1.reshape(1, -1)))             # [n1, nw*ne]
    V = engine.zgemm(M, U1).contiguous()                                                # [n2, nw*ne] on the device
    dev = V.device
    alpha = _dev.to_dev((Es[e_idx][None, :] + 0.5 * omegaps[:, None]).reshape(-1), np.float64, dev)
    dEf = _dev.to_dev(Es[f_idx], np.float64, dev)
    dt2 = _dev.to_dev(t2, np.float64, dev)
    out = torch.empty((nw * ne, nf), dtype=torch.complex128, device=dev)
    check(lib().limeb200_etpa_reduce(_dev.ptr(V), _dev.ptr(dt2), len(t2), _dev.ptr(alpha), nw * ne, _dev.ptr(dEf), nf,
                                     _dev.ptr(out), _dev.stream_ptr(dev)))
    S = out.cpu().numpy().reshape(nw, ne, nf)
    D = edip[e_idx, g][:, None] * edip[f_idx[None, :], e_idx[:, None]]                  # mu_eg mu_fe  [ne, nf]
    return np.einsum('jef,ef->j', S, D)


def etpa(omegaps, mol, epp, g_idx, e_idx, f_idx):
    """lime/signal/sos.py:1139-1167: ETPA signal of `mol` for the biphoton state `epp` (its get_jta() supplies the
    joint temporal amplitude on the host)"""
    Es = mol.eigenenergies()
    edip = mol.edip
    t1, t2, jta = epp.get_jta()
    return _etpa(omegaps, Es, edip, jta, t1, t2, g_idx, e_idx, f_idx)
