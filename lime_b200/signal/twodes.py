"""
Time-domain third-order response functions R(t1, t2, t3) with the signatures of lime/signal/2DES.py
(`G`, `ESA`, `GSB`, `SE`, lime/signal/2DES.py:37-247), evaluated on the device in factorised form.

lime's module cannot be imported (it executes undefined names at :249-263); its functions read the
level energies and decay rates from the module globals `en` and `decay` (G, :37-60) and ignore their
own `evals` / `gamma` arguments.  Here `en` / `decay` are module globals too: when they are set they
are used, exactly as in lime; otherwise the `evals` / `gamma` arguments are.

Each pathway is  sum_b A_b(t1) * sum_{c,d} W_{b,cd}(t2) G_{cd}(t3)  -- the same rank-|e| outer
product as the frequency-domain functions of sos.py, with exponential instead of Lorentzian
factors (limeb200_sos_factor_time + limeb200_sos_outer).

Arguments t1, t3: scalar, 1-D array, or the broadcastable pair (1, n1) / (n3, 1) one would hand to
lime; the result is indexed [t3, t1] (what NumPy broadcasting gives lime).  [ext] t2 may be a 1-D
array of waiting times -> [T, n3, n1].
"""
import numpy as np

from .. import engine
from .. import _dev

en = None
decay = None


def _levels(evals, gamma):
    E = np.asarray(en if en is not None else evals, dtype=float)
    g = np.asarray(decay if decay is not None else gamma, dtype=float)
    if g.ndim == 0:
        g = np.full(E.shape, float(g))
    return E, g


def _axis(t):
    """-> (1-D float array, was_scalar)"""
    a = np.asarray(t, dtype=float)
    if a.ndim == 0:
        return a.reshape(1), True
    if a.ndim == 2 and 1 in a.shape:
        a = a.reshape(-1)
    if a.ndim != 1:
        raise ValueError('t1 / t3 must be scalars, 1-D arrays or (1,n) / (n,1) arrays')
    return np.ascontiguousarray(a), False


def _Gh(E, g, a, b, t):
    """host copy of lime's G for the (tiny) t2 weights"""
    t = np.asarray(t, dtype=float)
    return -1j * np.heaviside(t, 1) * np.exp(-1j * (E[a] - E[b]) * t - (g[a] + g[b]) / 2. * t)


def G(a, b, t, gamma=None):
    """-i theta(t) exp(-i (E_a - E_b) t - (g_a + g_b)/2 t); lime/signal/2DES.py:37-60 (globals en, decay)"""
    if en is None or decay is None:
        raise NameError("set lime_b200.signal.twodes.en and .decay (module globals, as in lime)")
    t1, scalar = _axis(t)
    F = engine.sos_factor_time(_dev.to_dev(t1, np.float64), np.ones((1, 1, 1), dtype=complex),
                               [[[en[a] - en[b], (decay[a] + decay[b]) / 2.]]])
    out = F.cpu().numpy()[0, 0]
    return out[0] if scalar else out.reshape(np.shape(t))


def _evaluate(E, g, dip, b_list, terms, t1, t2, t3, sign):
    """terms(b, taus) -> (W[T, D], poles[D, 2]) for the t3 factor of row b"""
    t1a, s1 = _axis(t1)
    t3a, s3 = _axis(t3)
    taus = np.atleast_1d(np.asarray(t2, dtype=float))
    single = np.ndim(t2) == 0
    R = len(b_list)
    if R == 0:
        z = np.zeros((len(taus), len(t3a), len(t1a)), dtype=complex)
    else:
        Ws, Ps = zip(*[terms(b, taus) for b in b_list])
        W = sign * np.stack(Ws, axis=1)                       # [T, R, D]
        P = np.stack(Ps, axis=0)                              # [R, D, 2]
        A = engine.sos_factor_time(_dev.to_dev(t1a, np.float64), np.ones((1, R, 1), dtype=complex),
                                   np.array([[[E[0] - E[b], (g[0] + g[b]) / 2.]] for b in b_list]))
        Bf = engine.sos_factor_time(_dev.to_dev(t3a, np.float64), W, P)
        z = engine.sos_outer(Bf, A, len(taus)).cpu().numpy()  # [T, n3, n1]
    if single:
        z = z[0]
    if s1 and s3:
        return z[..., 0, 0]
    if s1:
        return z[..., :, 0]
    if s3:
        return z[..., 0, :]
    return z


def ESA(evals, dip, g_idx, e_idx, f_idx, gamma, t1, t2, t3):
    """excited-state absorption gg -> ge -> e'e -> fe -> ee (sign -1); lime/signal/2DES.py:99-155"""
    E, g = _levels(evals, gamma)
    dip = np.asarray(dip)
    a = 0
    e_idx, f_idx = list(e_idx), list(f_idx)

    def terms(b, taus):
        W = np.zeros((len(taus), len(f_idx)), dtype=complex)
        for k, d in enumerate(f_idx):
            for c in e_idx:
                W[:, k] += dip[b, a] * dip[c, a] * dip[d, c] * dip[b, d] * _Gh(E, g, c, b, taus)
        return W, np.array([[E[d] - E[b], (g[d] + g[b]) / 2.] for d in f_idx])
    return _evaluate(E, g, dip, e_idx, terms, t1, t2, t3, -1.0)


def GSB(evals, dip, g_idx, e_idx, gamma, t1, t2, t3):
    """ground-state bleach gg -> ge -> gg' -> e'g' -> g'g'; lime/signal/2DES.py:158-203"""
    E, g = _levels(evals, gamma)
    dip = np.asarray(dip)
    a = 0
    e_idx, g_idx = list(e_idx), list(g_idx)

    def terms(b, taus):
        pairs = [(c, d) for c in g_idx for d in e_idx]
        W = np.zeros((len(taus), len(pairs)), dtype=complex)
        for k, (c, d) in enumerate(pairs):
            W[:, k] = dip[a, b] * dip[b, c] * dip[c, d] * dip[d, a] * _Gh(E, g, a, c, taus)
        return W, np.array([[E[d] - E[c], (g[d] + g[c]) / 2.] for c, d in pairs])
    return _evaluate(E, g, dip, e_idx, terms, t1, t2, t3, 1.0)


def SE(evals, dip, g_idx, e_idx, t1, t2, t3, gamma=None):
    """stimulated emission gg -> ge -> e'e -> g'e -> g'g'; lime/signal/2DES.py:207-247
    (lime's SE has no gamma argument -- it relies on the global `decay`; [ext] keyword `gamma`)"""
    E, g = _levels(evals, gamma if gamma is not None else 0.0)
    dip = np.asarray(dip)
    a = 0
    e_idx, g_idx = list(e_idx), list(g_idx)

    def terms(b, taus):
        pairs = [(c, d) for c in e_idx for d in g_idx]
        W = np.zeros((len(taus), len(pairs)), dtype=complex)
        for k, (c, d) in enumerate(pairs):
            W[:, k] = dip[a, b] * dip[c, a] * dip[d, c] * dip[b, d] * _Gh(E, g, c, b, taus)
        return W, np.array([[E[c] - E[d], (g[c] + g[d]) / 2.] for c, d in pairs])
    return _evaluate(E, g, dip, e_idx, terms, t1, t2, t3, 1.0)
