"""
HEOM index tables, Matsubara coefficients and the multi-index hierarchy propagator.

    state_number_enumerate, enr_state_dictionaries   lime/heom/heom.py:21-108   (bit-exact)
    _calc_matsubara_params                           lime/heom/heom.py:110-140
    HEOM [ext]                                       coupling rules lime/heom/heom.py:156-216
                                                     + system term lime/oqs.py:1854, RK4 lime/phys.py:636-649

The tables come from the native builder (limeb200_heom_build_tables); lime's dictionaries
are rebuilt from them so that keys, values and iteration order are identical.
"""
import numpy as np

from .. import engine
from .. import _dev


def state_number_enumerate(dims, excitations=None, state=None, idx=0):
    """Lexicographic enumeration (last index fastest) of the number states of `dims`
    restricted to sum(n) <= excitations (falsy `excitations`: no restriction);
    lime/heom/heom.py:21-72.  Yields ndarrays when excitations is None, tuples otherwise."""
    if state is not None or idx != 0:
        raise NotImplementedError('the internal recursion arguments of lime are not supported')
    states, _, _ = engine.heom_tables(dims, excitations)
    for row in states:
        if excitations is None:
            yield np.array(row, dtype=int)
        else:
            yield tuple(np.int64(x) for x in row)


def enr_state_dictionaries(dims, excitations):
    """(nstates, state2idx, idx2state); lime/heom/heom.py:78-108.  excitations=None raises
    TypeError exactly as lime does (ndarray keys are unhashable, lime/heom/heom.py:66,104)."""
    nstates = 0
    state2idx = {}
    idx2state = {}
    for state in state_number_enumerate(dims, excitations):
        state2idx[state] = nstates
        idx2state[nstates] = state
        nstates += 1
    return nstates, state2idx, idx2state


def _calc_matsubara_params(N_exp, coup_strength, cut_freq, temperature):
    """Drude-Lorentz Matsubara expansion, lime/heom/heom.py:110-140: returns lists (c, nu)"""
    c = []
    nu = []
    lam0 = coup_strength
    gam = cut_freq
    hbar = 1.
    beta = 1.0 / temperature
    g = 2 * np.pi / (beta * hbar)
    for k in range(N_exp):
        if k == 0:
            nu.append(gam)
            c.append(lam0 * gam * (1.0 / np.tan(gam * hbar * beta / 2.0) - 1j) / hbar)
        else:
            nu.append(k * g)
            c.append(4 * lam0 * gam * nu[k] / ((nu[k] ** 2 - gam ** 2) * beta * hbar ** 2))
    return c, nu


class HEOM:
    """[ext] multi-index hierarchy for `nbath` independent Drude-Lorentz baths, each expanded
    in N_exp exponentials (modes ordered bath-major: k = bath*N_exp + j), truncated at
    sum(n) <= N_cut.  Q: one coupling operator (lime's case) or a list of `nbath` operators.

    pref_dn / pref_up default to -i / -i, the rules as written in lime/heom/heom.py:181-216."""

    def __init__(self, H, Q, coup_strength, cut_freq, temperature, N_exp=2, N_cut=4,
                 pref_dn=-1j, pref_up=-1j, device_index=None, row_range=None):
        self.H = _dev.as_c128(H)
        n = self.H.shape[0]
        if isinstance(Q, (list, tuple)):
            Qs = [_dev.as_c128(q) for q in Q]
        else:
            Qs = [_dev.as_c128(Q)]
        nbath = len(Qs)
        lam = np.broadcast_to(np.asarray(coup_strength, dtype=float), (nbath,))
        gam = np.broadcast_to(np.asarray(cut_freq, dtype=float), (nbath,))
        c, nu, qmap = [], [], []
        for b in range(nbath):
            cb, nub = _calc_matsubara_params(N_exp, lam[b], gam[b], temperature)
            c += cb
            nu += nub
            qmap += [b] * N_exp
        self.c = np.array(c, dtype=complex)
        self.nu = np.array(nu, dtype=float)
        self.qmap = np.array(qmap, dtype=np.int32)
        self.N_cut = N_cut
        self.nmodes = nbath * N_exp
        dims = [N_cut + 1] * self.nmodes
        self.states, self.dn, self.up = engine.heom_tables(dims, N_cut)
        self.nhe = self.states.shape[0]
        self.n = n
        self.Q = np.stack(Qs)
        self.plan = engine.HeomPlan(self.H, self.Q, self.qmap, self.c, self.nu, self.states, self.dn, self.up,
                                    pref_dn=pref_dn, pref_up=pref_up, device_index=device_index,
                                    row_range=row_range)

    def initial(self, rho0):
        ado = np.zeros((self.nhe, self.n, self.n), dtype=np.complex128)
        ado[0] = rho0
        return ado

    def evolve(self, rho0, dt, Nt, e_ops=None, store_states=True):
        """returns a lime Result: observables (Nt,E) and rholist of the REDUCED density matrix
        (tier 0) after each step; the full final hierarchy is kept in result.ado"""
        from ..mol import Result
        ado, obs, traj = self.plan.run(self.initial(rho0), dt, Nt, e_ops=e_ops,
                                       traj_every=1 if store_states else 0)
        result = Result(dt=dt, Nt=Nt, rho0=rho0)
        result.observables = obs if obs is not None else np.zeros((Nt, 0), dtype=complex)
        result.rholist = [traj[k] for k in range(Nt)] if traj is not None else None
        result.ado = ado
        return result
