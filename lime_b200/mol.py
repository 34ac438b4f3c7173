"""Result container of the density-matrix solvers (lime/mol.py:78-104).

lime's constructor does arithmetic on nout=None / t0=None for every density-matrix solver
(lime/mol.py:92 called from lime/oqs.py:449,1670,1780) and raises TypeError; here the
defaults are nout=1, t0=0.0 -- the fields and their meaning are unchanged."""
import numpy as np


class Result:
    def __init__(self, description=None, psi0=None, rho0=None, dt=None, Nt=None, times=None,
                 t0=None, nout=None):
        self.description = description
        self.dt = dt
        self.timesteps = Nt
        self.observables = None
        self.rholist = None
        self.psilist = [psi0]
        self.psi = None
        self.rho0 = rho0
        self.psi0 = psi0
        nout = 1 if nout is None else nout
        t0 = 0.0 if t0 is None else t0
        self.nout = nout
        if Nt is None and times is not None:
            Nt = len(times)
        self.times = t0 + np.arange(Nt // nout) * dt * nout if (Nt is not None and dt is not None) else times

    def expect(self):
        return self.observables


# ---------------------------------------------------------------------------------------
# wave-function solver
# ---------------------------------------------------------------------------------------
def _dag(a):
    return a.conjugate().transpose()


def _quantum_dynamics(H, psi0, dt=0.001, Nt=1, e_ops=[], t0=0.0, nout=1, store_states=True, output='obs.dat'):
    """RK4 propagation of i d psi/dt = H psi (rk4 + tdse, lime/phys.py:636-649,902-903) with lime's sampling:
    entry 0 of `psilist` / `observables` is the INITIAL state, entry k the state after k*nout steps, for
    k < Nt//nout (lime/mol.py:1304-1372).  psi0 may be [N] or [B, N] ([ext]: a batch in one launch; then the
    states come back as [Nt//nout, B, N]).  The time loop is ONE launch of the Liouville-space RK4 kernel with
    R = -iH (limeb200_liouville_rk4_csr); the bilinear observables <psi|e|psi> are taken from the stored states."""
    from scipy.sparse import csr_matrix
    from . import engine
    psi0 = np.asarray(psi0, dtype=complex)
    single = psi0.ndim == 1
    v0 = psi0[None] if single else psi0
    nblk = Nt // nout
    R = csr_matrix(-1j * (H if hasattr(H, 'tocsr') else np.asarray(H, dtype=complex)))
    nsteps = max(nblk - 1, 0) * nout
    if nsteps > 0:
        vf, _, traj = engine.liouville_rk4(R, v0, dt, nsteps, traj_every=nout)
        states = np.concatenate([v0[None], traj], axis=0)          # [nblk, B, N]
    else:
        vf = v0.copy()
        states = v0[None][:max(nblk, 0)]
    e_ops = [] if e_ops is None else list(e_ops)
    obs = np.zeros((states.shape[0], v0.shape[0], len(e_ops)), dtype=complex)
    for j, e in enumerate(e_ops):
        ed = e.toarray() if hasattr(e, 'toarray') else np.asarray(e)
        obs[:, :, j] = np.einsum('kbi,ij,kbj->kb', states.conj(), ed, states)
    if not store_states:
        with open(output, 'w') as f:
            fmt = '{} ' * (len(e_ops) + 1) + '\n'
            for k in range(1, states.shape[0]):
                f.write(fmt.format(t0 + k * dt * nout, *obs[k, 0]))
        return vf[0] if single else vf
    result = Result(dt=dt, Nt=Nt, psi0=psi0, t0=t0, nout=nout)
    result.observables = obs[:, 0] if single else obs
    result.psilist = [s[0] for s in states] if single else list(states)
    return result


def driven_dynamics(H, psi0, dt=0.001, Nt=1, e_ops=None, nout=1, t0=0.0, return_result=True, sparse=True,
                    strict_parity=False):
    """Laser-driven wave-function dynamics, lime/mol.py:1473-1588: H = [H0, [H1, f1], ...], H(t) = H0 - sum_i f_i(t) H_i
    evaluated at the start time of the current block of `nout` steps (lime advances t once per block) and frozen over the
    four RK4 stages; entry 0 of the output is
    the initial state, entry k the state after k * nout steps, k < Nt // nout.

    The whole loop is one launch of the driven master-equation kernel (limeb200_qme_run with per-step drive
    coefficients): psi is column 0 of an N x N state with generator G(t) = -i H(t) and NO right generator, so that
    d/dt column = -i H(t) column.

    lime's calcH reads `Ht = H[0]; Ht += ...` (lime/mol.py:1510-1517); lime only runs with scipy.sparse operands here
    (its psi is a CSR column; an ndarray Hamiltonian raises ValueError), and for those `+=` rebinds Ht, so H(t) is
    evaluated afresh each step -- which is what happens here for sparse AND dense operands.  strict_parity=True applies
    the accumulation an aliased ndarray H[0] would see.  return_result=False writes psi.dat / obs.dat as lime does and
    returns None."""
    from . import engine, _dev
    e_ops = [] if e_ops is None else e_ops
    H0 = _dev.as_c128(H[0])
    N = H0.shape[0]
    psi0v = np.asarray(psi0.toarray() if hasattr(psi0, 'toarray') else psi0, dtype=complex).reshape(-1)
    nd = len(H) - 1

    def sample(nsteps, first):
        # lime advances t once per block of nout steps (`t += dt * nout` after the inner loop, lime/mol.py:1541-1547):
        # every step of a block sees H at the block's start time
        times = t0 + dt * nout * ((np.arange(nsteps) + first) // nout)
        f = np.array([[complex(H[i][1](t)) for i in range(1, len(H))] for t in times], dtype=complex).reshape(nsteps, nd)
        return np.cumsum(f, axis=0) if strict_parity else f

    plan = engine.QmePlan(N)
    plan.set_generator(-1j * H0)
    plan.set_right_generator(np.zeros((N, N), dtype=complex))
    for i in range(1, len(H)):
        Hi = _dev.as_c128(H[i][0])
        plan.add_drive(1j * Hi, np.zeros((N, N), dtype=complex))
    plan.finalize()
    rho = np.zeros((N, N), dtype=complex)
    rho[:, 0] = psi0v

    def expect(psi):
        return [np.vdot(psi, (e @ psi) if not hasattr(e, 'toarray') else e.dot(psi)) for e in e_ops]

    if return_result:
        nblk = Nt // nout
        nsteps = max(nblk - 1, 0) * nout
        result = Result(dt=dt, Nt=Nt, psi0=psi0, t0=t0, nout=nout)
        observables = np.zeros((nblk, len(e_ops)), dtype=complex)
        if nblk > 0:
            observables[0, :] = expect(psi0v)
        if nsteps > 0:
            _, _, traj = plan.run(rho, dt, nsteps, coef=sample(nsteps, 0), traj_every=nout)
            for k in range(1, nblk):
                psi = traj[k - 1][:, 0]
                observables[k, :] = expect(psi)
                result.psilist.append(psi.copy())
        result.observables = observables
        return result
    nblk = int(Nt / nout)
    nsteps = nblk * nout
    with open('psi.dat', 'w') as f_dm, open('obs.dat', 'w') as f_obs:
        if nsteps > 0:
            _, _, traj = plan.run(rho, dt, nsteps, coef=sample(nsteps, 0), traj_every=nout)
            fmt = '{} ' * (len(e_ops) + 1) + '\n'
            fmt_dm = '{} ' * (N + 1) + '\n'
            for k in range(nblk):
                psi = traj[k][:, 0]
                t = t0 + dt * nout * (k + 1)
                f_dm.write(fmt_dm.format(t, *psi))
                f_obs.write(fmt.format(t, *expect(psi)))
    return None


class SESolver:
    """time-dependent Schroedinger equation, lime/mol.py:1094-1277: time-independent Hamiltonians through the Liouville
    RK4 kernel, laser-driven ones (`pulse=`) through `driven_dynamics`"""

    def __init__(self, H=None):
        self.H = H
        self.groundstate = None

    def run(self, psi0=None, dt=0.01, Nt=1, e_ops=None, nout=1, t0=0.0, edip=None, pulse=None):
        if psi0 is None:
            psi0 = self.groundstate
        if pulse is not None:
            if edip is None:
                raise ValueError('Electric dipole must be provided for \
                                 laser-driven dynamics.')
            if isinstance(pulse, list):
                H = [self.H]
                for i in range(len(pulse)):
                    H.append([edip[i], pulse[i].efield])
            else:
                H = [self.H, [edip, pulse.efield]]
            return driven_dynamics(H=H, psi0=psi0, dt=dt, Nt=Nt, e_ops=e_ops, nout=nout, t0=t0)
        return _quantum_dynamics(self.H, psi0, dt=dt, Nt=Nt, e_ops=e_ops, nout=nout, t0=t0)

    def propagator(self, dt, Nt):
        """[U(0), U(dt), ...] with U(t) = exp(-iHt) by RK4 on the identity, lime/mol.py:1279-1301: the N columns are
        one batched launch"""
        N = self.H.shape[-1]
        r = _quantum_dynamics(self.H, np.identity(N, dtype=complex), dt=dt, Nt=Nt)
        return [np.ascontiguousarray(s.T) for s in r.psilist]          # batch member b = column b of U

    def correlation_2op_1t(self):
        pass

    def correlation_3op_1t(self, psi0, oplist, dt, Nt):
        """<A B(t) C>, lime/mol.py:1185-1211: ket and bra propagated together as a batch of two"""
        a_op, b_op, c_op = oplist
        batch = np.stack([c_op @ psi0, _dag(a_op) @ psi0])
        st = np.array(_quantum_dynamics(self.H, batch, dt=dt, Nt=Nt).psilist)       # [Nt, 2, N]
        return np.einsum('ki,ij,kj->k', st[:, 1].conj(), np.asarray(b_op), st[:, 0])

    def correlation_3op_2t(self, psi0, oplist, dt, Nt, Ntau):
        """<A(t) B(t+tau) C(t)>, lime/mol.py:1213-1246: lime runs 2 Nt propagations of Ntau steps in a Python loop;
        here they are ONE batched launch"""
        a_op, b_op, c_op = oplist
        psi_t = np.array(_quantum_dynamics(self.H, psi0, dt=dt, Nt=Nt).psilist)     # [Nt, N]
        batch = np.concatenate([psi_t @ np.asarray(c_op).T, psi_t @ _dag(np.asarray(a_op)).T], axis=0)
        st = np.array(_quantum_dynamics(self.H, batch, dt=dt, Nt=Ntau).psilist)     # [Ntau, 2 Nt, N]
        ket, bra = st[:, :Nt], st[:, Nt:]
        return np.ascontiguousarray(np.einsum('jti,ik,jtk->tj', bra.conj(), np.asarray(b_op), ket))

    def correlation_4op_1t(self, psi0, oplist, dt=0.005, Nt=1):
        a_op, b_op, c_op, d_op = oplist
        return self.correlation_3op_1t(psi0, [a_op, b_op @ c_op, d_op], dt, Nt)

    def correlation_4op_2t(self, psi0, oplist, dt=0.005, Nt=1, Ntau=1):
        a_op, b_op, c_op, d_op = oplist
        return self.correlation_3op_2t(psi0, [a_op, b_op @ c_op, d_op], dt, Nt, Ntau)
