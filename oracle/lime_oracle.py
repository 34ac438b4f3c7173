"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the lime density-matrix hot path.

A NumPy/SciPy restatement of the algorithms binggu56/lime runs on the path named
in BASELINE.json (Redfield/Lindblad RK4, HEOM, sum-over-states response
functions).  It exists so that the CUDA path can be checked on a machine where
/root/reference is absent.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it; the product
package lime_b200 never does.

Pinning: every function below that has a counterpart in the reference is
checked against the imported reference (oracle/ref_shim.py) and against the
reference's only golden data (examples/cor.dat + examples/dm.dat) by
oracle/gen_golden.py and tests/test_oracle_vs_golden.py; the frozen outputs live
in tests/golden/.  The functions deliberately perform the *same sequence of
NumPy/SciPy calls* as the reference (e.g. l^dag l is recomputed on every call,
Tr(d rho) forms the full product) so that (i) results agree with it to the last
bit where NumPy is deterministic and (ii) timing the oracle is a fair stand-in
for timing lime's CPU path.

The multi-index hierarchy (heom_rhs / heom_rk4): lime ships the index tables, the
Matsubara coefficients and an excerpt stating the coupling rules inside a dead
__main__ block (lime/heom/heom.py:142-216) that uses names lime never defines.
oracle/gen_golden.py exec's that rule loop (:156-216) VERBATIM with those names
supplied and freezes the hierarchy Liouvillian it builds
(tests/golden/heom_rules.npz); heom_rhs minus the system term equals that matrix
applied to vec(ADOs) to 2e-16 -- PINNED on the reference's own rules for lime's
single-Q case and for bath-dependent coupling operators.  The system term
-i[H, rho_n] is lime/oqs.py:1854 and the integrator lime/phys.py:636-649 (both
pinned through _heom_dl / rk4 elsewhere).
PARITY UNPINNED: rkf45 (lime ships only examples/rkf45_test.py, which imports a
module that is not in the tree) -- accuracy pinned on that file's analytic problems.
All citations are file:line in /root/reference.
"""
import numpy as np
from scipy.sparse import issparse, identity, kron, csr_matrix, lil_matrix
import scipy.linalg

# lime/units.py:2,5,6,8,9
au2fs = 2.41888432651e-2
au2k = 315775.13
au2ev = 27.211386
au2mev = 27211.386
au2wavenumber = 219474.6305


# --------------------------------------------------------------------------
# L1 primitives                                             lime/phys.py
# --------------------------------------------------------------------------
def dag(a):
    """conjugate transpose, lime/phys.py:758-759"""
    return a.conjugate().transpose()


def comm(A, B):
    """AB - BA through np.dot, lime/phys.py:741-743"""
    assert A.shape == B.shape
    return np.dot(A, B) - np.dot(B, A)


def anticomm(A, B):
    """AB + BA through np.dot, lime/phys.py:746-748"""
    assert A.shape == B.shape
    return np.dot(A, B) + np.dot(B, A)


def commutator(A, B):
    """AB - BA through .dot (works for scipy.sparse), lime/phys.py:736-738"""
    assert A.shape == B.shape
    return A.dot(B) - B.dot(A)


def anticommutator(A, B):
    """lime/phys.py:750-752"""
    assert A.shape == B.shape
    return A.dot(B) + B.dot(A)


def transform(A, v):
    """v^dag A v, lime/phys.py:706-718"""
    return dag(v).dot(A.dot(v))


def obs_dm(rho, d):
    """Tr(d rho) via the full product, lime/phys.py:837-844"""
    return d.dot(rho).diagonal().sum()


def isherm(a):
    """lime/phys.py:1429"""
    return np.allclose(a, dag(a))


def rk4(y, fun, dt, *args):
    """classical RK4 on an autonomous RHS; y updated IN PLACE and returned,
    lime/phys.py:636-649"""
    half = dt / 2.0
    k1 = fun(y, *args)
    k2 = fun(y + k1 * half, *args)
    k3 = fun(y + k2 * half, *args)
    k4 = fun(y + k3 * dt, *args)
    y += (k1 + 2 * k2 + 2 * k3 + k4) / 6. * dt
    return y


def pauli():
    """real s0, sx, complex sy, real sz, lime/phys.py:773-786"""
    s0 = np.identity(2)
    sx = np.array([[0., 1.], [1., 0.]])
    sy = np.array([[0., -1j], [1j, 0.]])
    sz = np.array([[1., 0.], [0., -1.]])
    return s0, sx, sy, sz


def basis(N, j):
    """lime/phys.py:879-899"""
    b = np.zeros(N)
    b[j] = 1.0
    return b


def ket2dm(psi):
    """lime/phys.py:579-594"""
    return np.einsum("i, j -> ij", psi, psi.conj())


def destroy(N):
    """CSR annihilation operator, lime/phys.py:615-633"""
    a = lil_matrix((N, N))
    a.setdiag(np.sqrt(np.arange(1, N)), 1)
    return a.tocsr()


# --------------------------------------------------------------------------
# L4 Lindblad                                                lime/oqs.py
# --------------------------------------------------------------------------
def lindbladian(l, rho):
    """l rho l^dag - 1/2 {l^dag l, rho}; recomputes l^dag l, lime/oqs.py:716-723"""
    return np.dot(l, np.dot(rho, dag(l))) - 0.5 * anticomm(np.dot(dag(l), l), rho)


def liouvillian(rho, H, c_ops):
    """-i[H,rho] + sum_m D[l_m](rho), lime/oqs.py:706-713"""
    out = -1j * comm(H, rho)
    for c in c_ops:
        out += lindbladian(c, rho)
    return out


def lindbladian_sp(l, rho):
    """`.dot` flavour used for scipy.sparse operands, lime/phys.py:570-577"""
    return l.dot(rho.dot(dag(l))) - 0.5 * anticomm_any(dag(l).dot(l), rho)


def anticomm_any(A, B):
    if issparse(A) or issparse(B):
        return A.dot(B) + B.dot(A)
    return anticomm(A, B)


def liouvillian_sp(rho, H, c_ops):
    """lime/phys.py:561-568 (the version the golden cor.dat/dm.dat was made with;
    all operands CSR so np.dot semantics == .dot)"""
    out = -1j * (H.dot(rho) - rho.dot(H))
    for c in c_ops:
        out += lindbladian_sp(c, rho)
    return out


def lindblad(H, rho0, c_ops, e_ops=None, Nt=1, dt=0.005):
    """_lindblad with return_result=True, lime/oqs.py:1590-1688.
    Returns (observables (Nt,E) c128, rholist): sample k is taken AFTER step k+1."""
    rho = rho0.copy().astype(complex)
    if e_ops is None:
        e_ops = []
    observables = np.zeros((Nt, len(e_ops)), dtype=complex)
    rholist = []
    for k in range(Nt):
        rho = rk4(rho, liouvillian, dt, H, c_ops)
        rholist.append(rho.copy())
        observables[k, :] = [obs_dm(rho, op) for op in e_ops]
    return observables, rholist


def lindblad_driven(H, rho0, c_ops=None, e_ops=None, Nt=1, dt=0.005, t0=0.,
                    strict_parity=False):
    """_lindblad_driven, lime/oqs.py:1691-1800.  H = [H0, [H1, f1], ...];
    Ht = H0 - sum_i f_i(t) H_i evaluated once per step at t+dt and frozen over
    the four stages.  strict_parity=True reproduces the aliasing bug at
    lime/oqs.py:1717-1724 (`Ht = H[0]; Ht += ...` accumulates into H[0], which the
    caller's list then sees); the default builds a fresh Ht each step."""
    if c_ops is None:
        c_ops = []
    if e_ops is None:
        e_ops = []
    rho = rho0.copy().astype(complex)
    t = t0
    observables = np.zeros((Nt, len(e_ops)), dtype=complex)
    rholist = []
    for k in range(Nt):
        t += dt
        if strict_parity:
            Ht = H[0]
            for i in range(1, len(H)):
                Ht += - H[i][1](t) * H[i][0]
        else:
            Ht = H[0].astype(complex)
            for i in range(1, len(H)):
                Ht = Ht - H[i][1](t) * H[i][0]
        rho = rk4(rho, liouvillian, dt, Ht, c_ops)
        rholist.append(rho.copy())
        observables[k, :] = [obs_dm(rho, op) for op in e_ops]
    return observables, rholist


def correlation_3p_1t(H, rho0, ops, c_ops, tlist):
    """<A B(t) C> by quantum regression on CSR operands -- the generator of the
    reference's golden examples/cor.dat + dm.dat, lime/correlation.py:17-70 with
    dyn = lime/phys.py:561-568.  Returns (t[Nt], cor[Nt], rho[Nt,n,n])."""
    A, B, C = ops
    rho = C.dot(rho0.dot(A))
    Nt = len(tlist)
    dt = tlist[1] - tlist[0]
    t = 0.0
    ts, cors, rhos = [], [], []
    for k in range(Nt):
        t += dt
        rho = rk4(rho, liouvillian_sp, dt, H, c_ops)
        cors.append(B.dot(rho).diagonal().sum())
        ts.append(t)
        rhos.append(rho.toarray() if issparse(rho) else np.array(rho))
    return np.array(ts), np.array(cors), np.array(rhos)


def correlation_2p_1t(H, rho0, ops, c_ops, dt, Nt, output=None):
    """<A(t) B> by quantum regression with CSR operands and a CSR rho, lime/oqs.py:726-800;
    returns cor[Nt] and writes `t cor` lines to `output` when given"""
    A, B = ops
    rho = csr_matrix(B.dot(rho0))
    H = csr_matrix(H)
    A = csr_matrix(A)
    c_sp = [csr_matrix(c) for c in c_ops]
    t = 0.0
    cor = np.zeros(Nt, dtype=complex)
    lines = []
    for k in range(Nt):
        t += dt
        rho = rk4(rho, liouvillian_sp, dt, H, c_sp)
        tmp = A.dot(rho).diagonal().sum()                     # obs_dm on sparse operands, lime/phys.py:837-844
        cor[k] = tmp
        lines.append('{} {} \n'.format(t, tmp))
    if output is not None:
        with open(output, 'w') as f:
            f.writelines(lines)
    return cor


def getG(L, t, w=None, domain='time'):
    """lime/oqs.py:474-526 (dense eig of L, U2 = inv(U1), the two einsums as written)"""
    evals1, U1 = scipy.linalg.eig(L.todense() if issparse(L) else L)
    U2 = scipy.linalg.inv(U1)
    if domain == 'time':
        U = -1j * np.exp(-1j * evals1[:, np.newaxis] * t[np.newaxis, :])
        return np.einsum('aj, jk, jb -> abk', U1, U, U2)
    W = 1. / ((w[:, np.newaxis] - evals1[np.newaxis, :]))
    return np.einsum('an, nk, bn ->abk', U1, W, U2.conj())


def lindblad_correlation_3op_1t(H, c_ops, rho0, oplist, dt, Nt):
    """<A B(t) C>, lime/oqs.py:1227-1246"""
    a_op, b_op, c_op = oplist
    obs, _ = lindblad(H, c_op @ rho0 @ a_op, c_ops, e_ops=[b_op], dt=dt, Nt=Nt)
    return obs[:, 0]


def lindblad_correlation_3op_2t(H, c_ops, rho0, ops, dt, Nt, Ntau):
    """<A(t) B(t+tau) C(t)>, lime/oqs.py:1268-1299: Nt independent runs of Ntau steps"""
    _, rho_t = lindblad(H, rho0, c_ops, dt=dt, Nt=Nt)
    a_op, b_op, c_op = ops
    corr = np.zeros([Nt, Ntau], dtype=complex)
    for i, rho in enumerate(rho_t):
        obs, _ = lindblad(H, c_op @ rho @ a_op, c_ops, e_ops=[b_op], dt=dt, Nt=Ntau)
        corr[i, :] = obs[:, 0]
    return corr


def etpa_core(omegaps, Es, edip, jta, t1, t2, g_idx, e_idx, f_idx):
    """_etpa, lime/signal/sos.py:1171-1223: the double time integrals as lime's loops write them"""
    T1, T2 = np.meshgrid(t1, t2)
    theta = np.heaviside(T2 - T1, 0.5)
    signal = np.zeros(len(omegaps), dtype=complex)
    g = g_idx
    for j, omegap in enumerate(omegaps):
        omega1 = omega2 = omegap / 2.
        for f in f_idx:
            for e in e_idx:
                detuning2 = Es[f] - Es[e] - omega2
                detuning1 = Es[e] - Es[g] - omega1
                D = edip[e, g] * edip[f, e]
                signal[j] += (D * np.sum(theta * np.exp(1j * detuning2 * T2 + 1j * detuning1 * T1) * jta)).item()
                detuning2 = Es[f] - Es[e] - omega1
                detuning1 = Es[e] - Es[g] - omega2
                signal[j] += (D * np.sum(theta * np.exp(1j * detuning2 * T2 + 1j * detuning1 * T1) * jta.T)).item()
    return signal


# --------------------------------------------------------------------------
# L2 superoperator algebra (row-major vec)           lime/superoperator.py
# --------------------------------------------------------------------------
def dm2vec(rho):
    """lime/superoperator.py:131-151"""
    if issparse(rho):
        n, m = rho.shape
        return rho.tolil().reshape((n * m, 1))
    return rho.flatten()


def operator_to_superoperator(a, kind='commutator'):
    """kron(a,I) -/+ kron(I,a^T); lime/superoperator.py:201-247"""
    N = a.shape[-1]
    one = identity(N)
    if kind in ['commutator', 'c', '-']:
        return kron(a, one) - kron(one, a.T)
    elif kind in ['left', 'l']:
        return kron(a, one)
    elif kind in ['right', 'r']:
        return kron(one, a.T)
    elif kind in ['anticommutator', 'a', '+']:
        return kron(a, one) + kron(one, a.T)
    raise ValueError('Error: superoperator {} does not exist.'.format(kind))


op2sop = operator_to_superoperator


def left(a):
    """lime/superoperator.py:257-262"""
    n = a.toarray().shape[-1] if issparse(a) else a.shape[-1]
    return kron(a, identity(n))


def right(a):
    """lime/superoperator.py:264-271"""
    n = a.toarray().shape[-1] if issparse(a) else a.shape[-1]
    return kron(identity(n), a.T)


def lindblad_dissipator(l):
    """lime/superoperator.py:250-252"""
    return kron(l, l.conj()) - 0.5 * operator_to_superoperator(dag(l).dot(l), kind='anticommutator')


def liouvillian_super(H, c_ops):
    """lime/superoperator.py:30-59"""
    if c_ops is None:
        c_ops = []
    L = -1j * operator_to_superoperator(H)
    for c in c_ops:
        L = L + lindblad_dissipator(c)
    return L


# --------------------------------------------------------------------------
# L4 Redfield                                                lime/oqs.py
# --------------------------------------------------------------------------
def redfield_tensor(H, a_ops, spectra, secular=False):
    """eigenbasis Redfield generator, d/dt vec(rho) = R vec(rho); `secular` is
    ignored exactly as in lime/oqs.py:528-579.  Returns (csr R, evecs)."""
    for a in a_ops:
        dense = a.todense() if issparse(a) else a
        if not isherm(dense):
            raise TypeError("Operators in a_ops must be Hermitian.")
    evals, evecs = scipy.linalg.eigh(H.todense() if issparse(H) else H)
    W = np.real(evals[:, np.newaxis] - evals[np.newaxis, :])
    K, N = len(a_ops), len(evals)
    C = []
    for k in range(K):
        c = np.zeros((N, N))
        for n in range(N):
            for m in range(N):
                c[n, m] = spectra[k](-W[n, m])
        C.append(c)
    A = [transform(a, evecs) for a in a_ops]
    Lam = [C[k] * A[k] for k in range(K)]
    R = 0
    for k in range(K):
        R += op2sop(A[k]).dot(left(Lam[k]) - right(dag(Lam[k])))
    return csr_matrix(-1j * op2sop(np.diag(evals)) - R), evecs


def redfield_parts(H, a_ops, spectra):
    """(evals, evecs, A_k, Lambda_k): the O(K N^2) operator form of the same
    generator (what lime's `func`, lime/oqs.py:840-850, evaluates)."""
    evals, evecs = scipy.linalg.eigh(H.todense() if issparse(H) else H)
    W = np.real(evals[:, np.newaxis] - evals[np.newaxis, :])
    N = len(evals)
    A, Lam = [], []
    for k, a in enumerate(a_ops):
        c = np.zeros((N, N))
        for n in range(N):
            for m in range(N):
                c[n, m] = spectra[k](-W[n, m])
        ak = transform(a, evecs)
        A.append(ak)
        Lam.append(c * ak)
    return evals, evecs, A, Lam


def _rhs(v, R):
    """lime/oqs.py:471-472"""
    return R.dot(v)


def redfield(R, rho0, evecs=None, Nt=1, dt=0.005, t0=0, e_ops=[]):
    """_redfield with return_result=True, lime/oqs.py:373-468.  Observables are
    evaluated in the eigenbasis, rholist is transformed back with v rho v^dag."""
    N = rho0.shape[0]
    if e_ops is None:
        e_ops = []
    if evecs is not None:
        rho0 = transform(rho0, evecs)
        e_ops = [transform(e, evecs) for e in e_ops]
    v = dm2vec(rho0.copy()).astype(complex)
    observables = np.zeros((Nt, len(e_ops)), dtype=complex)
    rholist = []
    for k in range(Nt):
        v = rk4(v, _rhs, dt, R)
        m = np.reshape(v, (N, N))
        rholist.append(transform(m, dag(evecs)))
        observables[k, :] = [obs_dm(m, e) for e in e_ops]
    return observables, rholist


def expm_eom(A, t):
    """U(t_k) = e^{A t_k} by RK4 on the identity, dt = t[1]-t[0]; U(t_0) = 1,
    lime/phys.py:1352-1406 (method 'EOM')"""
    U = identity(A.shape[-1], dtype=complex).tocsr()
    out = []
    dt = t[1] - t[0]
    for k in range(len(t)):
        out.append(U.copy())
        U = rk4(U, lambda b, M: M.dot(b), dt, A)
    return out


def redfield_propagator(R, t, method='SOS'):
    """Redfield_solver.propagator, lime/oqs.py:169-223: U[a,b,k] = (e^{R t_k})_{ab}"""
    t = np.asarray(t)
    if method == 'EOM':
        return np.dstack([u.toarray() for u in expm_eom(R, t)])
    evals, U1 = scipy.linalg.eig(R.toarray())
    U2 = scipy.linalg.inv(U1)
    E = np.exp(evals[:, np.newaxis] * t[np.newaxis, :])
    return np.einsum('aj, jk, jb -> abk', U1, E, U2, optimize=True)


def redfield_expect(U, evecs, rho0, e_ops):
    """Redfield_solver.expect (ndarray propagator branch), lime/oqs.py:225-252"""
    v0 = dm2vec(transform(rho0, evecs))
    e_ops = [transform(e, evecs) for e in e_ops]
    Nt = U.shape[-1]
    out = np.zeros((Nt, len(e_ops)), dtype=complex)
    rho = np.tensordot(U, v0, axes=([1], [0]))
    for j, e in enumerate(e_ops):
        out[:, j] = dm2vec(e).dot(rho)
    return out


def redfield_correlation_4op_3t(G, dim, rho0, oplist, signature):
    """<<I|A G(t3) B G(t2) C G(t1) D|rho0>>, lime/oqs.py:277-366.  G = -i U is
    (N^2,N^2,Nt); result[i,j,k]: axis 0 is the LAST interval.  No basis change
    is applied to rho0 / operators (the reference does not either)."""
    if len(oplist) != 4:
        raise ValueError('Number of operators is not 4.')
    a, b, c, d = [operator_to_superoperator(o, s) for o, s in zip(oplist, signature)]
    one = dm2vec(np.identity(dim))
    rho = d.dot(dm2vec(rho0.toarray() if issparse(rho0) else rho0))
    x = np.tensordot(G, rho, axes=((1), (0)))
    x = c.dot(x)
    x = np.tensordot(G, x, axes=([1], [0]))
    x = np.tensordot(np.asarray(b.todense()), x, axes=([1], [0]))
    x = np.tensordot(G, x, axes=([1], [0]))
    return np.einsum('a, ab, bijk -> ijk', one, np.asarray(a.todense()), x, optimize=True)


# --------------------------------------------------------------------------
# HEOM                                      lime/oqs.py, lime/heom/heom.py
# --------------------------------------------------------------------------
def heom_dl(H, rho0, sz, temperature, cutoff, reorganization, nado, dt, nt):
    """_heom_dl, lime/oqs.py:1802-1865: single Drude mode, high-temperature,
    in-place Gauss-Seidel Euler sweep; tier 0 is advanced twice per step
    (:1850-1851 and n=0 of :1853-1857, where the n*(...) term vanishes and
    ado[:,:,-1] is touched only through a zero factor); last tier is never
    updated.  Returns (final ado[n,n,nado], trajectory[nt,n,n] of tier 0)."""
    nst = H.shape[0]
    ado = np.zeros((nst, nst, nado), dtype=np.complex128)
    ado[:, :, 0] = rho0
    gamma = cutoff
    T = temperature / au2k
    a = np.pi * reorganization * T
    b = 0.0
    traj = np.zeros((nt, nst, nst), dtype=complex)
    for k in range(nt):
        ado[:, :, 0] += -1j * commutator(H, ado[:, :, 0]) * dt - \
            commutator(sz, ado[:, :, 1]) * dt
        for n in range(nado - 1):
            ado[:, :, n] += -1j * commutator(H, ado[:, :, n]) * dt + \
                (- commutator(sz, ado[:, :, n + 1]) - n * gamma * ado[:, :, n] + n *
                 (a * commutator(sz, ado[:, :, n - 1]) +
                  1j * b * anticommutator(sz, ado[:, :, n - 1]))) * dt
        traj[k] = ado[:, :, 0]
    return ado, traj


def state_number_enumerate(dims, excitations=None, state=None, idx=0):
    """lexicographic enumeration (last index fastest) with prefix pruning
    sum(state[:idx]) > excitations; falsy `excitations` disables the restriction.
    lime/heom/heom.py:21-72"""
    if state is None:
        state = np.zeros(len(dims), dtype=int)
    if excitations and sum(state[0:idx]) > excitations:
        return
    if idx == len(dims):
        if excitations is None:
            yield np.array(state)
        else:
            yield tuple(state)
        return
    for n in range(dims[idx]):
        state[idx] = n
        yield from state_number_enumerate(dims, excitations, state, idx + 1)


def enr_state_dictionaries(dims, excitations):
    """(nstates, state->idx, idx->state), lime/heom/heom.py:78-108"""
    n = 0
    s2i, i2s = {}, {}
    for s in state_number_enumerate(dims, excitations):
        s2i[s] = n
        i2s[n] = s
        n += 1
    return n, s2i, i2s


def enr_states_fast(dims, excitations):
    """same ordering as enr_state_dictionaries, returned as an (N_he, N_m) int64
    array; iterative (used for the large tables in tests)."""
    dims = list(dims)
    nm = len(dims)
    out = []
    state = [0] * nm

    def rec(idx, tot):
        if excitations and tot > excitations:
            return
        if idx == nm:
            out.append(tuple(state))
            return
        for n in range(dims[idx]):
            state[idx] = n
            rec(idx + 1, tot + n)
        state[idx] = 0
    # NOTE: the reference prunes on sum(state[:idx]) *before* looking at idx, so
    # the leaf test uses the full sum: identical to passing tot here.
    rec(0, 0)
    return np.array(out, dtype=np.int64).reshape(len(out), nm)


def calc_matsubara_params(N_exp, coup_strength, cut_freq, temperature):
    """Drude-Lorentz Matsubara expansion, lime/heom/heom.py:110-140
    (nu_0 = gamma, c_0 = lam*gam*(cot(gam*beta/2) - i); nu_k = 2 pi k / beta,
    c_k = 4 lam gam nu_k / ((nu_k^2 - gam^2) beta))"""
    c, nu = [], []
    lam0, gam = coup_strength, cut_freq
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


def heom_tables(dims, excitations):
    """(states[N_he,N_m], dn[N_he,N_m], up[N_he,N_m]) neighbour tables: dn[a,k] =
    index of n - e_k (or -1 when n_k == 0), up[a,k] = index of n + e_k (or -1
    when sum(n) > N_c - 1 or n_k + 1 == dims[k]).  Connectivity per
    lime/heom/heom.py:176-216."""
    nhe, s2i, i2s = enr_state_dictionaries(dims, excitations)
    nm = len(dims)
    states = np.array([i2s[i] for i in range(nhe)], dtype=np.int64).reshape(nhe, nm)
    dn = -np.ones((nhe, nm), dtype=np.int64)
    up = -np.ones((nhe, nm), dtype=np.int64)
    for a in range(nhe):
        s = list(states[a])
        tot = sum(s)
        for k in range(nm):
            if s[k] >= 1:
                s[k] -= 1
                dn[a, k] = s2i[tuple(s)]
                s[k] += 1
            if tot <= excitations - 1 and s[k] + 1 < dims[k]:
                s[k] += 1
                up[a, k] = s2i[tuple(s)]
                s[k] -= 1
    return states, dn, up


def heom_rhs(ado, H, Q, qmap, c, nu, states, dn, up, pref_dn=-1j, pref_up=-1j):
    """Multi-index HEOM right-hand side (pinned on the reference's rule loop, tests/golden/heom_rules.npz).
    ado: (N_he, n, n).  Q: (N_q, n, n) coupling operators, qmap[k] -> which Q mode
    k uses (lime has a single Q: qmap = 0).  Rules, lime/heom/heom.py:156-216:
      diagonal   -sum_k n_k nu_k rho_n                                   :167-173
      down n_k>=1  pref_dn * n_k (c_k Q rho_{n-e_k} - c_k^* rho_{n-e_k} Q) :181-198
      up           pref_up * [Q, rho_{n+e_k}]                             :200-216
    plus the system term -i[H, rho_n] (lime/oqs.py:1854)."""
    out = -1j * (np.matmul(H, ado) - np.matmul(ado, H))
    damp = states.astype(float) @ np.asarray(nu, dtype=float)
    out -= damp[:, None, None] * ado
    nm = states.shape[1]
    for k in range(nm):
        q = Q[qmap[k]]
        ck = complex(c[k])
        has = dn[:, k] >= 0
        if has.any():
            nb = ado[dn[has, k]]
            nk = states[has, k].astype(float)[:, None, None]
            out[has] += pref_dn * nk * (ck * np.matmul(q, nb) - np.conj(ck) * np.matmul(nb, q))
        has = up[:, k] >= 0
        if has.any():
            nb = ado[up[has, k]]
            out[has] += pref_up * (np.matmul(q, nb) - np.matmul(nb, q))
    return out


def heom_rk4(ado0, H, Q, qmap, c, nu, states, dn, up, dt, nsteps,
             pref_dn=-1j, pref_up=-1j, e_ops=None, store=False):
    """RK4 (lime/phys.py:636-649) on the full ADO tensor.  Returns
    (ado_final, observables[nsteps,E] of tier 0, [tier-0 trajectory])."""
    ado = np.array(ado0, dtype=complex)
    e_ops = [] if e_ops is None else e_ops
    obs = np.zeros((nsteps, len(e_ops)), dtype=complex)
    traj = []
    for s in range(nsteps):
        ado = rk4(ado, heom_rhs, dt, H, Q, qmap, c, nu, states, dn, up, pref_dn, pref_up)
        obs[s, :] = [obs_dm(ado[0], e) for e in e_ops]
        if store:
            traj.append(ado[0].copy())
    return ado, obs, traj


# --------------------------------------------------------------------------
# L5 sum-over-states response functions              lime/signal/sos.py
# --------------------------------------------------------------------------
def _grid(omega1, omega3):
    # np.meshgrid default indexing='xy': arrays are (len(omega3), len(omega1))
    return np.meshgrid(omega1, omega3)


def GSB(evals, dip, omega1, omega3, tau2, g_idx, e_idx, gamma):
    """lime/signal/sos.py:478-528 (tau2 unused; a = c = 0)"""
    signal = np.zeros((len(omega1), len(omega3)), dtype=complex)
    a = c = 0
    pump, probe = _grid(omega1, omega3)
    for b in e_idx:
        G_ab = 1. / (pump - (evals[a] - evals[b]) + 1j * (gamma[a] + gamma[b]) / 2.0)
        for d in e_idx:
            G_dc = 1. / (probe - (evals[d] - evals[c]) + 1j * (gamma[d] + gamma[c]) / 2.0)
            signal += dip[a, b] * dip[b, c] * dip[c, d] * dip[d, a] * G_dc * G_ab
    return signal


def SE(evals, dip, omega1, omega3, tau2, g_idx, e_idx, gamma):
    """lime/signal/sos.py:576-635"""
    signal = np.zeros((len(omega1), len(omega3)), dtype=complex)
    a = 0
    pump, probe = _grid(omega1, omega3)
    for b in e_idx:
        G_ab = 1. / (pump - (evals[a] - evals[b]) + 1j * (gamma[a] + gamma[b]) / 2.0)
        for c in e_idx:
            U_cb = -1j * np.exp(-1j * (evals[c] - evals[b]) * tau2 - (gamma[c] + gamma[b]) / 2. * tau2)
            for d in g_idx:
                G_cd = 1. / (probe - (evals[c] - evals[d]) + 1j * (gamma[c] + gamma[d]) / 2.0)
                signal += dip[a, b] * dip[c, a] * dip[d, c] * dip[b, d] * G_cd * U_cb * G_ab
    return signal


def ESA(evals, dip, omega1, omega3, tau2, g_idx, e_idx, f_idx, gamma):
    """lime/signal/sos.py:348-407 (overall sign -1)"""
    signal = np.zeros((len(omega1), len(omega3)), dtype=complex)
    a = 0
    pump, probe = _grid(omega1, omega3)
    for b in e_idx:
        G_ab = 1. / (pump - (evals[a] - evals[b]) + 1j * (gamma[a] + gamma[b]) / 2.0)
        for c in e_idx:
            U_cb = -1j * np.exp(-1j * (evals[c] - evals[b]) * tau2 - (gamma[c] + gamma[b]) / 2. * tau2)
            for d in f_idx:
                G_db = 1. / (probe - (evals[d] - evals[b]) + 1j * (gamma[d] + gamma[b]) / 2.0)
                signal += dip[b, a] * dip[c, a] * dip[d, c] * dip[b, d] * G_db * U_cb * G_ab
    return -1 * signal


def photon_echo_core(evals, edip, omega1, omega3, t2, g_idx, e_idx, f_idx, gamma):
    """_photon_echo = GSB + SE + ESA, lime/signal/sos.py:695-729"""
    return (GSB(evals, edip, omega1, omega3, t2, g_idx, e_idx, gamma)
            + SE(evals, edip, omega1, omega3, t2, g_idx, e_idx, gamma)
            + ESA(evals, edip, omega1, omega3, t2, g_idx, e_idx, f_idx, gamma))


def SE_t3(E, dip, omega1, omega2, t3, g_idx, e_idx, gamma, dephasing=10 / au2mev):
    """_SE: (omega1, omega2) grid at detection time t3 with pure dephasing,
    lime/signal/sos.py:638-692"""
    signal = np.zeros((len(omega2), len(omega1)), dtype=complex)
    a = 0
    pump, probe = np.meshgrid(omega1, omega2)
    N = len(E)
    gD = np.ones((N, N)) * dephasing
    np.fill_diagonal(gD, 0)
    for b in e_idx:
        G_ab = 1. / (pump - (E[a] - E[b]) + 1j * ((gamma[a] + gamma[b]) / 2.0 + gD[a, b]))
        for c in e_idx:
            U_cb = 1. / (probe - (E[c] - E[b]) + 1j * ((gamma[c] + gamma[b]) / 2. + gD[c, b]))
            for d in g_idx:
                G_cd = -1j * np.exp(-1j * (E[c] - E[d]) * t3 - ((gamma[c] + gamma[d]) / 2.0 + gD[c, d]) * t3)
                signal += dip[a, b] * dip[c, a] * dip[d, c] * dip[b, d] * G_cd * U_cb * G_ab
    return signal


def ESA_t3(evals, dip, omega1, omega2, t3, g_idx, e_idx, f_idx, gamma, dephasing=10 / au2mev):
    """_ESA, lime/signal/sos.py:410-476"""
    signal = np.zeros((len(omega2), len(omega1)), dtype=complex)
    a = 0
    pump, probe = np.meshgrid(omega1, omega2)
    N = len(evals)
    gD = np.ones((N, N), dtype=float) * dephasing
    np.fill_diagonal(gD, 0)
    for b in e_idx:
        G_ab = 1. / (pump - (evals[a] - evals[b]) + 1j * ((gamma[a] + gamma[b]) / 2.0 + gD[a, b]))
        for c in e_idx:
            U_cb = 1. / (probe - (evals[c] - evals[b]) + 1j * ((gamma[c] + gamma[b]) / 2. + gD[c, b]))
            for d in f_idx:
                G_db = -1j * np.exp(-1j * (evals[d] - evals[b]) * t3 -
                                    ((gamma[d] + gamma[b]) / 2.0 + gD[d, b]) * t3)
                signal += dip[b, a] * dip[c, a] * dip[d, c] * dip[b, d] * G_db * U_cb * G_ab
    return -1 * signal


def DQC_R1(evals, dip, omega1=None, omega2=[], omega3=None, tau1=None, tau3=None,
           g_idx=[0], e_idx=None, f_idx=None, gamma=None):
    """lime/signal/sos.py:904-992.  Quirk kept: in the (omega1, omega2; tau3) branch
    G_ba is evaluated at `probe` (omega2), not `pump` (:949)."""
    a = 0
    if omega3 is None and tau3 is not None:
        signal = np.zeros((len(omega1), len(omega2)), dtype=complex)
        for i in range(len(omega1)):
            for j in range(len(omega2)):
                probe = omega2[j]
                for b in e_idx:
                    G_ba = 1. / (probe - (evals[b] - evals[a]) + 1j * (gamma[b] + gamma[a]) / 2.0)
                    for c in f_idx:
                        G_ca = 1. / (probe - (evals[c] - evals[a]) + 1j * (gamma[c] + gamma[a]) / 2.0)
                        for d in e_idx:
                            U_cd = -1j * np.exp(-1j * (evals[c] - evals[d]) * tau3 - (gamma[c] + gamma[d]) / 2. * tau3)
                            signal[i, j] += dip[b, a] * dip[c, b] * dip[d, a] * dip[d, c] * G_ba * G_ca * U_cd
    elif omega1 is None and tau1 is not None:
        signal = np.zeros((len(omega2), len(omega3)), dtype=complex)
        for i in range(len(omega2)):
            pump = omega2[i]
            for j in range(len(omega3)):
                probe = omega3[j]
                for b in e_idx:
                    U_ba = -1j * np.exp(-1j * (evals[b] - evals[a]) * tau1 - (gamma[b] + gamma[a]) / 2. * tau1)
                    for c in f_idx:
                        G_ca = 1. / (pump - (evals[c] - evals[a]) + 1j * (gamma[c] + gamma[a]) / 2.0)
                        for d in e_idx:
                            G_cd = 1. / (probe - (evals[c] - evals[d]) + 1j * (gamma[c] + gamma[d]) / 2.0)
                            signal[i, j] += dip[b, a] * dip[c, b] * dip[d, a] * dip[d, c] * U_ba * G_ca * G_cd
    return -1 * signal


def DQC_R2(evals, dip, omega1=None, omega2=[], omega3=None, tau1=None, tau3=None,
           g_idx=[0], e_idx=None, f_idx=None, gamma=None):
    """lime/signal/sos.py:994-1099 (U_ba in the tau1 branch has no -i prefactor, :1056)"""
    a = 0
    if omega3 is None and tau3 is not None:
        signal = np.zeros((len(omega1), len(omega2)), dtype=complex)
        for i in range(len(omega1)):
            pump = omega1[i]
            for j in range(len(omega2)):
                probe = omega2[j]
                for b in e_idx:
                    G_ba = 1. / (pump - (evals[b] - evals[a]) + 1j * (gamma[b] + gamma[a]) / 2.0)
                    for c in f_idx:
                        G_ca = 1. / (probe - (evals[c] - evals[a]) + 1j * (gamma[c] + gamma[a]) / 2.0)
                        for d in e_idx:
                            U_da = -1j * np.exp(-1j * (evals[d] - evals[a]) * tau3 - (gamma[d] + gamma[a]) / 2. * tau3)
                            signal[i, j] += dip[b, a] * dip[c, b] * dip[d, c] * dip[a, d] * G_ba * G_ca * U_da
    elif omega1 is None and tau1 is not None:
        signal = np.zeros((len(omega2), len(omega3)), dtype=complex)
        for i in range(len(omega2)):
            pump = omega2[i]
            for j in range(len(omega3)):
                probe = omega3[j]
                for b in e_idx:
                    U_ba = np.exp(-1j * (evals[b] - evals[a]) * tau1 - (gamma[b] + gamma[a]) / 2. * tau1)
                    for c in f_idx:
                        G_ca = 1. / (pump - (evals[c] - evals[a]) + 1j * (gamma[c] + gamma[a]) / 2.0)
                        for d in e_idx:
                            G_da = 1. / (probe - (evals[d] - evals[a]) + 1j * (gamma[d] + gamma[a]) / 2.0)
                            signal[i, j] += dip[b, a] * dip[c, b] * dip[d, c] * dip[a, d] * U_ba * G_ca * G_da
    else:
        raise Exception('Input Error! Please specify either omega1, tau3 or omega3, tau1.')
    return 1 * signal


def lorentzian(x, width=1.):
    """width is the HWHM, lime/phys.py:669-688"""
    return 1. / np.pi * width / (width ** 2 + (x) ** 2)


def TPA(E, dip, omegap, g_idx, e_idx, f_idx, gamma, degenerate=True):
    """TPA signal with classical light (degenerate pump), lime/signal/sos.py:199-228"""
    if degenerate:
        omega1 = omegap * 0.5
        omega2 = omegap - omega1
    i = 0
    signal = 0
    for f in f_idx:
        tmp = 0.0
        for m in e_idx:
            p1 = dip[f, m] * dip[m, i] / (omega1 - (E[m] - E[i]) + 1j * gamma[m])
            p2 = dip[f, m] * dip[m, i] / (omega2 - (E[m] - E[i]) + 1j * gamma[m])
            tmp += (p1 + p2)
        signal += np.abs(tmp) ** 2 * lorentzian(omegap - E[f] + E[i], width=gamma[f])
    return signal


def TPA2D(E, dip, omegaps, omega1s, g_idx, e_idx, f_idx, gamma):
    """lime/signal/sos.py:230-256 (real output)"""
    g = 0
    signal = np.zeros((len(omegaps), len(omega1s)))
    for i, omegap in enumerate(omegaps):
        for j, omega1 in enumerate(omega1s):
            omega2 = omegap - omega1
            for f in f_idx:
                tmp = 0.
                for m in e_idx:
                    tmp += dip[f, m] * dip[m, g] * (1. / (omega1 - (E[m] - E[g]) + 1j * gamma[m])
                                                    + 1. / (omega2 - (E[m] - E[g]) + 1j * gamma[m]))
                signal[i, j] += np.abs(tmp) ** 2 * lorentzian(omegap - E[f] + E[g], width=gamma[f])
    return signal


def TPA2D_time_order(E, dip, omegaps, omega1s, g_idx, e_idx, f_idx, gamma):
    """lime/signal/sos.py:258-283"""
    g = 0
    signal = np.zeros((len(omegaps), len(omega1s)))
    for i in range(len(omegaps)):
        omegap = omegaps[i]
        for j in range(len(omega1s)):
            omega1 = omega1s[j]
            for f in f_idx:
                tmp = 0.
                for m in e_idx:
                    tmp += dip[f, m] * dip[m, g] * 1. / (omega1 - (E[m] - E[g]) + 1j * gamma[m])
                signal[i, j] += np.abs(tmp) ** 2 * lorentzian(omegap - E[f] + E[g], width=gamma[f])
    return signal


# --------------------------------------------------------------------------
# Liouvillian eigen-decomposition solver         lime/superoperator.py:456-773
# 
# --------------------------------------------------------------------------
def operator_to_vector(rho):
    """row-major flatten, lime/superoperator.py:112-129"""
    if issparse(rho):
        return rho.toarray().flatten()
    return np.asarray(rho).flatten()


class SuperLindblad:
    """lime.superoperator.Lindblad_solver: dense scipy.linalg.eig of L, exponential series"""

    def __init__(self, H, c_ops=None):
        self.H = H
        self.c_ops = c_ops
        self.dim = H.shape[-1] ** 2
        self.idv = operator_to_vector(np.identity(H.shape[-1]))
        self.L = None

    def eigenstates(self):
        """:490-523 (k=None branch); norm = Re diag(vl^H vr) as in lime (:508)"""
        self.L = liouvillian_super(self.H, self.c_ops)
        w, vl, vr = scipy.linalg.eig(self.L.toarray(), left=True, right=True)
        self.eigvals, self.left_eigvecs, self.right_eigvecs = w, vl, vr
        self.norm = np.diagonal(vl.conj().T.dot(vr)).real
        return w, vr, vl

    def evolve(self, rho0, tlist, e_ops):
        """:525-562"""
        evals, U1, U2, norm = self.eigvals, self.right_eigvecs, self.left_eigvecs, self.norm
        rho0 = operator_to_vector(rho0)
        k = U1.shape[-1]
        observables = np.zeros((len(tlist), len(e_ops)), dtype=complex)
        coeff = [np.vdot(U2[:, n], rho0) / norm[n] for n in range(k)]
        for i, t in enumerate(tlist):
            rho = U1.dot(coeff * np.exp(evals * t))
            observables[i, :] = [np.vdot(operator_to_vector(dag(e)), rho) for e in e_ops]
        return observables

    def _coeff1(self, bra_op, ket_vec):
        evals, U1, U2, norm = self.eigvals, self.right_eigvecs, self.left_eigvecs, self.norm
        k = U1.shape[-1]
        return np.array([np.vdot(self.idv, left(bra_op).dot(U1[:, n])) * np.vdot(U2[:, n], ket_vec) / norm[n]
                         for n in range(k)])

    def correlation_2op_1t(self, rho0, ops, tlist):
        """<A(t)B>, :566-605"""
        a, b = ops
        coeff = self._coeff1(a, operator_to_vector(b.dot(rho0)))
        return np.array([np.sum(np.exp(self.eigvals * t) * coeff) for t in tlist])

    def correlation_2op_1w(self, rho0, ops, w):
        """:607-641"""
        a, b = ops
        coeff = self._coeff1(a, operator_to_vector(b.dot(rho0)))
        return np.array([np.sum(-1. / (self.eigvals + 1j * wi) * coeff) for wi in w])

    def correlation_3op_1t(self, rho0, ops, t):
        """<A B(t) C>, :643-671"""
        a, b, c = ops
        coeff = self._coeff1(b, operator_to_vector(c @ rho0 @ a))
        return np.array([np.sum(np.exp(self.eigvals * ti) * coeff) for ti in t])

    def correlation_3op_1w(self, rho0, ops, w):
        """:673-701"""
        a, b, c = ops
        coeff = self._coeff1(b, operator_to_vector(c @ rho0 @ a))
        return np.array([np.sum(-1. / (self.eigvals + 1j * wi) * coeff) for wi in w])

    def correlation_3op_2t(self, rho0, ops, tlist, taulist):
        """<A(t)B(t+tau)C(t)>, :703-754; returns (len(taulist), len(tlist)) like lime"""
        a, b, c = ops
        rho0 = operator_to_vector(rho0)
        evals, U1, U2, norm, idv = self.eigvals, self.right_eigvecs, self.left_eigvecs, self.norm, self.idv
        k = self.dim
        coeff = np.zeros((k, k), dtype=complex)
        for m in range(k):
            for n in range(k):
                coeff[m, n] = np.vdot(idv, left(b).dot(U1[:, m])) * \
                    np.vdot(U2[:, m], right(a).dot(left(c).dot(U1[:, n]))) / norm[m] \
                    * np.vdot(U2[:, n], rho0) / norm[n]
        tmp1 = np.exp(np.outer(evals, taulist))
        tmp2 = np.exp(np.outer(evals, tlist))
        return tmp1.T @ coeff @ tmp2

    def correlation_4op_2t(self, rho0, ops, tlist, taulist):
        """:756-773"""
        if len(ops) != 4:
            raise ValueError('Number of operators is not 4.')
        a, b, c, d = ops
        return self.correlation_3op_2t(rho0, [a, b @ c, d], tlist, taulist)


# --------------------------------------------------------------------------
# spectral post-processing                                   lime/fft.py
# 
# --------------------------------------------------------------------------
def fft(f, x=None, axis=-1, **kwargs):
    """lime/fft.py:15-58"""
    nx = np.asarray(f).shape[axis]
    if x is None:
        x = np.arange(nx)
    dx = x[1] - x[0]
    g = np.fft.fft(f, axis=axis, **kwargs)
    g = np.fft.fftshift(g, axes=(axis, ))
    g *= dx
    freq = 2. * np.pi * np.fft.fftshift(np.fft.fftfreq(nx, d=dx))
    g *= np.exp(-1j * freq * x[0])
    return g, freq


def ifft(f, x=None, axis=-1):
    """lime/fft.py:61-86"""
    nx = np.asarray(f).shape[axis]
    if x is None:
        x = np.arange(nx)
    dx = x[1] - x[0]
    g = np.fft.ifft(f, axis=axis)
    g = np.fft.ifftshift(g)
    g = g * dx / 2. / np.pi * len(x)
    freq = 2. * np.pi * np.fft.ifftshift(np.fft.fftfreq(nx, d=dx))
    return g * np.exp(1j * freq * x[0]), freq


def fft2(f, dx=1, dy=1):
    """lime/fft.py:88-110"""
    nx, ny = f.shape
    g = np.fft.fft2(f)
    g = np.fft.fftshift(g)
    g = g * dx * dy
    freqx = 2. * np.pi * np.fft.fftshift(np.fft.fftfreq(nx, d=dx))
    freqy = 2. * np.pi * np.fft.fftshift(np.fft.fftfreq(nx, d=dy))
    return freqx, freqy, g


def dft(x, f, k):
    """lime/fft.py:112-124 (without the figure)"""
    dx = (x[1] - x[0]).real
    g = np.zeros(len(k), dtype=np.complex128)
    for i in range(len(k)):
        g[i] = np.sum(f * np.exp(-1j * k[i] * x)) * dx
    return g


def dft2(x, y, f, kx, ky):
    """lime/fft.py:126-137"""
    dx = x[1] - x[0]
    dy = y[1] - y[0]
    X, Y = np.meshgrid(x, y)
    g = np.zeros((len(kx), len(ky)), dtype=complex)
    for i in range(len(kx)):
        for j in range(len(ky)):
            g[i, j] = np.sum(f * np.exp(-1j * kx[i] * X - 1j * ky[j] * Y)) * dx * dy
    return g


# --------------------------------------------------------------------------
# wave-function solver                                lime/mol.py:1094-1391
# 
# --------------------------------------------------------------------------
def tdse(wf, h):
    """lime/phys.py:902-903"""
    return -1j * h.dot(wf)


def obs_psi(psi, a):
    """<psi|a|psi>, lime/phys.py:846-862"""
    return np.vdot(psi, a.dot(psi))


def quantum_dynamics(H, psi0, dt=0.001, Nt=1, e_ops=[], t0=0.0, nout=1):
    """_quantum_dynamics with store_states=True, lime/mol.py:1304-1372: returns (observables
    [Nt//nout, E], psilist [Nt//nout]); entry 0 is the initial state"""
    psi = psi0.copy()
    observables = np.zeros((Nt // nout, len(e_ops)), dtype=complex)
    psilist = [psi0]
    observables[0, :] = [obs_psi(psi, e) for e in e_ops]
    for k1 in range(1, Nt // nout):
        for k2 in range(nout):
            psi = rk4(psi, tdse, dt, H)
        observables[k1, :] = [obs_psi(psi, e) for e in e_ops]
        psilist.append(psi.copy())
    return observables, psilist


def driven_dynamics(H, psi0, dt=0.001, Nt=1, e_ops=None, nout=1, t0=0.0, strict=False):
    """laser-driven wave-function dynamics (return_result=True), lime/mol.py:1473-1560: H = [H0, [H1, f1], ...],
    Ht = H0 - sum_i f_i(t) H_i at the start time of the current block of nout steps.  lime only runs with scipy.sparse
    operands here (its psi is a CSR column; an ndarray H raises), and for those `Ht = H[0]; Ht += ...` REBINDS Ht, so the
    drive does not accumulate; strict=True models the accumulation an ndarray H[0] would see (unreachable in lime)."""
    e_ops = [] if e_ops is None else e_ops
    psi = np.array(psi0, dtype=complex)
    H0 = np.array(H[0].toarray() if issparse(H[0]) else H[0], dtype=complex)

    def calcH(t):
        nonlocal H0
        Ht = H0 if strict else H0.copy()
        for i in range(1, len(H)):
            Hi = H[i][0].toarray() if issparse(H[i][0]) else np.asarray(H[i][0])
            Ht += -H[i][1](t) * Hi
        return Ht
    observables = np.zeros((Nt // nout, len(e_ops)), dtype=complex)
    psilist = [np.array(psi0)]
    t = t0
    observables[0, :] = [obs_psi(psi, e) for e in e_ops]
    for k1 in range(1, Nt // nout):
        for k2 in range(nout):
            ht = calcH(t)            # lime evaluates at the block's start time for every step of the block (t += dt*nout after)
            psi = rk4(psi, tdse, dt, ht)
        t += dt * nout
        observables[k1, :] = [obs_psi(psi, e) for e in e_ops]
        psilist.append(psi.copy())
    return observables, psilist


def se_correlation_3op_1t(H, psi0, oplist, dt, Nt):
    """<A B(t) C>, lime/mol.py:1185-1211"""
    a_op, b_op, c_op = oplist
    ket = quantum_dynamics(H, c_op @ psi0, dt=dt, Nt=Nt)[1]
    bra = quantum_dynamics(H, dag(a_op) @ psi0, dt=dt, Nt=Nt)[1]
    return np.array([np.vdot(bra[j], b_op @ ket[j]) for j in range(Nt)])


def se_correlation_3op_2t(H, psi0, oplist, dt, Nt, Ntau):
    """<A(t) B(t+tau) C(t)>, lime/mol.py:1213-1246"""
    psi_t = quantum_dynamics(H, psi0, dt=dt, Nt=Nt)[1]
    a_op, b_op, c_op = oplist
    corr = np.zeros([Nt, Ntau], dtype=complex)
    for i, psi in enumerate(psi_t):
        ket = quantum_dynamics(H, c_op @ psi, dt=dt, Nt=Ntau)[1]
        bra = quantum_dynamics(H, dag(a_op) @ psi, dt=dt, Nt=Ntau)[1]
        corr[i, :] = [np.vdot(bra[j], b_op @ ket[j]) for j in range(Ntau)]
    return corr


# --------------------------------------------------------------------------
# time-domain third-order response functions              lime/signal/2DES.py
# (the module itself is not importable -- it executes undefined names at :249-263 --
#  but the functions :37-247 are; oracle/gen_golden.py execs exactly those lines of the
#  reference source to pin this restatement)
# --------------------------------------------------------------------------
def td_G(en, decay, a, b, t):
    """-i theta(t) exp(-i (E_a - E_b) t - (g_a + g_b)/2 t), lime/signal/2DES.py:37-60
    (`en`, `decay` are module globals there)"""
    return -1j * np.heaviside(t, 1) * np.exp(-1j * (en[a] - en[b]) * t - (decay[a] + decay[b]) / 2. * t)


def td_ESA(en, decay, dip, g_idx, e_idx, f_idx, t1, t2, t3):
    """lime/signal/2DES.py:99-155: gg -> ge -> e'e -> fe -> ee, sign -1"""
    signal = 0
    a = 0
    for b in e_idx:
        G_ab = td_G(en, decay, a, b, t1)
        for c in e_idx:
            G_cb = td_G(en, decay, c, b, t2)
            for d in f_idx:
                G_db = td_G(en, decay, d, b, t3)
                signal += dip[b, a] * dip[c, a] * dip[d, c] * dip[b, d] * G_db * G_cb * G_ab
    return -1 * signal


def td_GSB(en, decay, dip, g_idx, e_idx, t1, t2, t3):
    """lime/signal/2DES.py:158-203: gg -> ge -> gg' -> e'g' -> g'g'"""
    signal = 0
    a = 0
    for b in e_idx:
        G_ab = td_G(en, decay, a, b, t1)
        for c in g_idx:
            G_ac = td_G(en, decay, a, c, t2)
            for d in e_idx:
                G_dc = td_G(en, decay, d, c, t3)
                signal += dip[a, b] * dip[b, c] * dip[c, d] * dip[d, a] * G_dc * G_ac * G_ab
    return signal


def td_SE(en, decay, dip, g_idx, e_idx, t1, t2, t3):
    """lime/signal/2DES.py:207-247: gg -> ge -> e'e -> g'e -> g'g'"""
    signal = 0.0
    a = 0
    for b in e_idx:
        G_ab = td_G(en, decay, a, b, t1)
        for c in e_idx:
            G_cb = td_G(en, decay, c, b, t2)
            for d in g_idx:
                G_cd = td_G(en, decay, c, d, t3)
                signal += dip[a, b] * dip[c, a] * dip[d, c] * dip[b, d] * G_cd * G_cb * G_ab
    return signal


# --------------------------------------------------------------------------
# model builders used by the BASELINE configs (host-side in lime too)
# --------------------------------------------------------------------------
def jaynes_cummings(omega0, omegac, g, ncav, kappa, rwa=False):
    """Config-2 operators: 2-level molecule (x) ncav-level cavity, index =
    i_mol*ncav + n (kron(mol, cav), lime/cavity.py:76).  H_mol = omega0/2 (1 - sz),
    H_cav = omegac a^dag a (lime/phys.py:789-806 without ZPE), coupling
    g sx (x) (a + a^dag) (lime/cavity.py:57-97) or its RWA part; c_op = sqrt(kappa) 1 (x) a.
    Returns CSR (H, [c_op], [a^dag a, sigma^+ sigma^-])."""
    a = destroy(ncav)
    num = csr_matrix(np.diag(np.arange(ncav, dtype=float)))
    s0, sx, sy, sz = pauli()
    hmol = csr_matrix(0.5 * omega0 * (s0 - sz))
    ic = identity(ncav, format='csr')
    im = identity(2, format='csr')
    H = kron(hmol, ic) + kron(im, omegac * num)
    if rwa:
        sm = csr_matrix(np.array([[0., 1.], [0., 0.]]))     # |g><e|, g = index 0
        H = H + g * (kron(sm.T, a) + kron(sm, a.T))
    else:
        H = H + g * kron(csr_matrix(sx), a + a.T)
    c = np.sqrt(kappa) * kron(im, a)
    pe = csr_matrix(np.array([[0., 0.], [0., 1.]]))
    return csr_matrix(H), [csr_matrix(c)], [csr_matrix(kron(im, num)), csr_matrix(kron(pe, ic))]
