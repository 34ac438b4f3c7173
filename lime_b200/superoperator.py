"""
Liouville-space algebra (row-major vec; kron(a, I) is left multiplication) with the
semantics of lime/superoperator.py:30-59,131-151,201-271.  Set-up only: these build the
sparse superoperators that the device propagators consume.
"""
import numpy as np
from scipy.sparse import identity, kron, issparse

from .phys import dag


def operator_to_vector(rho):
    """lime/superoperator.py:112-129"""
    if isinstance(rho, np.ndarray):
        return rho.flatten()
    return rho.toarray().flatten()


def dm2vec(rho):
    """lime/superoperator.py:131-151"""
    if issparse(rho):
        n, m = rho.shape
        return rho.tolil().reshape((n * m, 1))
    return rho.flatten()


def operator_to_superoperator(a, kind='commutator'):
    """lime/superoperator.py:201-247"""
    N = a.shape[-1]
    idm = identity(N)
    if kind in ['commutator', 'c', '-']:
        return kron(a, idm) - kron(idm, a.T)
    elif kind in ['left', 'l']:
        return kron(a, idm)
    elif kind in ['right', 'r']:
        return kron(idm, a.T)
    elif kind in ['anticommutator', 'a', '+']:
        return kron(a, idm) + kron(idm, a.T)
    raise ValueError('Error: superoperator {} does not exist.'.format(kind))


def op2sop(a, kind='commutator'):
    return operator_to_superoperator(a, kind=kind)


to_super = op2sop


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


def liouvillian(H, c_ops):
    """lime/superoperator.py:30-59"""
    if c_ops is None:
        c_ops = []
    l = -1j * operator_to_superoperator(H)
    for c_op in c_ops:
        l = l + lindblad_dissipator(c_op)
    return l


def kraus(a):
    """lime/superoperator.py:273-291"""
    return right(dag(a)).dot(left(a))


def obs(rho, a):
    """lime/superoperator.py:314-315"""
    return np.vdot(operator_to_vector(dag(a)), rho)


def cdot(a, b):
    """a^H b, lime/superoperator.py:371-390"""
    return dag(a).dot(b)


class Lindblad_solver:
    """Liouville-space solver by full eigen-decomposition of the Liouvillian, lime/superoperator.py:456-773.

    The decomposition itself is the host LAPACK call lime makes (scipy.linalg.eig with left and right vectors,
    norm = Re diag(vl^H vr) as at :508) -- so the eigenvectors, their phases and the `.real` quirk are lime's.
    Everything after it that is O(k^2) or worse in k = N^2 runs on the device as plain complex GEMMs on the FP64
    tensor cores (limeb200_zgemm): the exponential series of `evolve` (U1 @ E) and the two-time function
    (coeff = (U2 diag(conj a))^H (R_a L_c U1 diag(beta)), cor = tmp1.T @ coeff @ tmp2), which lime builds with an
    O(k^2) Python loop of sparse mat-vecs (:734-741).  `eigenstates(k=...)` with k given is broken in lime
    (returns an undefined name, :518-523) and raises NotImplementedError here."""

    def __init__(self, H, c_ops=None):
        self.H = H
        self.c_ops = c_ops
        self.L = None
        self.dim = H.shape[-1] ** 2
        self.idv = operator_to_vector(np.identity(H.shape[-1]))
        self.left_eigvecs = None
        self.right_eigvecs = None
        self.eigvals = None
        self.norm = None

    def liouvillian(self):
        self.L = liouvillian(self.H, self.c_ops)
        return self.L

    def eigenstates(self, k=None):
        import scipy.linalg
        L = self.liouvillian() if self.L is None else self.L
        if k is not None:
            raise NotImplementedError('eigenstates(k=...) is not functional in lime (superoperator.py:510-523)')
        w, vl, vr = scipy.linalg.eig(L.toarray(), left=True, right=True)
        self.eigvals, self.left_eigvecs, self.right_eigvecs = w, vl, vr
        self.norm = np.diagonal(cdot(vl, vr)).real
        return w, vr, vl

    def _need(self):
        if self.eigvals is None:
            raise TypeError("eigenstates() has not been called")      # lime fails on None here as well

    def evolve(self, rho0, tlist, e_ops):
        """observables[i, :] = <e_op>(t_i) from the exponential series; returns a Result (:525-562)"""
        from .mol import Result
        from . import engine
        self._need()
        tlist = np.asarray(tlist, dtype=float)
        evals, U1, U2, norm = self.eigvals, self.right_eigvecs, self.left_eigvecs, self.norm
        coeff = (U2.conj().T @ operator_to_vector(np.asarray(rho0))) / norm
        E = coeff[:, None] * np.exp(evals[:, None] * tlist[None, :])           # [k, Nt]
        rho_t = engine.zgemm(np.ascontiguousarray(U1), np.ascontiguousarray(E))  # device [k, Nt]
        if len(e_ops):
            ev = np.ascontiguousarray(np.stack([np.conj(operator_to_vector(dag(np.asarray(e)))) for e in e_ops]))
            obs_t = engine.zgemm(ev, rho_t).cpu().numpy().T                     # [Nt, E]
        else:
            obs_t = np.zeros((len(tlist), 0), dtype=complex)
        result = Result(times=tlist, dt=(tlist[1] - tlist[0]) if len(tlist) > 1 else 0.0, Nt=len(tlist))
        result.times = tlist
        result.observables = obs_t
        return result

    def _coeff1(self, bra_op, ket_vec):
        evals, U1, U2, norm = self.eigvals, self.right_eigvecs, self.left_eigvecs, self.norm
        return (np.conj(self.idv) @ left(bra_op).dot(U1)) * (U2.conj().T @ ket_vec) / norm

    def correlation_2op_1t(self, rho0, ops, tlist):
        """<A(t)B>, :566-605 (O(k Nt): host)"""
        self._need()
        a, b = ops
        coeff = self._coeff1(a, operator_to_vector(b.dot(rho0)))
        return np.exp(np.outer(np.asarray(tlist), self.eigvals)) @ coeff

    def correlation_2op_1w(self, rho0, ops, w):
        """:607-641"""
        self._need()
        a, b = ops
        coeff = self._coeff1(a, operator_to_vector(b.dot(rho0)))
        return (-1. / (self.eigvals[None, :] + 1j * np.asarray(w)[:, None])) @ coeff

    def correlation_3op_1t(self, rho0, ops, t):
        """<A B(t) C>, :643-671"""
        self._need()
        a, b, c = ops
        coeff = self._coeff1(b, operator_to_vector(c @ rho0 @ a))
        return np.exp(np.outer(np.asarray(t), self.eigvals)) @ coeff

    def correlation_3op_1w(self, rho0, ops, w):
        """:673-701"""
        self._need()
        a, b, c = ops
        coeff = self._coeff1(b, operator_to_vector(c @ rho0 @ a))
        return (-1. / (self.eigvals[None, :] + 1j * np.asarray(w)[:, None])) @ coeff

    def correlation_3op_2t(self, rho0, ops, tlist, taulist, k=None):
        """<A(t)B(t+tau)C(t)>, :703-754.  Returns (len(taulist), len(tlist)) as lime does (:749-752)."""
        from . import engine
        self._need()
        a, b, c = ops
        rho0v = operator_to_vector(np.asarray(rho0))
        evals, U1, U2, norm, idv = self.eigvals, self.right_eigvecs, self.left_eigvecs, self.norm, self.idv
        alpha = (np.conj(idv) @ left(b).dot(U1)) / norm                       # [k]  (row factor of coeff)
        beta = (U2.conj().T @ rho0v) / norm                                    # [k]  (column factor)
        X = right(a).dot(left(c).dot(U1))                                      # sparse x dense, [k, k]
        lhs = np.ascontiguousarray((U2 * np.conj(alpha)[None, :]).conj().T)    # diag(alpha) U2^H
        rhs = np.ascontiguousarray(np.asarray(X) * beta[None, :])
        coeff = engine.zgemm(lhs, rhs)                                         # device [k, k]
        tmp1T = np.ascontiguousarray(np.exp(np.outer(np.asarray(taulist), evals)))   # [Ntau, k]
        tmp2 = np.ascontiguousarray(np.exp(np.outer(evals, np.asarray(tlist))))      # [k, Nt]
        left_part = engine.zgemm(tmp1T, coeff)                                 # [Ntau, k]
        return engine.zgemm(left_part, tmp2).cpu().numpy()                     # [Ntau, Nt]

    def correlation_4op_2t(self, rho0, ops, tlist, taulist, k=None):
        """:756-773"""
        if len(ops) != 4:
            raise ValueError('Number of operators is not 4.')
        a, b, c, d = ops
        return self.correlation_3op_2t(rho0=rho0, ops=[a, b @ c, d], tlist=tlist, taulist=taulist, k=k)
