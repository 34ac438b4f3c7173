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
