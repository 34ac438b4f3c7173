"""
L1 primitives of lime/phys.py that sit on the density-matrix path.

The algebraic helpers (comm, dag, transform ...) are host-side conveniences with lime's
exact semantics (they are used for set-up: building generators, changing basis).  The two
functions lime's solvers actually spend their time in -- `liouvillian` and `rk4` applied to
it -- are served by the CUDA engine: see lime_b200.oqs.
"""
import numpy as np
from scipy.sparse import lil_matrix, issparse


def dag(a):
    """lime/phys.py:758-759"""
    return a.conjugate().transpose()


dagger = dag


_DEVICE_MIN_N = 64      # dense products from this size on run as ONE batched FP64 tensor-core GEMM (A.B and B.A together)


def _pair_products(A, B):
    """(A.B, B.A) for square dense ndarrays: both products in one batched limeb200_zgemm launch when the matrices are
    large enough to matter; sparse operands and the small matrices of model set-up stay NumPy/SciPy, as in lime"""
    if (isinstance(A, np.ndarray) and isinstance(B, np.ndarray) and A.ndim == 2 and A.shape[0] == A.shape[1]
            and A.shape[0] >= _DEVICE_MIN_N):
        from . import engine
        a = np.ascontiguousarray(A, dtype=np.complex128)
        b = np.ascontiguousarray(B, dtype=np.complex128)
        out = engine.zgemm(np.stack([a, b]), np.stack([b, a])).cpu().numpy()
        if not (np.iscomplexobj(A) or np.iscomplexobj(B)):
            out = out.real
        return out[0], out[1]
    return None


def comm(A, B):
    """lime/phys.py:741-743"""
    assert A.shape == B.shape
    p = _pair_products(A, B)
    if p is not None:
        return p[0] - p[1]
    return np.dot(A, B) - np.dot(B, A)


def anticomm(A, B):
    """lime/phys.py:746-748"""
    assert A.shape == B.shape
    p = _pair_products(A, B)
    if p is not None:
        return p[0] + p[1]
    return np.dot(A, B) + np.dot(B, A)


def commutator(A, B):
    """lime/phys.py:736-738"""
    assert A.shape == B.shape
    p = _pair_products(A, B)
    if p is not None:
        return p[0] - p[1]
    return A.dot(B) - B.dot(A)


def anticommutator(A, B):
    """lime/phys.py:750-752"""
    assert A.shape == B.shape
    p = _pair_products(A, B)
    if p is not None:
        return p[0] + p[1]
    return A.dot(B) + B.dot(A)


def transform(A, v):
    """v^dag A v, lime/phys.py:706-718"""
    return dag(v).dot(A.dot(v))


def obs_dm(rho, d):
    """Tr(d rho), lime/phys.py:837-844 (O(N^2) form of the same number)"""
    if issparse(d) or issparse(rho):
        return d.dot(rho).diagonal().sum()
    return np.sum(np.asarray(d).T * np.asarray(rho))


def isherm(a):
    """lime/phys.py:1429"""
    return np.allclose(a, dag(a))


def pauli():
    """lime/phys.py:773-786"""
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
    """lime/phys.py:615-633"""
    a = lil_matrix((N, N))
    a.setdiag(np.sqrt(np.arange(1, N)), 1)
    return a.tocsr()


def lorentzian(x, width=1.):
    """lime/phys.py:669-688"""
    return 1. / np.pi * width / (width ** 2 + (x) ** 2)


def rk4(rho, fun, dt, *args):
    """lime/phys.py:636-649: one classical RK4 step; updates `rho` IN PLACE and returns the same object.

    When `fun` is this package's `liouvillian` (args = H, c_ops) -- the combination lime's `_lindblad` loops over,
    lime/oqs.py:1676 -- the step is one launch of the fused CUDA propagator.  For any other Python right-hand side
    this is lime's generic driver: the four evaluations are whatever `fun` does (e.g. `oqs.liouvillian` evaluates
    on the device); only the axpy algebra of an arbitrary user callback is NumPy.  None of the solvers in this
    package loops over this function: their whole time loop runs inside one CUDA launch."""
    from . import oqs
    if fun is oqs.liouvillian and len(args) == 2 and isinstance(rho, np.ndarray):
        plan = oqs._lindblad_plan_cached(args[0], args[1], None)      # uploaded once, reused by every step of a user loop
        out, _, _ = plan.run(rho, dt, 1)
        rho[...] = out
        return rho
    dt2 = dt / 2.0
    k1 = fun(rho, *args)
    k2 = fun(rho + k1 * dt2, *args)
    k3 = fun(rho + k2 * dt2, *args)
    k4 = fun(rho + k3 * dt, *args)
    rho += (k1 + 2 * k2 + 2 * k3 + k4) / 6. * dt
    return rho
