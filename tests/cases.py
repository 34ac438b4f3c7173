"""
Seeded input builders shared by oracle/gen_golden.py (which freezes reference
outputs into tests/golden/) and by the parity tests.  NumPy only.
"""
import numpy as np
from scipy.sparse import csr_matrix, lil_matrix


def rand_herm(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    return scale * (a + a.conj().T) / 2


def rand_cplx(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return scale * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))


def rand_dm(n, seed):
    """random Hermitian, positive, trace-1 matrix"""
    a = rand_cplx(n, seed)
    rho = a @ a.conj().T
    return rho / np.trace(rho)


def rand_dm_batch(B, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((B, n, n)) + 1j * rng.standard_normal((B, n, n))
    rho = a @ a.conj().transpose(0, 2, 1)
    tr = np.einsum('bii->b', rho)
    return rho / tr[:, None, None]


# ---- reference golden: examples/test_cavity.py:70-92 (thermal cavity) --------
def thermal_cavity(N=10, kappa=0.25, n_th=2.0):
    """operators of the run that produced examples/cor.dat / dm.dat:
    H = diag(n), c_ops = [sqrt(kappa(1+n_th)) a, sqrt(kappa n_th) a^dag],
    rho0 = thermal(n_th=2), ops = [a^dag, a^dag a, a], tlist = linspace(0,10,200)"""
    a = lil_matrix((N, N))
    a.setdiag(np.sqrt(np.arange(1, N)), 1)
    a = a.tocsr()
    ad = a.conjugate().transpose().tocsr()
    H = csr_matrix(np.diag(np.arange(N) * 1.0))
    num = csr_matrix(np.diag(np.arange(N) * 1.0))
    c_ops = [np.sqrt(kappa * (1 + n_th)) * a, np.sqrt(kappa * n_th) * ad]
    w = np.array([np.exp(-n / 2) for n in range(N)])
    w /= np.sum(w)
    rho0 = lil_matrix((N, N))
    rho0.setdiag(w)
    rho0 = rho0.tocsr()
    tlist = np.linspace(0, 10, 200)
    return H, rho0, [ad, num, a], c_ops, tlist


# ---- config 1: examples/redfield.py:14-47 -------------------------------------
def redfield_example():
    delta = 0.2 * 2 * np.pi
    eps0 = 1.0 * 2 * np.pi
    gamma1 = 0.5
    sx = np.array([[0., 1.], [1., 0.]])
    sz = np.array([[1., 0.], [0., -1.]])
    H = - delta / 2.0 * sx - eps0 / 2.0 * sz

    def ohmic_spectrum(w):
        return gamma1 / 2 * (w / (2 * np.pi))

    psi0 = np.zeros(2)
    psi0[1] = 1.0
    rho0 = np.einsum("i, j -> ij", psi0, psi0.conj()).astype(complex)
    Nt = 200
    tlist = np.linspace(0, 20, Nt)
    dt = tlist[1] - tlist[0]
    return H, [sx], [ohmic_spectrum], rho0, dt, Nt, [sx], tlist


def redfield_multilevel(n=5, seed=11):
    """physical-ish n-level Redfield problem with two coupling operators"""
    H = rand_herm(n, seed, 1.0).real + np.diag(np.arange(n) * 1.5)
    a1 = rand_herm(n, seed + 1, 0.3)
    a2 = rand_herm(n, seed + 2, 0.2).real

    def s1(w):
        return 0.05 * (1.0 + np.tanh(w))

    def s2(w):
        return 0.02 * np.exp(-abs(w) / 3.0) * (1.0 if w >= 0 else 0.5)

    return H, [a1, a2], [s1, s2], rand_dm(n, seed + 3)


# ---- Lindblad ----------------------------------------------------------------
def lindblad_dense(n=6, M=2, E=2, seed=5):
    H = rand_herm(n, seed)
    c_ops = [rand_cplx(n, seed + 1 + m, 0.2) for m in range(M)]
    e_ops = [rand_herm(n, seed + 20 + e) for e in range(E)]
    rho0 = rand_dm(n, seed + 40)
    return H, c_ops, e_ops, rho0


def jc_point(ncav=8, g=0.1, detuning=0.05, kappa=0.05, omega0=1.0, rwa=False):
    """dense restatement-free JC builder (same layout as lime: index = i_mol*ncav + n)"""
    a = np.diag(np.sqrt(np.arange(1, ncav)), 1)
    num = np.diag(np.arange(ncav) * 1.0)
    sx = np.array([[0., 1.], [1., 0.]])
    sz = np.array([[1., 0.], [0., -1.]])
    s0 = np.identity(2)
    hmol = 0.5 * omega0 * (s0 - sz)
    H = np.kron(hmol, np.identity(ncav)) + np.kron(s0, (omega0 + detuning) * num)
    if rwa:
        sm = np.array([[0., 1.], [0., 0.]])
        H = H + g * (np.kron(sm.T, a) + np.kron(sm, a.T))
    else:
        H = H + g * np.kron(sx, a + a.T)
    c = np.sqrt(kappa) * np.kron(s0, a)
    pe = np.array([[0., 0.], [0., 1.]])
    e_ops = [np.kron(s0, num), np.kron(pe, np.identity(ncav))]
    rho0 = np.zeros((2 * ncav, 2 * ncav), dtype=complex)
    rho0[ncav, ncav] = 1.0          # |e,0><e,0|
    return H, [c], e_ops, rho0


# ---- HEOM --------------------------------------------------------------------
def spin_boson_heom(depth=4, K=2, lam=0.2, gam=1.0, beta=1.0):
    sx = np.array([[0., 1.], [1., 0.]], dtype=complex)
    sz = np.array([[1., 0.], [0., -1.]], dtype=complex)
    H = 0.5 * 1.0 * sz + 0.5 * 0.5 * sx
    rho0 = np.zeros((2, 2), dtype=complex)
    rho0[0, 0] = 1.0
    return H, sz, rho0, depth, K, lam, gam, 1.0 / beta


# ---- sum over states ------------------------------------------------------------
def sos_system(N=8, ne=3, seed=0):
    """E[0]=0, e-manifold around 1.5-2 eV, f-manifold around 3.2-3.8 eV (a.u.)"""
    au2ev = 27.211386
    rng = np.random.default_rng(seed)
    nf = N - 1 - ne
    E = np.zeros(N)
    E[1:1 + ne] = np.sort(rng.uniform(1.5, 2.0, ne)) / au2ev
    E[1 + ne:] = np.sort(rng.uniform(3.2, 3.8, nf)) / au2ev
    dip = np.zeros((N, N))
    ge = rng.standard_normal(ne)
    dip[0, 1:1 + ne] = ge
    dip[1:1 + ne, 0] = ge
    ef = rng.standard_normal((ne, nf))
    dip[1:1 + ne, 1 + ne:] = ef
    dip[1 + ne:, 1:1 + ne] = ef.T
    gamma = np.ones(N) * 0.05 / au2ev
    gamma[0] = 0.0
    g_idx = [0]
    e_idx = list(range(1, 1 + ne))
    f_idx = list(range(1 + ne, N))
    return E, dip, gamma, g_idx, e_idx, f_idx


class DuckMol:
    """the attributes lime's Mol-based wrappers read (lime/signal/sos.py:731-902): eigvals(), edip_rms, gamma,
    dephasing, nstates"""

    def __init__(self, E, dip, gamma, dephasing=0.0):
        self._E, self.edip_rms, self.gamma, self.dephasing = E, dip, gamma, dephasing
        self.nstates = len(E)

    def eigvals(self):
        return self._E
