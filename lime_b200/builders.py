"""
Host-side builders for the synthetic inputs of the BASELINE configs (in lime these are
Python model classes: lime/cavity.py:26-97 Composite.getH, :380-437 Cavity; they stay on
the host here too).  Index convention of kron(mol, cav): i = i_mol * ncav + n
(lime/cavity.py:76).
"""
import numpy as np
from scipy.sparse import csr_matrix, identity, kron, diags

from .units import au2fs, au2k, au2wavenumber


def jaynes_cummings_ops(ncav, kappa):
    """(A, Bm, Cm, c_ops, e_ops, rho0) with H = omega0*A + omegac*Bm + g*Cm,
    A = sigma+sigma- (x) 1, Bm = 1 (x) a^dag a, Cm = sigma_x (x) (a + a^dag) (no RWA, as
    Composite.getH([sx],[a+a^dag],[g]) builds it); c_op = sqrt(kappa) 1 (x) a;
    e_ops = [a^dag a, sigma+sigma-]; rho0 = |e,0><e,0|."""
    a = diags(np.sqrt(np.arange(1, ncav)), 1, format='csr')
    num = diags(np.arange(ncav, dtype=float), 0, format='csr')
    sx = csr_matrix(np.array([[0., 1.], [1., 0.]]))
    pe = csr_matrix(np.array([[0., 0.], [0., 1.]]))
    i2, ic = identity(2, format='csr'), identity(ncav, format='csr')
    A = csr_matrix(kron(pe, ic))
    Bm = csr_matrix(kron(i2, num))
    Cm = csr_matrix(kron(sx, a + a.T))
    c_ops = [csr_matrix(np.sqrt(kappa) * kron(i2, a))]
    e_ops = [csr_matrix(kron(i2, num)), A]
    N = 2 * ncav
    rho0 = np.zeros((N, N), dtype=np.complex128)
    rho0[ncav, ncav] = 1.0
    return A, Bm, Cm, c_ops, e_ops, rho0


def jaynes_cummings_batch(omega0, omegac, g, ncav, kappa):
    """Hamiltonian batch in (pattern, values[B,nnz]) form for B = len(g) parameter points."""
    A, Bm, Cm, c_ops, e_ops, rho0 = jaynes_cummings_ops(ncav, kappa)
    pat = csr_matrix(abs(A) + abs(Bm) + abs(Cm) + identity(2 * ncav, format='csr'))
    pat.sum_duplicates()
    pat.sort_indices()
    N = 2 * ncav
    rows = np.repeat(np.arange(N), np.diff(pat.indptr))
    cols = pat.indices

    def onto(m):
        return np.asarray(m.todense())[rows, cols]
    va, vb, vc = onto(A), onto(Bm), onto(Cm)
    omegac = np.asarray(omegac, dtype=float).reshape(-1)
    g = np.asarray(g, dtype=float).reshape(-1)
    vals = omega0 * va[None, :] + omegac[:, None] * vb[None, :] + g[:, None] * vc[None, :]
    pat.data[:] = 1.0
    return pat, vals.astype(np.complex128), c_ops, e_ops, rho0


def jc_grid(ncav=64, ng=64, ndet=64, omega0=1.0, Q=None, kappa=0.05):
    """config 2: 64 couplings g in linspace(0.01,0.2)*omega0 x 64 detunings in
    linspace(-0.2,0.2)*omega0 -> 4096 points"""
    gs = np.linspace(0.01, 0.2, ng) * omega0
    dets = np.linspace(-0.2, 0.2, ndet) * omega0
    G, Dt = np.meshgrid(gs, dets, indexing='ij')
    return jaynes_cummings_batch(omega0, omega0 + Dt.reshape(-1), G.reshape(-1), ncav, kappa)


# FMO 7-site Hamiltonian (Adolphs & Renger 2006, cm^-1), the customary HEOM benchmark
FMO_CM = np.array([
    [200., -87.7, 5.5, -5.9, 6.7, -13.7, -9.9],
    [-87.7, 320., 30.8, 8.2, 0.7, 11.8, 4.3],
    [5.5, 30.8, 0., -53.5, -2.2, -9.6, 6.0],
    [-5.9, 8.2, -53.5, 110., -70.7, -17.0, -63.3],
    [6.7, 0.7, -2.2, -70.7, 270., 81.1, -1.3],
    [-13.7, 11.8, -9.6, -17.0, 81.1, 420., 39.7],
    [-9.9, 4.3, 6.0, -63.3, -1.3, 39.7, 230.]])


def fmo_heom_inputs(temperature_K=300.0, reorg_cm=35.0, tau_fs=50.0):
    """config 4: (H [a.u.], Q list of 7 site projectors, lambda, gamma, kT) in atomic units"""
    H = FMO_CM / au2wavenumber
    Q = [np.diag((np.arange(7) == j).astype(float)) for j in range(7)]
    lam = reorg_cm / au2wavenumber
    gam = 1.0 / (tau_fs / au2fs)
    kT = temperature_K / au2k
    return H.astype(np.complex128), Q, lam, gam, kT
