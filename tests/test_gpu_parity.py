"""GPU parity: the CUDA path (through the C ABI / lime-compatible Python surface) against the
NumPy oracle on the same seeded inputs and against the frozen reference outputs in tests/golden.
Tolerance: <= 1e-10 relative in complex128 (BASELINE.json north_star); integer tables bit-exact."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix

import cases
import lime_oracle as lo
from conftest import golden, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


# ---------------------------------------------------------------- Lindblad
def test_reference_golden_cavity(cuda):
    """examples/cor.dat + dm.dat: <a^dag (a^dag a)(t) a>, thermal cavity N=10, CSR operands"""
    from lime_b200.oqs import Lindblad_solver
    g = golden('cavity_cor')
    H, rho0, ops, c_ops, tlist = cases.thermal_cavity()
    dt = tlist[1] - tlist[0]
    A, B, C = ops
    s = Lindblad_solver(H, c_ops=c_ops)
    res = s.evolve(C.dot(rho0.dot(A)).toarray(), dt, len(tlist), e_ops=[B])
    assert relerr(res.observables[:, 0], g['cor']) <= TOL
    assert relerr(np.array(res.rholist), g['dm']) <= TOL
    cor = s.correlation_3op_1t(rho0.toarray(), [A.toarray(), B.toarray(), C.toarray()], dt=dt, Nt=len(tlist))
    assert relerr(cor, g['cor']) <= TOL


@pytest.mark.parametrize('path', [0, 1, 2])
def test_lindblad_dense_paths(cuda, path):
    from lime_b200 import oqs
    g = golden('lindblad_dense')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense()
    plan = oqs._lindblad_plan(H, c_ops, e_ops, path=path)
    rho_f, obs, traj = plan.run(rho0, 0.01, 80, traj_every=1)
    assert relerr(obs, g['observables']) <= TOL
    assert relerr(traj, g['rholist']) <= TOL
    assert relerr(rho_f, g['rholist'][-1]) <= TOL
    assert relerr(plan.rhs(rho0), g['rhs']) <= 1e-13
    assert plan.last_launches >= 1


def test_lindblad_drop_in_surface(cuda):
    from lime_b200 import oqs
    g = golden('lindblad_dense')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense()
    rho_in = rho0.copy()
    res = oqs._lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=80, dt=0.01)
    assert np.array_equal(rho0, rho_in)                       # input is copied, never mutated
    assert res.observables.shape == (80, 2) and res.observables.dtype == np.complex128
    assert len(res.rholist) == 80 and res.rholist[0].shape == (6, 6)
    assert relerr(res.observables, g['observables']) <= TOL
    assert relerr(oqs.liouvillian(rho0, H, c_ops), g['rhs']) <= 1e-13
    ref = lo.lindbladian(c_ops[0], rho0)
    assert relerr(oqs.lindbladian(c_ops[0], rho0), ref) <= 1e-13
    res2 = oqs.Lindblad_solver(H, c_ops).evolve(rho0, 0.01, 80, e_ops=e_ops)
    assert relerr(res2.observables, g['observables']) <= TOL
    # rk4(rho, liouvillian, dt, H, c_ops): one fused device step, in place, same object returned (lime/phys.py:636-649)
    from lime_b200 import phys
    r = rho0.copy()
    ref = lo.rk4(rho0.copy(), lo.liouvillian, 0.01, H, c_ops)
    assert phys.rk4(r, oqs.liouvillian, 0.01, H, c_ops) is r and relerr(r, ref) <= 1e-13
    r = rho0.copy()          # generic driver with a user callback whose evaluations run on the device
    assert relerr(phys.rk4(r, lambda x, h, c: oqs.liouvillian(x, h, c), 0.01, H, c_ops), ref) <= 1e-13
    # no e_ops / no c_ops edge cases
    r3 = oqs._lindblad(H, rho0, [], e_ops=None, Nt=5, dt=0.01)
    o3, l3 = lo.lindblad(H, rho0, [], None, Nt=5, dt=0.01)
    assert r3.observables.shape == (5, 0) and relerr(r3.rholist[-1], l3[-1]) <= TOL


@pytest.mark.parametrize('path', [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize('ncav', [8, 16, 32, 37, 64])
def test_lindblad_jc_every_kernel(cuda, path, ncav):
    """Jaynes-Cummings (config-2 shape, small cutoffs) through every kernel family"""
    from lime_b200 import oqs
    H, c_ops, e_ops, rho0 = cases.jc_point(ncav=ncav)
    obs_o, rl_o = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=60, dt=0.01)
    if path == 1 and 2 * ncav > 64:
        pytest.skip('dense on-chip path is N <= 64')
    sp = path in (3, 4, 5, 6)
    Hs = csr_matrix(H) if sp else H
    cs = [csr_matrix(c) for c in c_ops] if sp else c_ops
    plan = oqs._lindblad_plan(Hs, cs, e_ops, path=path)
    assert plan.path == path
    rho_f, obs, traj = plan.run(rho0, 0.01, 60, traj_every=20)
    assert relerr(obs, obs_o) <= TOL
    assert relerr(rho_f, rl_o[-1]) <= TOL
    assert relerr(traj, np.array([rl_o[19], rl_o[39], rl_o[59]])) <= TOL
    if ncav == 8:
        g = golden('lindblad_jc')
        p2 = oqs._lindblad_plan(Hs, cs, e_ops, path=path)
        rf, ob, _ = p2.run(rho0, 0.01, 100)
        assert relerr(ob, g['observables']) <= TOL and relerr(rf, g['rho_final']) <= TOL


@pytest.mark.parametrize('path', [1, 3, 4, 5])
def test_lindblad_batch_parameter_scan(cuda, path):
    """[ext] batch over coupling/detuning points with per-point Hamiltonian values"""
    from lime_b200.oqs import Lindblad_solver
    pts = [(0.05, -0.1), (0.1, 0.0), (0.15, 0.1), (0.2, 0.2), (0.08, 0.05)]
    Hs, refs = [], []
    for gc, det in pts:
        H, c_ops, e_ops, rho0 = cases.jc_point(ncav=12, g=gc, detuning=det)
        Hs.append(csr_matrix(H) if path != 1 else H)
        refs.append(lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=40, dt=0.02))
    cs = [csr_matrix(c) for c in c_ops] if path != 1 else c_ops
    s = Lindblad_solver(None, c_ops=cs)
    rho_f, obs, _ = s.evolve_batch(rho0, 0.02, 40, e_ops=e_ops, H_batch=Hs, path=path)
    assert obs.shape == (40, len(pts), 2)
    # same operator objects again: the cached plan is reused and gives the same numbers
    plan_id = id(s._batch_plan[1])
    rho_f2, obs2, _ = s.evolve_batch(rho0, 0.02, 40, e_ops=e_ops, H_batch=Hs, path=path)
    assert id(s._batch_plan[1]) == plan_id and np.array_equal(obs2, obs) and np.array_equal(rho_f2, rho_f)
    for b, (o, rl) in enumerate(refs):
        assert relerr(obs[:, b], o) <= TOL
        assert relerr(rho_f[b], rl[-1]) <= TOL


def test_evolve_batch_plan_cache_follows_operator_content(cuda):
    """the plan cached on the solver is keyed on operator CONTENT: repeated calls reuse it, an in-place edit of the scan
    values or of a collapse operator rebuilds it (no stale results)"""
    from lime_b200 import builders, oqs
    pat, vals, c_ops, e_ops, rho0 = builders.jaynes_cummings_batch(1.0, 1.0 + np.linspace(-0.1, 0.1, 6), np.linspace(0.02, 0.1, 6), 8, 0.05)
    s = oqs.Lindblad_solver(None, c_ops=c_ops)
    r1, o1, _ = s.evolve_batch(rho0, 0.01, 10, e_ops=e_ops, H_batch=(pat, vals))
    plan1 = s._batch_plan[1]
    r1b, o1b, _ = s.evolve_batch(rho0, 0.01, 10, e_ops=e_ops, H_batch=(pat, vals))
    assert s._batch_plan[1] is plan1 and np.array_equal(r1, r1b)
    vals[:, :] *= 1.1                                     # in-place edit: same objects, new physics
    r2, o2, _ = s.evolve_batch(rho0, 0.01, 10, e_ops=e_ops, H_batch=(pat, vals))
    assert s._batch_plan[1] is not plan1 and relerr(r2, r1) > 1e-6
    Hs = [csr_matrix((vals[b], pat.indices, pat.indptr), shape=pat.shape) for b in range(6)]
    for b in (0, 5):
        o, rl = lo.lindblad(Hs[b].toarray(), rho0, [c.toarray() for c in c_ops], [e.toarray() for e in e_ops], Nt=10, dt=0.01)
        assert relerr(r2[b], rl[-1]) <= TOL and relerr(o2[:, b], o) <= TOL


def test_lindblad_non_hermitian_rho_and_ragged_sizes(cuda):
    """correlation functions propagate non-Hermitian 'density matrices'; odd N; N=1"""
    from lime_b200 import oqs
    for n, M in [(1, 1), (3, 2), (5, 0), (7, 3), (11, 1), (33, 1)]:
        H, c_ops, e_ops, _ = cases.lindblad_dense(n=n, M=M, E=1, seed=100 + n)
        rho0 = cases.rand_cplx(n, 7)
        o, rl = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=12, dt=0.01)
        for path in ([1, 2] if n < 33 else [1, 2]):
            plan = oqs._lindblad_plan(H, c_ops, e_ops, path=path)
            rf, ob, _ = plan.run(rho0, 0.01, 12)
            assert relerr(ob, o) <= TOL, (n, path)
            assert relerr(rf, rl[-1]) <= TOL, (n, path)


def test_lindblad_non_hermitian_hamiltonian(cuda):
    """lime evaluates the plain commutator -i[H,rho] even when H is not Hermitian
    (lime/oqs.py:706-713): the right generator is then not G^dag"""
    from lime_b200 import oqs
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=5, M=1, E=1, seed=9)
    H = H + 0.1j * cases.rand_herm(5, 10)
    o, rl = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=30, dt=0.01)
    for path in (1, 2):
        plan = oqs._lindblad_plan(H, c_ops, e_ops, path=path)
        rf, ob, _ = plan.run(rho0, 0.01, 30)
        assert relerr(ob, o) <= TOL and relerr(rf, rl[-1]) <= TOL
    res = oqs._lindblad(csr_matrix(H), rho0, [csr_matrix(c) for c in c_ops], e_ops=e_ops, Nt=30, dt=0.01)
    assert relerr(res.observables, o) <= TOL


def test_lindblad_band_chain_variant(cuda, monkeypatch):
    """opt-in sliding-window ("chain") variant of the register-tiled kernel: interleaved path ordering"""
    from lime_b200 import oqs
    monkeypatch.setenv('LIMEB200_BAND_CHAIN', '1')
    for ncav in (32, 64):
        H, c_ops, e_ops, rho0 = cases.jc_point(ncav=ncav)
        rho0 = rho0 + 0.01 * cases.rand_cplx(2 * ncav, 5)
        o, rl = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=25, dt=0.01)
        plan = oqs._lindblad_plan(csr_matrix(H), [csr_matrix(c) for c in c_ops], e_ops, path=5)
        rf, ob, _ = plan.run(rho0, 0.01, 25)
        assert relerr(ob, o) <= TOL and relerr(rf, rl[-1]) <= TOL


def test_lindblad_band_variants_and_sparse_rhs(cuda):
    """register-tiled cluster kernel: RWA Hamiltonian (other sparsity), complex couplings (generic
    complex off-diagonal G), two collapse operators, no collapse operator, non-Hermitian rho;
    liouvillian() on sparse operands goes through the permuted-basis right-hand side"""
    from lime_b200 import oqs
    for ncav, rwa, cplx_g, M in [(20, True, False, 1), (20, False, True, 2), (33, False, False, 0), (64, True, True, 2)]:
        H, c_ops, e_ops, rho0 = cases.jc_point(ncav=ncav, rwa=rwa)
        N = 2 * ncav
        if cplx_g:
            ph = np.exp(1j * np.linspace(0, 1, N))
            H = ph[:, None] * H * ph.conj()[None, :]
            H = (H + H.conj().T) / 2
        a = c_ops[0]
        cs = [a, 0.3 * a.conj().T * (1 + (0.5j if cplx_g else 0))][:M]
        rho0 = rho0 + 0.01 * cases.rand_cplx(N, 3)
        o, rl = lo.lindblad(H, rho0, cs, e_ops=e_ops, Nt=25, dt=0.01)
        plan = oqs._lindblad_plan(csr_matrix(H), [csr_matrix(c) for c in cs], e_ops, path=5)
        assert plan.path == 5
        rf, ob, _ = plan.run(rho0, 0.01, 25)
        assert relerr(ob, o) <= TOL, (ncav, rwa, cplx_g, M)
        assert relerr(rf, rl[-1]) <= TOL, (ncav, rwa, cplx_g, M)
        assert relerr(plan.rhs(rho0), lo.liouvillian(rho0, H, cs)) <= 1e-13


def test_lindblad_driven(cuda):
    from lime_b200 import oqs
    g = golden('lindblad_driven')
    H0, c_ops, e_ops, rho0 = cases.lindblad_dense(n=4, M=1, E=1, seed=77)

    def f1(t):
        return 0.3 * np.exp(-(t - 0.4) ** 2 / 0.02) * np.exp(-1j * 2.0 * t)
    res = oqs._lindblad_driven([H0.copy(), [g['H1'], f1]], rho0, c_ops=c_ops, e_ops=e_ops, Nt=60, dt=0.01,
                               t0=0.1, strict_parity=True)
    assert relerr(res.observables, g['observables_strict']) <= TOL
    assert relerr(res.rholist[-1], g['rho_final_strict']) <= TOL
    o, rl = lo.lindblad_driven([H0.copy(), [g['H1'], f1]], rho0, c_ops, e_ops, Nt=60, dt=0.01, t0=0.1)
    res = oqs.Lindblad_solver([H0.copy(), [g['H1'], f1]], c_ops=c_ops).evolve(rho0, 0.01, 60, t0=0.1, e_ops=e_ops)
    assert relerr(res.observables, o) <= TOL and relerr(res.rholist[-1], rl[-1]) <= TOL


def test_lindblad_driven_csr_operands(cuda):
    """_lindblad_driven with CSR operands and a real envelope: the sparse kernel with one set of generator values per
    step (limeb200_qme_set_step_values); a complex envelope falls to the dense kernel with its own right generator"""
    from scipy.sparse import csr_matrix
    from lime_b200 import oqs
    g = golden('lindblad_driven_csr')
    Hj, cj, ej, rj = cases.jc_point(ncav=6)
    fdr = lambda t: 0.2 * np.exp(-(t - 0.2) ** 2 / 0.02) * np.cos(3 * t)
    Hl = [csr_matrix(Hj), [csr_matrix(g['H1']), fdr]]
    res = oqs._lindblad_driven(Hl, csr_matrix(rj), c_ops=[csr_matrix(c) for c in cj], e_ops=[csr_matrix(e) for e in ej],
                               Nt=40, dt=0.01, t0=0.05)
    assert relerr(res.observables, g['obs']) <= TOL
    assert relerr(res.rholist[-1], g['rho_final']) <= TOL and relerr(res.rholist[19], g['rho_mid']) <= TOL
    # same problem, larger cavity (N = 64: the cluster / tile regime for static operators), sparse vs oracle
    Hb, cb, eb, rb = cases.jc_point(ncav=32)
    H1b = np.kron(np.array([[0, 1.], [1, 0]]), np.identity(32))
    o, rl = lo.lindblad_driven([Hb.copy(), [H1b, fdr]], rb, cb, eb, Nt=25, dt=0.01, t0=0.0)
    res = oqs._lindblad_driven([csr_matrix(Hb), [csr_matrix(H1b), fdr]], rb, c_ops=[csr_matrix(c) for c in cb], e_ops=eb,
                               Nt=25, dt=0.01)
    assert relerr(res.observables, o) <= TOL and relerr(res.rholist[-1], rl[-1]) <= TOL
    # complex envelope with CSR operands: dense kernel
    fc = lambda t: 0.2 * np.exp(-(t - 0.2) ** 2 / 0.02) * np.exp(-3j * t)
    o, rl = lo.lindblad_driven([Hj.copy(), [g['H1'], fc]], rj, cj, ej, Nt=20, dt=0.01, t0=0.0)
    res = oqs._lindblad_driven([csr_matrix(Hj), [csr_matrix(g['H1']), fc]], rj, c_ops=[csr_matrix(c) for c in cj], e_ops=ej,
                               Nt=20, dt=0.01)
    assert relerr(res.observables, o) <= TOL and relerr(res.rholist[-1], rl[-1]) <= TOL


def test_lindblad_correlations(cuda):
    from lime_b200.oqs import Lindblad_solver
    g = golden('lindblad_corr')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=4, M=2, E=1, seed=31)
    ops3 = [g['A'], g['B'], g['C']]
    s = Lindblad_solver(H, c_ops=c_ops)
    assert relerr(s.correlation_3op_1t(rho0, ops3, dt=0.02, Nt=30), g['c3op1t']) <= TOL
    c2 = s.correlation_3op_2t(rho0, ops3, 0.02, 6, 7)
    assert c2.shape == (6, 7) and relerr(c2, g['c3op2t']) <= TOL
    with pytest.raises(ValueError):
        s.correlation_4op_1t(rho0, ops3, 0.02, 4)
    ops4 = [g['A'], g['B'], g['C'], g['A']]
    ref = lo.lindblad_correlation_3op_2t(H, c_ops, rho0, [ops4[0], ops4[1] @ ops4[2], ops4[3]], 0.02, 4, 5)
    assert relerr(s.correlation_4op_2t(rho0, ops4, 0.02, 4, 5), ref) <= TOL


def test_lindblad_long_run_stability(cuda):
    """2e4 steps, N=2: accumulated difference to the oracle stays far below 1e-10"""
    from lime_b200 import oqs
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=2, M=1, E=1, seed=3)
    o, rl = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=20000, dt=0.005)
    plan = oqs._lindblad_plan(H, c_ops, e_ops)
    rf, ob, _ = plan.run(rho0, 0.005, 20000)
    assert relerr(ob, o) <= TOL and relerr(rf, rl[-1]) <= TOL
    assert abs(np.trace(rf) - 1) < 1e-11


# ---------------------------------------------------------------- Redfield
def test_redfield_example_config1(cuda):
    from lime_b200.oqs import Redfield_solver
    g = golden('redfield_example')
    H, a_ops, spectra, rho0, dt, Nt, e_ops, tlist = cases.redfield_example()
    s = Redfield_solver(H, c_ops=a_ops, spectra=spectra)
    R, evecs = s.redfield_tensor()
    res = s.evolve(rho0, evecs=evecs, dt=dt, Nt=Nt, e_ops=e_ops)
    assert res.observables.shape == (200, 1)
    assert relerr(res.observables, g['observables']) <= TOL
    assert relerr(np.array(res.rholist), g['rholist']) <= TOL
    t8 = tlist[:8]
    assert relerr(s.propagator(t8, 'EOM'), g['U_eom']) <= TOL
    U = s.propagator(t8, 'SOS')
    assert relerr(U, g['U_sos']) <= 1e-10
    assert relerr(s.expect(rho0, e_ops), g['expect']) <= 1e-10
    assert relerr(s.correlation_4op_3t(rho0, [e_ops[0]] * 4, 'llll', t8), g['corr4']) <= 1e-10


def test_redfield_multilevel_tensor_and_operator_forms(cuda):
    from lime_b200.oqs import Redfield_solver, _redfield
    g = golden('redfield_multilevel')
    H, a_ops, spectra, rho0 = cases.redfield_multilevel()
    s = Redfield_solver(H, c_ops=a_ops, spectra=spectra)
    R, evecs = s.redfield_tensor()
    e_ops = [a_ops[0], g['e1']]
    res = s.evolve(rho0, dt=0.02, Nt=60, e_ops=e_ops)
    assert relerr(res.observables, g['observables']) <= TOL
    assert relerr(np.array(res.rholist), g['rholist']) <= TOL
    res = _redfield(R, rho0, evecs=evecs, Nt=60, dt=0.02, e_ops=e_ops)
    assert relerr(res.observables, g['observables']) <= TOL
    # [ext] batched, both forms; final state returned in the eigenbasis
    batch = cases.rand_dm_batch(9, 5, 4)
    for form in ('tensor', 'operator'):
        out, obs = s.evolve_batch(batch, 0.02, 30, e_ops=e_ops, form=form)
        for b in range(9):
            o, rl = lo.redfield(R, batch[b], evecs=evecs, Nt=30, dt=0.02, e_ops=e_ops)
            assert relerr(obs[:, b], o) <= TOL, form
            assert relerr(evecs @ out[b] @ evecs.conj().T, rl[-1]) <= TOL, form


# ---------------------------------------------------------------- HEOM
def test_heom_dl_exact(cuda):
    import io, contextlib, os, tempfile
    from lime_b200.oqs import _heom_dl
    g = golden('heom_dl')
    H, sz, rho0 = cases.spin_boson_heom()[:3]
    fn = os.path.join(tempfile.mkdtemp(), 'heom.dat')
    with contextlib.redirect_stdout(io.StringIO()):
        out = _heom_dl(H, rho0, sz, None, 300.0, 0.002, 0.0005, 12, 0.05, 200, fn)
    assert relerr(out, g['rho_final']) <= TOL
    traj = np.genfromtxt(fn, dtype=complex)[:, 1:].reshape(-1, 2, 2)
    assert relerr(traj, g['traj']) <= TOL


@pytest.mark.parametrize('path', [1, 2, 3, 4])
def test_heom_spin_boson_multi_index(cuda, path):
    """config-3 shape: K=2 Matsubara terms, depth 12 -> 91 ADOs, 312 couplings (diagonal Q)"""
    from lime_b200.heom.heom import HEOM
    H, sz, rho0, depth, K, lam, gam, T = cases.spin_boson_heom(depth=12)
    h = HEOM(H, sz, lam, gam, T, N_exp=K, N_cut=depth)
    assert h.nhe == 91
    h.plan.set_path(path)
    sx = np.array([[0, 1.], [1, 0]])
    res = h.evolve(rho0, 0.01, 50, e_ops=[sz, sx])
    ado_o, obs_o, traj_o = lo.heom_rk4(h.initial(rho0), H, h.Q, h.qmap, h.c, h.nu, h.states.astype(np.int64),
                                       h.dn.astype(np.int64), h.up.astype(np.int64), 0.01, 50,
                                       e_ops=[sz, sx], store=True)
    assert h.plan.path == path
    assert relerr(res.ado, ado_o) <= TOL
    assert relerr(res.observables, obs_o) <= TOL
    assert relerr(np.array(res.rholist), np.array(traj_o)) <= TOL
    assert relerr(h.plan.rhs(ado_o), lo.heom_rhs(ado_o, H, h.Q, h.qmap, h.c, h.nu, h.states.astype(np.int64),
                                                 h.dn.astype(np.int64), h.up.astype(np.int64))) <= 1e-12


@pytest.mark.parametrize('path', [1, 2, 3])
def test_heom_dense_q_and_multibath(cuda, path):
    """non-diagonal coupling operators, 2 baths with their own Q, n=3"""
    from lime_b200.heom.heom import HEOM
    n = 3
    H = cases.rand_herm(n, 1)
    Q = [cases.rand_herm(n, 2), cases.rand_herm(n, 3)]
    rho0 = cases.rand_dm(n, 4)
    h = HEOM(H, Q, [0.1, 0.2], [0.8, 1.3], 1.5, N_exp=2, N_cut=3)
    h.plan.set_path(path)
    res = h.evolve(rho0, 0.01, 25, e_ops=[H])
    st = h.states.astype(np.int64)
    ado_o, obs_o, _ = lo.heom_rk4(h.initial(rho0), H, h.Q, h.qmap, h.c, h.nu, st, h.dn.astype(np.int64),
                                  h.up.astype(np.int64), 0.01, 25, e_ops=[H])
    assert relerr(res.ado, ado_o) <= TOL and relerr(res.observables, obs_o) <= TOL


@pytest.mark.parametrize('name', ['k2_d3_n2', 'k4_d2_n3', 'b2k2_d3_n3', 'b2k2_d2_proj'])
def test_heom_rhs_against_reference_rule_matrix(cuda, name):
    """device RHS of every HEOM kernel family against the REFERENCE's own hierarchy Liouvillian
    (lime/heom/heom.py:156-216 exec'd in oracle/gen_golden.py, frozen in heom_rules.npz) plus -i[H, rho_n]"""
    from lime_b200 import engine
    from test_oracle_vs_golden import heom_rule_case
    k = heom_rule_case(name)
    n, nhe = k['n'], k['st'].shape[0]
    H = cases.rand_herm(n, 31)
    rng = np.random.default_rng(8)
    ado = rng.standard_normal((nhe, n, n)) + 1j * rng.standard_normal((nhe, n, n))
    ref = (k['L'] @ ado.reshape(-1)).reshape(nhe, n, n) - 1j * (H @ ado - ado @ H)
    plan = engine.HeomPlan(H, k['Q'], k['qmap'], k['c'], k['nu'], k['st'], k['dn'], k['up'])
    assert relerr(plan.rhs(ado), ref) <= 1e-12
    # and one propagated trajectory per run path against RK4 of the reference matrix
    Lfull = k['L'].copy()
    nn = n * n
    for a in range(nhe):
        Lfull[a * nn:(a + 1) * nn, a * nn:(a + 1) * nn] += -1j * (np.kron(H, np.eye(n)) - np.kron(np.eye(n), H.T))
    y = ado.reshape(-1).copy() * 0.1
    y0 = y.copy()
    for _ in range(20):
        y = lo.rk4(y, lambda v: Lfull @ v, 0.01)
    for path in (1, 2, 3):
        pl = engine.HeomPlan(H, k['Q'], k['qmap'], k['c'], k['nu'], k['st'], k['dn'], k['up'])
        pl.set_path(path)
        out, _, _ = pl.run(y0.reshape(nhe, n, n), 0.01, 20)
        assert relerr(out.reshape(-1), y) <= TOL, path


@pytest.mark.parametrize('path', [2, 3, 4])
def test_heom_config4_full_size(cuda, path):
    """config 4 at its real size: FMO, 7 baths x K = 2, depth 4 -> 3060 ADOs of 7x7, 19 040 couplings; 50 RK4 steps
    of the stage-wise (2), barrier-persistent (3) and dataflow-persistent (4) kernels against the oracle"""
    from lime_b200 import builders
    from lime_b200.heom.heom import HEOM
    from lime_b200.units import au2fs
    Hm, Q, lam, gam, kT = builders.fmo_heom_inputs()
    h = HEOM(Hm, Q, lam, gam, kT, N_exp=2, N_cut=4)
    assert h.nhe == 3060 and int((h.dn >= 0).sum() + (h.up >= 0).sum()) == 19040
    h.plan.set_path(path)
    rho0 = np.zeros((7, 7), dtype=complex)
    rho0[0, 0] = 1.0
    dt = 0.5 / au2fs
    res = h.evolve(rho0, dt, 50, e_ops=[Q[0], Q[1]], store_states=False)
    assert h.plan.path == path
    st = h.states.astype(np.int64)
    ado_o, obs_o, _ = lo.heom_rk4(h.initial(rho0), Hm, h.Q, h.qmap, h.c, h.nu, st, h.dn.astype(np.int64),
                                  h.up.astype(np.int64), dt, 50, e_ops=[Q[0], Q[1]])
    assert relerr(res.ado, ado_o) <= TOL and relerr(res.observables, obs_o) <= TOL


def test_heom_solver_dl_class(cuda, tmp_path):
    """HEOMSolverDL (lime/oqs.py:1335-1431): `solve` runs the Lindblad propagator exactly as lime's does (:1364-1366),
    the regression correlation functions, and the [ext] hierarchy entry point solve_heom"""
    from lime_b200 import oqs
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=6)
    s = oqs.HEOMSolverDL(H=H, c_ops=c_ops, e_ops=e_ops)
    res = s.solve(rho0, 0.01, 30, True)
    obs_o, rl_o = lo.lindblad(H, rho0, c_ops, e_ops, Nt=30, dt=0.01)
    assert relerr(res.observables, obs_o) <= TOL and relerr(np.array(res.rholist), np.array(rl_o)) <= TOL
    s.set_e_ops(e_ops[:1]); s.set_c_ops(c_ops); s.setH(H); s.configure(c_ops, e_ops)
    g = golden('api_r2')
    cor = s.correlation_2op_1t(rho0, g['A'], g['B'], 0.01, 30, output=str(tmp_path / 'c.dat'))
    assert relerr(cor, g['cor2']) <= TOL
    ops = [e_ops[0], e_ops[1], c_ops[0]]
    assert relerr(s.correlation_3op_2t(rho0, ops, 0.01, 5, 7),
                  lo.lindblad_correlation_3op_2t(H, c_ops, rho0, ops, 0.01, 5, 7)) <= TOL
    Hs, sz, r0, depth, K, lam, gam, T = cases.spin_boson_heom(depth=5)
    hs = oqs.HEOMSolverDL(H=Hs, c_ops=[sz], e_ops=[sz])
    r = hs.solve_heom(r0, 0.01, 40, lam, gam, T, N_exp=K, N_cut=depth)
    st, dn, up = lo.heom_tables([depth + 1] * K, depth)
    c, nu = lo.calc_matsubara_params(K, lam, gam, T)
    ado0 = np.zeros((st.shape[0], 2, 2), dtype=complex)
    ado0[0] = r0
    ado_o, obs_o, tr_o = lo.heom_rk4(ado0, Hs, sz[None], [0] * K, c, nu, st, dn, up, 0.01, 40, e_ops=[sz], store=True)
    assert relerr(r.ado, ado_o) <= TOL and relerr(r.observables, obs_o) <= TOL
    assert relerr(np.array(r.rholist), np.array(tr_o)) <= TOL


def test_lindblad_correlation_2op_1t(cuda, tmp_path):
    """<A(t)B>, lime/oqs.py:726-800, 1196-1225: values and the cor.dat side effect, dense and CSR operands"""
    from scipy.sparse import csr_matrix
    from lime_b200 import oqs
    g = golden('api_r2')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=6)
    fn = str(tmp_path / 'cor.dat')
    cor = oqs.Lindblad_solver(H, c_ops).correlation_2op_1t(rho0, g['A'], g['B'], 0.01, 30, output=fn)
    assert relerr(cor, g['cor2']) <= TOL
    got = np.array([complex(l.split()[1]) for l in open(fn)])
    ref = np.array([complex(l.split()[1]) for l in str(g['cor_txt']).splitlines()])
    tt = np.array([float(l.split()[0]) for l in open(fn)])
    tr = np.array([float(l.split()[0]) for l in str(g['cor_txt']).splitlines()])
    assert np.array_equal(tt, tr) and relerr(got, ref) <= TOL          # same file layout, times bit-identical
    Hj, cj, ej, rj = cases.jc_point(ncav=8)
    corj = oqs._correlation_2p_1t(csr_matrix(Hj), rj, [ej[0], cj[0] / np.sqrt(0.05)], [csr_matrix(c) for c in cj], 0.01, 25,
                                  output=str(tmp_path / 'c2.dat'))
    assert relerr(corj, g['cor2_jc']) <= TOL
    with pytest.raises(SystemExit):
        oqs._correlation_2p_1t(H, rho0, [g['A'], g['B']], c_ops, 0.01, 3, method='redfield', output=str(tmp_path / 'c3.dat'))


def test_redfield_green_functions(cuda):
    """Redfield_solver.gf / getG (lime/oqs.py:145-167, 474-526) and correlation_2op_1t (:254-275)"""
    from lime_b200 import oqs
    g = golden('api_r2')
    Hr, a_ops, spectra, rr = cases.redfield_multilevel()
    sol = oqs.Redfield_solver(Hr, c_ops=a_ops, spectra=spectra)
    assert relerr(sol.gf(g['tg'], method='diag'), g['G_time']) <= 1e-11       # redfield_tensor is built on demand
    assert relerr(oqs.getG(1j * sol.R, g['tg'], w=g['wg'], domain='freq'), g['G_freq']) <= 1e-11
    with pytest.raises(ValueError):
        oqs.getG(1j * sol.R, g['tg'], w=g['wg'][:-1], domain='freq')           # lime's einsum needs len(w) == dim(L)
    # 'EOM': lime multiplies a Python list by -1j (always TypeError); here: the list lime meant to return
    t = np.linspace(0, 0.5, 6)
    Gl = sol.gf(t, method='EOM')
    U = lo.redfield_propagator(sol.R, t, method='EOM')
    assert len(Gl) == 6 and relerr(np.stack([x.toarray() for x in Gl], axis=-1), -1j * np.asarray(U)) <= TOL
    # two-point function through the propagator, lime/oqs.py:254-275: in lime `idm = dm2vec(identity(dim))` is a sparse
    # (dim^2, 1) column, so `idm.dot(...)` raises ValueError for every input -- reproduced as it is
    sol2 = oqs.Redfield_solver(Hr, c_ops=a_ops, spectra=spectra)
    sol2.redfield_tensor()
    from lime_b200.superoperator import operator_to_superoperator
    a = operator_to_superoperator(a_ops[0], 'left')
    b = operator_to_superoperator(a_ops[1], 'left')
    with pytest.raises(ValueError):
        sol2.correlation_2op_1t(rr, a, b, t)


def test_heom_dataflow_kernel_tiles_and_trajectory(cuda, monkeypatch):
    """path 4, tiled variant, on a hierarchy large enough that a CTA walks several tiles (FMO depth 5: 11 628 ADOs -> 79
    per CTA, 4 tiles of 20), random non-Hermitian ADOs, observables + tier-0 trajectory, against the barrier kernel"""
    from lime_b200 import builders
    import lime_b200.heom.heom as hh
    Hm, Q, lam, gam, kT = builders.fmo_heom_inputs()
    h = hh.HEOM(Hm, Q, lam, gam, kT, N_exp=2, N_cut=5)
    rng = np.random.default_rng(12)
    ado0 = 0.1 * (rng.standard_normal((h.nhe, 7, 7)) + 1j * rng.standard_normal((h.nhe, 7, 7)))
    dt = 20.0
    h.plan.set_path(3)
    o3, b3, t3 = h.plan.run(ado0, dt, 12, e_ops=[Q[0], Hm], traj_every=4)
    h.plan.set_path(4)
    o4, b4, t4 = h.plan.run(ado0, dt, 12, e_ops=[Q[0], Hm], traj_every=4)
    assert h.plan.path == 4
    assert relerr(o4, o3) <= 1e-13 and relerr(b4, b3) <= 1e-13 and relerr(t4, t3) <= 1e-13
    o4b, _, _ = h.plan.run(o4, dt, 3)              # a second launch on the same plan: tags continue
    h.plan.set_path(3)
    o3b, _, _ = h.plan.run(o3, dt, 3)
    assert relerr(o4b, o3b) <= 1e-13
    # the tiled variant forced on a hierarchy that would take the register-resident one
    monkeypatch.setenv('LIMEB200_HEOM_FLOW_TILED', '1')
    h2 = hh.HEOM(Hm, Q, lam, gam, kT, N_exp=2, N_cut=3)
    a0 = 0.1 * (rng.standard_normal((h2.nhe, 7, 7)) + 1j * rng.standard_normal((h2.nhe, 7, 7)))
    h2.plan.set_path(4)
    p4, q4, _ = h2.plan.run(a0, dt, 9, e_ops=[Q[1]])
    monkeypatch.delenv('LIMEB200_HEOM_FLOW_TILED')
    h2.plan.set_path(3)
    p3, q3, _ = h2.plan.run(a0, dt, 9, e_ops=[Q[1]])
    assert relerr(p4, p3) <= 1e-13 and relerr(q4, q3) <= 1e-13


def test_heom_dataflow_kernel_two_elements_per_thread(cuda):
    """path 4, register-resident variant with TWO matrix elements per thread: 7 sites, 11 bath modes (four baths with
    two exponentials, three with one), depth 5 -> 4368 ADOs = 30 per SM; against the barrier kernel (path 3) and the
    oracle"""
    from lime_b200 import builders, engine
    from lime_b200.heom.heom import _calc_matsubara_params
    Hm, Q, lam, gam, kT = builders.fmo_heom_inputs()
    qmap = [0, 0, 1, 1, 2, 2, 3, 3, 4, 5, 6]
    c, nu = [], []
    for b in range(7):
        cb, nub = _calc_matsubara_params(2, lam * (1 + 0.1 * b), gam, kT)
        k = 2 if b < 4 else 1
        c += cb[:k]
        nu += nub[:k]
    states, dn, up = engine.heom_tables([6] * 11, 5)
    assert states.shape[0] == 4368
    plan = engine.HeomPlan(Hm, np.stack(Q).astype(complex), qmap, np.array(c), np.array(nu), states, dn, up)
    rng = np.random.default_rng(3)
    ado0 = 0.1 * (rng.standard_normal((4368, 7, 7)) + 1j * rng.standard_normal((4368, 7, 7)))
    plan.set_path(3)
    o3, b3, _ = plan.run(ado0, 15.0, 10, e_ops=[Q[2]])
    plan.set_path(4)
    o4, b4, _ = plan.run(ado0, 15.0, 10, e_ops=[Q[2]])
    assert plan.path == 4
    assert relerr(o4, o3) <= 1e-13 and relerr(b4, b3) <= 1e-13
    st = states.astype(np.int64)
    ado_o, _, _ = lo.heom_rk4(ado0, Hm, np.stack(Q).astype(complex), qmap, np.array(c), np.array(nu), st,
                              dn.astype(np.int64), up.astype(np.int64), 15.0, 2)
    o2, _, _ = plan.run(ado0, 15.0, 2)
    assert relerr(o2, ado_o) <= TOL


@pytest.mark.parametrize('path', [0, 2, 3])
def test_heom_fmo_shape_stagewise_and_batch(cuda, path):
    """config-4 shape at reduced depth: 7 sites, 7 baths x K=2, depth 2 (120 ADOs of 7x7);
    projector coupling operators (diagonal-Q gather path); batch of 3 hierarchies"""
    from lime_b200 import engine
    import lime_b200.heom.heom as hh
    n = 7
    H = cases.rand_herm(n, 21, 0.5) + np.diag(np.arange(n) * 0.3)
    Q = [np.diag((np.arange(n) == j).astype(float)) for j in range(n)]
    h = hh.HEOM(H, Q, 0.05, 0.6, 1.2, N_exp=2, N_cut=2)
    assert h.nhe == 120
    st = h.states.astype(np.int64)
    rng = np.random.default_rng(5)
    ado0 = rng.standard_normal((3, h.nhe, n, n)) + 1j * rng.standard_normal((3, h.nhe, n, n))
    ado0 *= 0.1
    h.plan.set_path(path)
    out, obs, traj = h.plan.run(ado0, 0.02, 10, e_ops=[H], traj_every=5)
    assert relerr(traj[-1], out[:, 0]) == 0
    for b in range(3):
        ado_o, obs_o, _ = lo.heom_rk4(ado0[b], H, h.Q, h.qmap, h.c, h.nu, st, h.dn.astype(np.int64),
                                      h.up.astype(np.int64), 0.02, 10, e_ops=[H])
        assert relerr(out[b], ado_o) <= TOL and relerr(obs[:, b], obs_o) <= TOL


def test_heom_parameter_batch(cuda):
    """[ext] config-3 throughput variant: hierarchies differing in (lambda, beta)"""
    from lime_b200 import engine
    from lime_b200.heom.heom import _calc_matsubara_params
    H, sz, rho0, depth, K, lam, gam, T = cases.spin_boson_heom(depth=5)
    states, dn, up = engine.heom_tables([depth + 1] * K, depth)
    pars = [(0.1, 1.0), (0.2, 0.7), (0.3, 1.4), (0.05, 2.0)]
    cs, nus = [], []
    for lam_b, T_b in pars:
        c, nu = _calc_matsubara_params(K, lam_b, gam, T_b)
        cs.append(c)
        nus.append(nu)
    plan = engine.HeomPlan(H, sz, [0] * K, np.array(cs), np.array(nus), states, dn, up)
    ado0 = np.zeros((len(pars), states.shape[0], 2, 2), dtype=complex)
    ado0[:, 0] = rho0
    out, obs, _ = plan.run(ado0, 0.01, 40, e_ops=[sz])
    for b in range(len(pars)):
        ado_o, obs_o, _ = lo.heom_rk4(ado0[b], H, sz[None], [0] * K, cs[b], nus[b], states.astype(np.int64),
                                      dn.astype(np.int64), up.astype(np.int64), 0.01, 40, e_ops=[sz])
        assert relerr(out[b], ado_o) <= TOL and relerr(obs[:, b], obs_o) <= TOL


def test_heom_stagewise_batch_l2_slices(cuda, monkeypatch):
    """stage-wise batch path: (i) the packed-neighbour forms (LIMEB200_HEOM_STAGE_MODE 1: generic tile code with the
    packed gather, 2: heom_stage_fast_kernel with the coefficient table, 3: with n_k x base coefficients from shared
    memory) agree with the table-walking kernel (mode 0) to rounding, also in the generic persistent kernel; (ii) walked
    in slices of the batch (opt-in; forced slice of 2 on a batch of 5: ragged last slice) the results are identical to the
    batch-wide launches -- shared bath parameters (FMO shape, observables + trajectory) and per-hierarchy parameters
    (the slice offsets the parameter and coefficient tables)"""
    from lime_b200 import engine
    import lime_b200.heom.heom as hh
    from lime_b200.heom.heom import _calc_matsubara_params
    n = 7
    H = cases.rand_herm(n, 21, 0.5) + np.diag(np.arange(n) * 0.3)
    Q = [np.diag((np.arange(n) == j).astype(float)) for j in range(n)]
    rng = np.random.default_rng(6)
    for Hsys in (H.real.astype(complex), H):          # real Hamiltonian (the real -i[H, .] form) and a complex one
        h = hh.HEOM(Hsys, Q, 0.05, 0.6, 1.2, N_exp=2, N_cut=2)
        ado0 = 0.1 * (rng.standard_normal((5, h.nhe, n, n)) + 1j * rng.standard_normal((5, h.nhe, n, n)))
        h.plan.set_path(2)
        res = {}
        for mode in ('0', '1', '2', '3'):
            monkeypatch.setenv('LIMEB200_HEOM_STAGE_MODE', mode)
            res[mode] = h.plan.run(ado0, 0.02, 10, e_ops=[Hsys], traj_every=5)
            monkeypatch.setenv('LIMEB200_HEOM_L2_CHUNK', '2')
            sl = h.plan.run(ado0, 0.02, 10, e_ops=[Hsys], traj_every=5)
            monkeypatch.delenv('LIMEB200_HEOM_L2_CHUNK')
            for a, b in zip(res[mode], sl):
                assert np.array_equal(a, b)
            for a, b in zip(res['0'], res[mode]):     # same operations in the same order (coefficients tabulated)
                assert relerr(a, b) <= 1e-14
        ado_o, obs_o, _ = lo.heom_rk4(ado0[4], Hsys, h.Q, h.qmap, h.c, h.nu, h.states.astype(np.int64), h.dn.astype(np.int64),
                                      h.up.astype(np.int64), 0.02, 10, e_ops=[Hsys])
        for mode in ('0', '3'):
            assert relerr(res[mode][0][4], ado_o) <= TOL and relerr(res[mode][1][:, 4], obs_o) <= TOL
        # the generic persistent kernel (the one the ADO-sharded barrier path runs on large shards;
        # LIMEB200_HEOM_NO_CACHED keeps the cached kernel out) with and without the packed gather
        monkeypatch.setenv('LIMEB200_HEOM_NO_CACHED', '1')
        h.plan.set_path(3)
        for m in ('0', '1'):
            monkeypatch.setenv('LIMEB200_HEOM_STAGE_MODE', m)
            out3 = h.plan.run(ado0, 0.02, 10, e_ops=[Hsys], traj_every=5)
            assert h.plan.path == 3
            for a, b in zip(res['0'], out3):
                assert relerr(a, b) <= 1e-14
        monkeypatch.delenv('LIMEB200_HEOM_NO_CACHED')
    monkeypatch.setenv('LIMEB200_HEOM_STAGE_MODE', '2')
    monkeypatch.setenv('LIMEB200_HEOM_L2_CHUNK', '2')
    # per-hierarchy bath parameters
    Hs, sz, rho0, depth, K, lam, gam, T = cases.spin_boson_heom(depth=5)
    states, dn, up = engine.heom_tables([depth + 1] * K, depth)
    pars = [(0.1, 1.0), (0.2, 0.7), (0.3, 1.4), (0.05, 2.0), (0.15, 0.9)]
    cs, nus = zip(*[_calc_matsubara_params(K, l, gam, t) for l, t in pars])
    plan = engine.HeomPlan(Hs, sz, [0] * K, np.array(cs), np.array(nus), states, dn, up)
    plan.set_path(2)
    a0 = np.zeros((len(pars), states.shape[0], 2, 2), dtype=complex)
    a0[:, 0] = rho0
    out, obs, _ = plan.run(a0, 0.01, 40, e_ops=[sz])
    assert plan.path == 2
    for b in (0, 3, 4):
        ado_o, obs_o, _ = lo.heom_rk4(a0[b], Hs, sz[None], [0] * K, cs[b], nus[b], states.astype(np.int64),
                                      dn.astype(np.int64), up.astype(np.int64), 0.01, 40, e_ops=[sz])
        assert relerr(out[b], ado_o) <= TOL and relerr(obs[:, b], obs_o) <= TOL


# ---------------------------------------------------------------- SOS
def test_sos_all_pathways(cuda):
    from lime_b200.signal import sos
    g = golden('sos')
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    w1, w3, w2, w1b, t2 = g['w1'], g['w3'], g['w2'], g['w1b'], float(g['t2'])
    au2ev = 27.211386
    assert relerr(sos.GSB(E, dip, w1, w3, t2, g_idx, e_idx, gamma), g['GSB']) <= TOL
    assert relerr(sos.SE(E, dip, w1, w3, t2, g_idx, e_idx, gamma), g['SE']) <= TOL
    assert relerr(sos.ESA(E, dip, w1, w3, t2, g_idx, e_idx, f_idx, gamma), g['ESA']) <= TOL
    pe = sos._photon_echo(E, dip, -w1, w3, t2, g_idx, e_idx, f_idx, gamma)
    assert pe.shape == (12, 12) and pe.dtype == np.complex128
    assert relerr(pe, g['PE']) <= TOL
    assert relerr(sos._SE(E, dip, -w1, w3, t2, g_idx, e_idx, gamma, dephasing=0.01 / au2ev), g['SE_t3']) <= TOL
    assert relerr(sos._ESA(E, dip, -w1, w3, t2, g_idx, e_idx, f_idx, gamma, dephasing=0.01 / au2ev), g['ESA_t3']) <= TOL
    kw = dict(g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    assert relerr(sos.DQC_R1(E, dip, omega1=w1b, omega2=w2, tau3=1e-6, **kw), g['R1_t3']) <= TOL
    assert relerr(sos.DQC_R2(E, dip, omega1=w1b, omega2=w2, tau3=1e-6, **kw), g['R2_t3']) <= TOL
    assert relerr(sos.DQC_R1(E, dip, omega2=w2, omega3=w1b, tau1=50.0, **kw), g['R1_t1']) <= TOL
    assert relerr(sos.DQC_R2(E, dip, omega2=w2, omega3=w1b, tau1=50.0, **kw), g['R2_t1']) <= TOL
    assert relerr(sos.TPA2D(E, dip, w2, w1b, g_idx, e_idx, f_idx, gamma), g['TPA2D']) <= TOL
    assert relerr(sos.TPA2D_time_order(E, dip, w2, w1b, g_idx, e_idx, f_idx, gamma), g['TPA2D_to']) <= TOL
    with pytest.raises(Exception):
        sos.DQC_R2(E, dip, omega2=w2, **kw)


def test_sos_waiting_time_batch_and_ragged_grid(cuda):
    """[ext] vector of waiting times in one launch; non-multiple-of-tile grid; N=32 config-5 shape"""
    from lime_b200.signal import sos
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system(N=32, ne=15, seed=0)
    au2ev, au2fs = 27.211386, 2.41888432651e-2
    w = np.linspace(1.4, 2.1, 45) / au2ev
    T = np.linspace(0, 630, 5) / au2fs
    pe = sos._photon_echo(E, dip, -w, w, T, g_idx, e_idx, f_idx, gamma)
    assert pe.shape == (5, 45, 45)
    for t in (0, 2, 4):
        assert relerr(pe[t], lo.photon_echo_core(E, dip, -w, w, T[t], g_idx, e_idx, f_idx, gamma)) <= TOL
    # linearity in the dipole scale: S ~ mu^4
    pe2 = sos._photon_echo(E, 2.0 * dip, -w, w, T[1], g_idx, e_idx, f_idx, gamma)
    assert relerr(pe2, 16.0 * pe[1]) <= 1e-12
    # [ext] device-resident evaluation (what bench.py times) gives the same numbers, repeatedly
    grid = sos.PhotonEchoGrid(E, dip, -w, w, T, g_idx, e_idx, f_idx, gamma)
    assert relerr(grid.run().cpu().numpy(), pe) <= 1e-14
    assert relerr(grid.run().cpu().numpy(), pe) <= 1e-14


def test_time_domain_2des(cuda):
    """lime/signal/2DES.py:37-247 (G, ESA, GSB, SE in the time domain) against the frozen outputs of the
    reference functions themselves and against the oracle; scalars, vectors, waiting-time batch"""
    from lime_b200.signal import twodes
    g = golden('twodes_time')
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    t1, t3, tw = g['t1'], g['t3'], float(g['t2'])
    twodes.en, twodes.decay = None, None
    assert relerr(twodes.ESA(E, dip, g_idx, e_idx, f_idx, gamma, t1[None, :], tw, t3[:, None]), g['ESA']) <= TOL
    assert relerr(twodes.GSB(E, dip, g_idx, e_idx, gamma, t1, tw, t3), g['GSB']) <= TOL
    assert relerr(twodes.SE(E, dip, g_idx, e_idx, t1, tw, t3, gamma=gamma), g['SE']) <= TOL
    # module globals as in lime
    twodes.en, twodes.decay = E, gamma
    try:
        assert relerr(twodes.SE(None, dip, g_idx, e_idx, t1, tw, t3), g['SE']) <= TOL
        assert relerr(twodes.G(2, 0, t3), lo.td_G(E, gamma, 2, 0, t3)) <= 1e-13
        assert abs(twodes.G(1, 0, 3.0) - lo.td_G(E, gamma, 1, 0, 3.0)) <= 1e-13
        assert twodes.G(1, 0, -1.0) == 0
    finally:
        twodes.en, twodes.decay = None, None
    # scalar delays and a batch of waiting times
    s = twodes.ESA(E, dip, g_idx, e_idx, f_idx, gamma, 10.0, tw, 20.0)
    assert abs(s - lo.td_ESA(E, gamma, dip, g_idx, e_idx, f_idx, 10.0, tw, 20.0)) <= TOL * abs(s)
    tws = np.array([0.0, 25.0, 80.0])
    b = twodes.GSB(E, dip, g_idx, e_idx, gamma, t1, tws, t3)
    assert b.shape == (3, len(t3), len(t1))
    for k, t2 in enumerate(tws):
        assert relerr(b[k], lo.td_GSB(E, gamma, dip, g_idx, e_idx, t1[None, :], t2, t3[:, None])) <= TOL


def test_redfield_two_level_batch_thread_per_vector_kernel(cuda):
    """config-1 throughput variant: the Redfield tensor of examples/redfield.py on a batch large enough to
    take the thread-per-vector kernel; checked against the oracle on a few members"""
    from lime_b200.oqs import Redfield_solver
    H, a_ops, spectra, rho0, dt, Nt, e_ops, tlist = cases.redfield_example()
    s = Redfield_solver(H, c_ops=a_ops, spectra=spectra)
    R, evecs = s.redfield_tensor()
    batch = cases.rand_dm_batch(5000, 2, 9)
    out, obs = s.evolve_batch(batch, dt, 40, e_ops=e_ops)
    assert obs.shape == (40, 5000, 1)
    for b in (0, 17, 4999):
        o, rl = lo.redfield(R, batch[b], evecs=evecs, Nt=40, dt=dt, e_ops=e_ops)
        assert relerr(obs[:, b], o) <= TOL
        assert relerr(evecs @ out[b] @ evecs.conj().T, rl[-1]) <= TOL


@pytest.mark.parametrize('no_dmma', [False, True])
def test_lindblad_dense_stage_64_tile_ragged(cuda, monkeypatch, no_dmma):
    """dense stage-wise path at a size that is no multiple of the 64x64 tile: the FP64 tensor-core kernel
    (qme_dense_stage_dmma, N >= 96) and, with LIMEB200_DENSE_NO_DMMA, the register-tiled DFMA kernel it replaced"""
    from lime_b200 import oqs
    if no_dmma:
        monkeypatch.setenv('LIMEB200_DENSE_NO_DMMA', '1')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=101, M=2, E=1, seed=77)
    H = H / 10.0
    o, rl = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=6, dt=0.01)
    plan = oqs._lindblad_plan(H, c_ops, e_ops, path=2)
    rf, ob, _ = plan.run(rho0, 0.01, 6)
    assert relerr(ob, o) <= TOL and relerr(rf, rl[-1]) <= TOL
    assert relerr(plan.rhs(rho0), lo.liouvillian(rho0, H, c_ops)) <= 1e-12


@pytest.mark.parametrize('no_dmma', [False, True])
def test_lindblad_dense_config2prime_full_size(cuda, monkeypatch, no_dmma):
    """config 2' at its named size: N = 256, M = 2 dense operators, a batch of 3, 4 RK4 steps (DMMA and DFMA kernels);
    and the 32x32-tile kernel below N = 96 with the DMMA path disabled"""
    from lime_b200 import oqs
    if no_dmma:
        monkeypatch.setenv('LIMEB200_DENSE_NO_DMMA', '1')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=256, M=2, E=2, seed=123)
    H = H / 20.0
    c_ops = [c / 4.0 for c in c_ops]
    plan = oqs._lindblad_plan(H, c_ops, e_ops, path=2)
    r0 = np.stack([rho0, cases.rand_dm(256, 5), cases.rand_dm(256, 6)])
    rf, ob, _ = plan.run(r0, 0.01, 4)
    for b in (0, 2):
        o, rl = lo.lindblad(H, r0[b], c_ops, e_ops=e_ops, Nt=4, dt=0.01)
        assert relerr(ob[:, b], o) <= TOL and relerr(rf[b], rl[-1]) <= TOL
    if no_dmma:
        H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=80, M=1, E=1, seed=9)
        o, rl = lo.lindblad(H / 8.0, rho0, c_ops, e_ops=e_ops, Nt=5, dt=0.01)
        plan = oqs._lindblad_plan(H / 8.0, c_ops, e_ops, path=2)
        rf, ob, _ = plan.run(rho0, 0.01, 5)
        assert relerr(ob, o) <= TOL and relerr(rf, rl[-1]) <= TOL


def test_lindblad_config2_full_batch(cuda):
    """config 2 at its real batch: 4096 coupling x detuning points of N = 128 in ONE launch (the launch bench.py
    times), 12 RK4 steps; three of the points against the oracle, Tr rho = 1 on all of them"""
    import torch
    from lime_b200 import builders, oqs
    pat, vals, c_ops, e_ops, rho0 = builders.jc_grid()
    assert vals.shape[0] == 4096
    plan, B = oqs._lindblad_plan_batch((pat, vals), c_ops, e_ops)
    assert plan.path == 6
    rho = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(rho0, (4096, 128, 128)))).cuda()
    obs, _ = plan.run_device(rho, 0.01, 12)
    tr = torch.einsum('bii->b', rho).cpu().numpy()
    assert np.max(np.abs(tr - 1)) < 1e-12
    gs = np.linspace(0.01, 0.2, 64)
    dets = np.linspace(-0.2, 0.2, 64)
    for b in (0, 2077, 4095):
        Ho, co, eo = lo.jaynes_cummings(1.0, 1.0 + dets[b % 64], gs[b // 64], 64, 0.05)
        o, rl = lo.lindblad(Ho.toarray(), rho0, [c.toarray() for c in co], [e.toarray() for e in eo], Nt=12, dt=0.01)
        assert relerr(rho[b].cpu().numpy(), rl[-1]) <= TOL and relerr(obs[:, b].cpu().numpy(), o) <= TOL


@pytest.mark.parametrize('variant', [8])
def test_lindblad_tile_kernel_variants(cuda, monkeypatch, variant):
    """the tensor-memory-window variant of the register-patch kernel (LIMEB200_TILE_V=8: the thread's own rows of the
    stage vector and nothing else come from tensor memory, row coefficients from shared memory) does the same
    arithmetic: identical to the default kernel and within tolerance of the oracle, small cutoffs (one CTA, ragged
    padding, clusters of 2 and 4), with / without observables and trajectory, and on the full 4096-point batch at full
    occupancy"""
    import torch
    from lime_b200 import builders, oqs
    for ncav in (8, 16, 37, 64):
        H, c_ops, e_ops, rho0 = cases.jc_point(ncav=ncav)
        obs_o, rl_o = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=60, dt=0.01)
        Hs, cs = csr_matrix(H), [csr_matrix(c) for c in c_ops]
        out = {}
        for v in (0, variant):
            monkeypatch.setenv('LIMEB200_TILE_V', str(v))
            plan = oqs._lindblad_plan(Hs, cs, e_ops, path=6)
            assert plan.path == 6
            out[v] = plan.run(rho0, 0.01, 60, traj_every=20)
            plan2 = oqs._lindblad_plan(Hs, cs, None, path=6)          # no observables
            out[v] += (plan2.run(rho0, 0.01, 60)[0],)
        rho_f, obs, traj, rho_n = out[variant]
        assert relerr(obs, obs_o) <= TOL and relerr(rho_f, rl_o[-1]) <= TOL and relerr(rho_n, rl_o[-1]) <= TOL
        assert relerr(traj, np.array([rl_o[19], rl_o[39], rl_o[59]])) <= TOL
        for a, b in zip(out[variant], out[0]):
            assert np.array_equal(a, b)
    # no collapse operators (S = 0) and two of them (S = 2)
    H, c_ops, e_ops, rho0 = cases.jc_point(ncav=16)
    for cl in ([], [c_ops[0], 0.5 * c_ops[0]]):
        o, rl = lo.lindblad(H, rho0, cl, e_ops=e_ops, Nt=30, dt=0.01)
        monkeypatch.setenv('LIMEB200_TILE_V', str(variant))
        plan = oqs._lindblad_plan(csr_matrix(H), [csr_matrix(c) for c in cl], e_ops, path=6)
        rf, ob, _ = plan.run(rho0, 0.01, 30)
        assert relerr(ob, o) <= TOL and relerr(rf, rl[-1]) <= TOL
    # full batch, 200 steps: every CTA slot of the GPU busy, clusters of 4
    pat, vals, c_ops, e_ops, rho0 = builders.jc_grid()
    res = {}
    for v in (0, variant):
        monkeypatch.setenv('LIMEB200_TILE_V', str(v))
        plan, B = oqs._lindblad_plan_batch((pat, vals), c_ops, e_ops)
        assert plan.path == 6
        rho = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(rho0, (4096, 128, 128)))).cuda()
        obs, _ = plan.run_device(rho, 0.01, 200)
        res[v] = (rho, obs)
        del plan
    assert torch.equal(res[0][0], res[variant][0]) and torch.equal(res[0][1], res[variant][1])


def test_heom_config3_long_run(cuda):
    """config 3 (spin-boson, K = 2, depth 12, dt = 0.01): 4000 RK4 steps of the single hierarchy stay within
    1e-10 of the oracle (thread-per-ADO kernel, one launch)"""
    from lime_b200.heom.heom import HEOM
    H, sz, rho0, depth, K, lam, gam, T = cases.spin_boson_heom(depth=12)
    h = HEOM(H, sz, lam, gam, T, N_exp=K, N_cut=depth)
    res = h.evolve(rho0, 0.01, 4000, e_ops=[sz], store_states=False)
    st = h.states.astype(np.int64)
    ado_o, obs_o, _ = lo.heom_rk4(h.initial(rho0), H, h.Q, h.qmap, h.c, h.nu, st, h.dn.astype(np.int64),
                                  h.up.astype(np.int64), 0.01, 4000, e_ops=[sz])
    assert relerr(res.ado, ado_o) <= TOL and relerr(res.observables, obs_o) <= TOL
    assert abs(np.trace(res.ado[0]) - 1) < 1e-11


def test_heom_config3_1e5_steps_one_launch(cuda):
    """config 3 at its named length: 1e5 RK4 steps of the 91-ADO hierarchy in ONE launch.  The oracle needs minutes for
    that, so the full length is checked through properties: splitting the run into 25 launches of 4000 steps (whose
    first segment IS checked against the oracle above) gives the bit-identical state, the trace stays 1, tier 0 stays
    Hermitian and the populations approach a stationary value"""
    from lime_b200.heom.heom import HEOM
    H, sz, rho0, depth, K, lam, gam, T = cases.spin_boson_heom(depth=12)
    h = HEOM(H, sz, lam, gam, T, N_exp=K, N_cut=depth)
    one, obs1, _ = h.plan.run(h.initial(rho0), 0.01, 100000, e_ops=[sz])
    seg = h.initial(rho0)
    for _ in range(25):
        seg, _, _ = h.plan.run(seg, 0.01, 4000)
    assert np.array_equal(one, seg)
    assert abs(np.trace(one[0]) - 1) < 1e-10 and relerr(one[0], one[0].conj().T) < 1e-12
    assert abs(obs1[-1, 0] - obs1[-2000, 0]) < 1e-6 and np.all(np.isfinite(one))


def test_heom_parameter_batch_thread_per_ado_kernel(cuda):
    """batch large enough (>= 2 hierarchies per SM) to take the thread-per-ADO kernel; per-hierarchy bath parameters"""
    from lime_b200 import engine
    from lime_b200.heom.heom import _calc_matsubara_params
    H, sz, rho0, depth, K, lam, gam, T = cases.spin_boson_heom(depth=6)
    states, dn, up = engine.heom_tables([depth + 1] * K, depth)
    B = 600
    lams = np.linspace(0.05, 0.4, B)
    Ts = np.linspace(0.6, 2.0, B)
    cs, nus = [], []
    for lam_b, T_b in zip(lams, Ts):
        c, nu = _calc_matsubara_params(K, lam_b, gam, T_b)
        cs.append(c)
        nus.append(nu)
    plan = engine.HeomPlan(H, sz, [0] * K, np.array(cs), np.array(nus), states, dn, up)
    ado0 = np.zeros((B, states.shape[0], 2, 2), dtype=complex)
    ado0[:, 0] = rho0
    sx = np.array([[0, 1.], [1, 0]])
    out, obs, traj = plan.run(ado0, 0.01, 30, e_ops=[sz, sx], traj_every=10)
    assert plan.path == 1
    for b in (0, 299, 599):
        ado_o, obs_o, tr_o = lo.heom_rk4(ado0[b], H, sz[None], [0] * K, cs[b], nus[b], states.astype(np.int64),
                                         dn.astype(np.int64), up.astype(np.int64), 0.01, 30, e_ops=[sz, sx], store=True)
        assert relerr(out[b], ado_o) <= TOL and relerr(obs[:, b], obs_o) <= TOL
        assert relerr(traj[:, b], np.array([tr_o[9], tr_o[19], tr_o[29]])) <= TOL


def test_liouvillian_eigen_solver(cuda):
    """superoperator.Lindblad_solver (lime/superoperator.py:456-773).  Frozen reference outputs
    (another host may round the LAPACK eigenvectors differently: 1e-8) and the oracle on this host (1e-10);
    the device part is the DMMA ZGEMM"""
    from lime_b200.superoperator import Lindblad_solver
    from lime_b200 import engine
    g = golden('super_lindblad')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=3, M=2, E=2, seed=61)
    A, B, C = g['A'], g['B'], g['C']
    tl, taul, wl = g['tl'], g['taul'], g['wl']
    s = Lindblad_solver(H, c_ops)
    with pytest.raises(TypeError):
        s.evolve(rho0, tl, e_ops)
    s.eigenstates()
    o = lo.SuperLindblad(H, c_ops)
    o.eigenstates()
    got = {'evolve': s.evolve(rho0, tl, e_ops).observables,
           'c2_1t': s.correlation_2op_1t(rho0, [A, B], tl), 'c2_1w': s.correlation_2op_1w(rho0, [A, B], wl),
           'c3_1t': s.correlation_3op_1t(rho0, [A, B, C], tl), 'c3_1w': s.correlation_3op_1w(rho0, [A, B, C], wl),
           'c3_2t': s.correlation_3op_2t(rho0, [A, B, C], tl, taul),
           'c4_2t': s.correlation_4op_2t(rho0, [A, B, C, A], tl, taul)}
    want = {'evolve': o.evolve(rho0, tl, e_ops),
            'c2_1t': o.correlation_2op_1t(rho0, [A, B], tl), 'c2_1w': o.correlation_2op_1w(rho0, [A, B], wl),
            'c3_1t': o.correlation_3op_1t(rho0, [A, B, C], tl), 'c3_1w': o.correlation_3op_1w(rho0, [A, B, C], wl),
            'c3_2t': o.correlation_3op_2t(rho0, [A, B, C], tl, taul),
            'c4_2t': o.correlation_4op_2t(rho0, [A, B, C, A], tl, taul)}
    for k in got:
        assert got[k].shape == g[k].shape, k
        assert relerr(got[k], want[k]) <= TOL, k
        assert relerr(got[k], g[k]) <= 1e-8, k
    assert got['c3_2t'].shape == (len(taul), len(tl))
    with pytest.raises(ValueError):
        s.correlation_4op_2t(rho0, [A, B, C], tl, taul)
    # the GEMM primitive itself, ragged and batched shapes
    rng = np.random.default_rng(3)
    for (m, n, k, b) in [(1, 1, 1, 1), (70, 33, 129, 1), (64, 64, 64, 3), (5, 200, 17, 2)]:
        a = rng.standard_normal((b, m, k)) + 1j * rng.standard_normal((b, m, k))
        bb = rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n))
        c = engine.zgemm(a, bb).cpu().numpy()
        assert relerr(c, a @ bb) <= 1e-13


def test_sesolver_wavefunction_path(cuda):
    """lime.mol.SESolver (RK4 of the Schroedinger equation and its correlation functions)
    against frozen outputs of the reference and the oracle"""
    from lime_b200.mol import SESolver
    g = golden('sesolver')
    H = cases.rand_herm(5, 81)
    psi0 = cases.rand_cplx(5, 82)[:, 0]
    psi0 = psi0 / np.linalg.norm(psi0)
    e_ops = [cases.rand_herm(5, 83), cases.rand_herm(5, 84)]
    ops = [g['A'], g['B'], g['C']]
    s = SESolver(H)
    r = s.run(psi0=psi0, dt=0.01, Nt=40, e_ops=e_ops, nout=2)
    assert r.observables.shape == (20, 2) and len(r.psilist) == 20
    assert relerr(r.observables, g['obs']) <= TOL and relerr(np.array(r.psilist), g['psilist']) <= TOL
    assert relerr(s.correlation_3op_1t(psi0, ops, 0.01, 12), g['c3_1t']) <= TOL
    assert relerr(s.correlation_3op_2t(psi0, ops, 0.01, 5, 6), g['c3_2t']) <= TOL
    assert relerr(s.correlation_4op_2t(psi0, ops + [ops[0]], 0.01, 4, 3), g['c4_2t']) <= TOL
    U = s.propagator(0.01, 6)
    import scipy.linalg
    assert len(U) == 6 and relerr(U[5], scipy.linalg.expm(-1j * H * 0.05)) <= 1e-9      # RK4 truncation error
    with pytest.raises(ValueError):                      # laser-driven branch: edip is mandatory (lime/mol.py:1153-1156)
        s.run(psi0=psi0, pulse=object())


def test_sos_mol_wrappers_and_tpa(cuda, tmp_path):
    """photon_echo / photon_echo_t3 (Mol-based wrappers with their np.savez side effect) and the single-frequency
    TPA, lime/signal/sos.py:199-228, 731-902, against frozen reference outputs"""
    from lime_b200.signal import sos
    g = golden('sos_mol')
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    mol = cases.DuckMol(E, dip, gamma, dephasing=0.01 / 27.211386)
    wp = g['wp']
    pe = sos.photon_echo(mol, wp, wp, t2=30.0, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, fname=str(tmp_path / 's'))
    assert relerr(pe, g['PE']) <= TOL
    saved = np.load(str(tmp_path / 's.npz'))
    assert np.array_equal(saved['arr_2'], pe)
    t3 = sos.photon_echo_t3(mol, wp, wp, 20.0, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, fname=str(tmp_path / 't'))
    assert relerr(t3, g['PE_t3']) <= TOL
    se, esa = sos.photon_echo_t3(mol, wp, wp, 20.0, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, fname=None, separate=True)
    assert relerr(se + esa, g['PE_t3']) <= TOL
    tpa = np.array([sos.TPA(E, dip, w, g_idx, e_idx, f_idx, gamma) for w in g['wtpa']])
    assert relerr(tpa, g['TPA']) <= TOL
    mol.gamma = None
    with pytest.raises(ValueError):
        sos.photon_echo(mol, wp, wp)


def test_fft_module(cuda):
    """lime/fft.py -- fft, ifft, fft2 as ONE tensor-core GEMM with the Fourier matrix (shift, scale and
    phase folded in; no FFT library), dft, dft2 as separable GEMMs -- against frozen reference outputs"""
    from lime_b200 import fft as lfft
    g = golden('fft')
    xg, f1, f2 = g['xg'], g['f1'], g['f2']
    out, freq = lfft.fft(f1, xg)
    assert relerr(out, g['fft']) <= 1e-12 and np.array_equal(freq, g['fft_freq'])
    out0, _ = lfft.fft(f1.T.copy(), xg, axis=0)
    assert relerr(out0, g['fft'].T) <= 1e-12
    out, freq = lfft.ifft(f1[0], xg)
    assert relerr(out, g['ifft']) <= 1e-12 and np.array_equal(freq, g['ifft_freq'])
    fx, fy, out = lfft.fft2(f2, 0.1, 0.2)
    assert relerr(out, g['fft2']) <= 1e-12 and np.array_equal(fy, g['fft2_fy'])
    assert relerr(lfft.dft(xg, f1[1], g['kxs']), g['dft']) <= 1e-12
    assert relerr(lfft.dft2(g['xs'], g['ys'], g['fxy'], g['kxs'], g['kys']), g['dft2']) <= 1e-12
    # odd lengths, a 2-D input through ifft (lime shifts every axis), rectangular fft2 -- against the NumPy formulas of lime
    rng = np.random.default_rng(6)
    for n in (7, 33):
        x = np.linspace(0.3, 2.0, n)
        f = rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))
        dx = x[1] - x[0]
        fr = 2 * np.pi * np.fft.fftshift(np.fft.fftfreq(n, d=dx))
        ref = np.fft.fftshift(np.fft.fft(f, axis=-1), axes=(-1,)) * dx * np.exp(-1j * fr * x[0])
        assert relerr(lfft.fft(f, x)[0], ref) <= 1e-12
        fi = 2 * np.pi * np.fft.ifftshift(np.fft.fftfreq(n, d=dx))
        ref = np.fft.ifftshift(np.fft.ifft(f, axis=-1)) * dx / 2 / np.pi * n * np.exp(1j * fi * x[0])
        assert relerr(lfft.ifft(f, x)[0], ref) <= 1e-12
        f22 = rng.standard_normal((n, n + 2)) + 0j
        assert relerr(lfft.fft2(f22, 0.1, 0.3)[2], np.fft.fftshift(np.fft.fft2(f22)) * 0.03) <= 1e-12


def test_sesolver_laser_driven(cuda, tmp_path, monkeypatch):
    """SESolver.run(pulse=...) / driven_dynamics (lime/mol.py:1094-1171, 1473-1588): one launch of the driven kernel,
    single pulse with nout = 2 (lime holds H at the block's start time), two pulses, sparse and dense operands, the
    psi.dat / obs.dat branch"""
    from scipy.sparse import csr_matrix
    from lime_b200 import mol
    from test_oracle_vs_golden import _GoldPulse
    g = golden('sesolver_driven')
    p1, p2 = [_GoldPulse(*r) for r in g['pulses']]
    e_d = [g['e0'], g['e1']]
    r = mol.SESolver(csr_matrix(g['Hd'])).run(psi0=g['psi0'], dt=0.01, Nt=40, e_ops=[csr_matrix(e) for e in e_d], nout=2,
                                              edip=csr_matrix(g['mu']), pulse=p1)
    assert relerr(r.observables, g['obs1']) <= TOL and relerr(np.array(r.psilist), g['psi1']) <= TOL
    assert np.array_equal(r.times, 0.0 + np.arange(20) * 0.02)
    r = mol.SESolver(g['Hd']).run(psi0=g['psi0'], dt=0.01, Nt=30, e_ops=e_d, nout=1, edip=[g['mu'], g['mu2']], pulse=[p1, p2])
    assert relerr(r.observables, g['obs2']) <= TOL and relerr(np.array(r.psilist), g['psi2']) <= TOL
    with pytest.raises(ValueError):
        mol.SESolver(g['Hd']).run(psi0=g['psi0'], pulse=p1)
    monkeypatch.chdir(tmp_path)
    assert mol.driven_dynamics([g['Hd'], [g['mu'], p1.efield]], g['psi0'], dt=0.01, Nt=6, e_ops=e_d, return_result=False) is None
    rows = np.genfromtxt(tmp_path / 'obs.dat', dtype=complex)
    o, _ = lo.driven_dynamics([g['Hd'], [g['mu'], p1.efield]], g['psi0'], dt=0.01, Nt=8, e_ops=e_d, nout=1)
    assert rows.shape == (6, 3) and relerr(rows[:, 1:], o[1:7]) <= TOL


def test_etpa_double_time_integrals(cuda):
    """sos._etpa (lime/signal/sos.py:1171-1223): tensor-core GEMM + reduction kernel against the frozen reference output,
    and a rectangular-free larger grid against the oracle loops"""
    from lime_b200.signal import sos
    g = golden('etpa')
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    out = sos._etpa(g['wps'], E, dip, g['jta'], g['t1'], g['t2'], list(g_idx), list(e_idx), list(f_idx))
    assert relerr(out, g['etpa']) <= TOL
    rng = np.random.default_rng(4)
    nt = 70
    t = np.linspace(-30, 90, nt)
    jta = rng.standard_normal((nt, nt)) + 1j * rng.standard_normal((nt, nt))
    wps = np.linspace(0.1, 0.2, 9)
    ref = lo.etpa_core(wps, E, dip, jta, t, t, list(g_idx), list(e_idx), list(f_idx))
    assert relerr(sos._etpa(wps, E, dip, jta, t, t, list(g_idx), list(e_idx), list(f_idx)), ref) <= TOL
    assert np.array_equal(sos._etpa(wps, E, dip, jta, t, t, list(g_idx), [], list(f_idx)), np.zeros(9))
    with pytest.raises(ValueError):
        sos._etpa(wps, E, dip, jta, t, t, [0, 1], list(e_idx), list(f_idx))


def test_phys_algebra_helpers_device_path(cuda):
    """comm / anticomm / commutator / anticommutator (lime/phys.py:736-752): dense operands from N = 64 on take the
    batched FP64 tensor-core GEMM; small, sparse and mismatched operands behave as in lime"""
    from scipy.sparse import csr_matrix
    from lime_b200 import phys
    A, B = cases.rand_cplx(96, 1), cases.rand_cplx(96, 2)
    assert relerr(phys.comm(A, B), A @ B - B @ A) <= 1e-13
    assert relerr(phys.anticomm(A, B), A @ B + B @ A) <= 1e-13
    assert relerr(phys.commutator(A, B), A @ B - B @ A) <= 1e-13
    assert relerr(phys.anticommutator(A, B), A @ B + B @ A) <= 1e-13
    Ar, Br = A.real.copy(), B.real.copy()
    c = phys.comm(Ar, Br)
    assert c.dtype == np.float64 and relerr(c, Ar @ Br - Br @ Ar) <= 1e-13
    a, b = cases.rand_cplx(5, 3), cases.rand_cplx(5, 4)
    assert np.array_equal(phys.comm(a, b), np.dot(a, b) - np.dot(b, a))
    sa, sb = csr_matrix(a), csr_matrix(b)
    assert relerr(phys.commutator(sa, sb).toarray(), a @ b - b @ a) <= 1e-15
    with pytest.raises(AssertionError):
        phys.comm(A, a)
    assert np.array_equal(phys.dag(a), a.conj().T)
    # lime-style user loop over rk4(rho, liouvillian, ...) (lime/phys.py:100-105): the operator plan is built once
    from lime_b200 import oqs
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=6)
    oqs.clear_plan_cache()
    rho = rho0.copy()
    for _ in range(5):
        out = phys.rk4(rho, oqs.liouvillian, 0.01, H, c_ops)
        assert out is rho
    assert len(oqs._PLAN_CACHE) == 1
    _, rl = lo.lindblad(H, rho0, c_ops, [], Nt=5, dt=0.01)
    assert relerr(rho, rl[-1]) <= TOL
    H2 = H.copy()
    H2[0, 0] += 0.5                                   # a changed operator is a different plan (content fingerprint)
    oqs.liouvillian(rho0, H2, c_ops)
    assert len(oqs._PLAN_CACHE) == 2
    assert relerr(oqs.liouvillian(rho0, H2, c_ops), lo.liouvillian(rho0, H2, c_ops)) <= 1e-13


# ---------------------------------------------------------------- RKF45 (north_star; lime ships only examples/rkf45_test.py)
def test_rkf45_lime_example_problems(cuda):
    """examples/rkf45_test.py:72-119 (test04), :153-202 (test05), :283-330 (test06, single-step mode) through the
    device stage kernels; tolerances sqrt(eps) as in the example, errors against the analytic solutions"""
    from lime_b200 import rkf45 as rk
    tol = float(np.sqrt(np.finfo(np.double).eps))

    def f1(t, y):
        return np.array([0.25 * y[0] * (1.0 - y[0] / 20.0)])

    def exact1(t):
        return 20.0 / (1.0 + 19.0 * np.exp(-0.25 * t))

    y, flag, t = np.array([1.0]), 1, 0.0
    yp = f1(t, y)
    for i in range(1, 6):
        t, tout = (i - 1) * 4.0, i * 4.0
        y, yp, t, flag = rk.r8_rkf45(f1, 1, y, yp, t, tout, tol, tol, flag)
        assert flag == 2 and t == tout and abs(y[0] - exact1(t)) < 2e-6

    def f2(t, y):
        return np.array([y[1], -y[0]])
    y, flag = np.array([1.0, 0.0]), 1
    yp = f2(0.0, y)
    for i in range(1, 13):
        t, tout = (i - 1) * 2 * 3.14159265 / 12, i * 2 * 3.14159265 / 12
        y, yp, t, flag = rk.r8_rkf45(f2, 2, y, yp, t, tout, tol, tol, flag)
        assert flag == 2 and abs(y[0] - np.cos(t)) < 1e-6 and abs(y[1] + np.sin(t)) < 1e-6

    y, flag, t = np.array([1.0]), -1, 0.0
    yp = f1(t, y)
    for i in range(1, 6):
        tout = i * 4.0
        while flag < 0:
            y, yp, t, flag = rk.r8_rkf45(f1, 1, y, yp, t, tout, tol, tol, flag)
        assert flag == 2 and t == tout and abs(y[0] - exact1(t)) < 2e-6
        flag = -2
    # one uncontrolled Fehlberg step: fifth-order-pair solution of y' = y over h = 0.1
    fs = rk.r8_fehl(lambda t, y: y, 1, np.array([1.0]), 0.0, 0.1, np.array([1.0]))
    assert abs(fs[5][0] - np.exp(0.1)) < 1e-9


def test_rkf45_device_resident_lindblad_and_heom(cuda):
    """adaptive propagation with rho resident on the device: the right-hand side is limeb200_qme_rhs /
    limeb200_heom_rhs, the stage arithmetic limeb200_rkf45_*; compared with the exact propagator e^{Lt} (Lindblad,
    lime's superoperator) and with fine-step RK4 of the oracle (HEOM)"""
    import torch
    import scipy.linalg
    from lime_b200 import rkf45 as rk, oqs, engine
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=6)
    plan = oqs._lindblad_plan(H, c_ops, None)
    B = 3
    r0 = np.stack([rho0, cases.rand_dm(6, 91), cases.rand_dm(6, 92)])
    d = torch.from_numpy(r0).cuda()
    tl = np.linspace(0.0, 2.0, 5)
    states, S = rk.integrate(rk.QmeRHS(plan), d, tl, relerr=1e-10, abserr=1e-12)
    Lsup = lo.liouvillian_super(H, c_ops).toarray()
    for t, s in zip(tl, states):
        U = scipy.linalg.expm(Lsup * t)
        for b in range(B):
            ref = (U @ r0[b].reshape(-1)).reshape(6, 6)
            assert relerr(s[b].cpu().numpy(), ref) < 1e-8
    assert S.steps_accepted > 4 and S.nfe > 0
    # HEOM hierarchy, adaptive, against the oracle's RK4 with a small step
    Hs, sz, rr, depth, K, lam, gam, T = cases.spin_boson_heom(depth=4)
    from lime_b200.heom.heom import HEOM
    h = HEOM(Hs, sz, lam, gam, T, N_exp=K, N_cut=depth)
    a0 = torch.from_numpy(h.initial(rr)[None]).cuda()
    out, S = rk.integrate(rk.HeomRHS(h.plan), a0, [0.0, 0.5, 1.0], relerr=1e-10, abserr=1e-13)
    st = h.states.astype(np.int64)
    ado_o, _, _ = lo.heom_rk4(h.initial(rr), Hs, h.Q, h.qmap, h.c, h.nu, st, h.dn.astype(np.int64), h.up.astype(np.int64),
                              0.0005, 2000)
    assert relerr(out[-1][0].cpu().numpy(), ado_o) < 1e-8


def test_batch_apis_obs_every(cuda):
    """[ext] obs_every = k on the batch APIs returns exactly the observables after steps k, 2k, ... of the full run
    (sub-sampled on the device before the copy; the final states are untouched)"""
    from lime_b200.oqs import Lindblad_solver, Redfield_solver
    H, c_ops, e_ops, rho0 = cases.jc_point(ncav=8)
    s = Lindblad_solver(H, c_ops=c_ops)
    batch = np.stack([rho0, cases.rand_dm(16, 3)])
    rf, obs, _ = s.evolve_batch(batch, 0.01, 12, e_ops=e_ops)
    for pinned in (False, True):
        rf4, obs4, _ = s.evolve_batch(batch, 0.01, 12, e_ops=e_ops, obs_every=4, pinned=pinned)
        assert obs4.shape == (3, 2, 2) and np.array_equal(obs4, obs[3::4]) and np.array_equal(rf4, rf)
    Hs, a_ops, spectra, r0, dt, Nt, e_r, tlist = cases.redfield_example()
    rs = Redfield_solver(Hs, c_ops=a_ops, spectra=spectra)
    for form in ('tensor', 'operator'):
        out, ob = rs.evolve_batch(np.stack([r0, r0]), dt, 12, e_ops=e_r, form=form)
        out3, ob3 = rs.evolve_batch(np.stack([r0, r0]), dt, 12, e_ops=e_r, form=form, obs_every=3)
        assert ob3.shape == (4, 2, 1) and np.array_equal(ob3, ob[2::3]) and np.array_equal(out3, out)

