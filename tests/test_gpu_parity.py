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
