"""CPU: the NumPy oracle reproduces the frozen outputs of the real reference
(tests/golden/*.npz, written by oracle/gen_golden.py) and the reference's own golden
examples/cor.dat + dm.dat."""
import json
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

import cases
import lime_oracle as lo
from conftest import golden, relerr, GOLD

TOL = 1e-13


def test_pinning_report_is_clean():
    rep = json.load(open(os.path.join(GOLD, 'PINNING.json')))['oracle_vs_reference_max_rel_err']
    for case, errs in rep.items():
        for k, v in errs.items():
            if k in ('fixed_vs_strict', 'fmo_nhe', 'none_raises_TypeError', 'cases'):
                continue
            assert v <= (1e-14 if case != 'heom_rules' else 1e-15 * 4), (case, k, v)
    assert rep['heom_tables']['none_raises_TypeError'] == 1
    assert rep['heom_tables']['fmo_nhe'] == 3060


def test_reference_golden_cavity_bit_exact():
    g = golden('cavity_cor')
    H, rho0, ops, c_ops, tlist = cases.thermal_cavity()
    t, cor, dm = lo.correlation_3p_1t(H, rho0, ops, c_ops, tlist)
    assert np.array_equal(t, g['t'])
    assert np.array_equal(cor, g['cor'])
    assert np.array_equal(dm, g['dm'])


def test_redfield_example():
    g = golden('redfield_example')
    H, a_ops, spectra, rho0, dt, Nt, e_ops, tlist = cases.redfield_example()
    R, evecs = lo.redfield_tensor(H, a_ops, spectra)
    assert relerr(R.toarray(), g['R']) <= TOL
    obs, rl = lo.redfield(R, rho0, evecs=evecs, Nt=Nt, dt=dt, e_ops=e_ops)
    assert relerr(obs, g['observables']) <= TOL
    assert relerr(rl, g['rholist']) <= TOL
    t8 = tlist[:8]
    U = lo.redfield_propagator(R, t8, 'SOS')
    assert relerr(U, g['U_sos']) <= 1e-12
    assert relerr(lo.redfield_propagator(R, t8, 'EOM'), g['U_eom']) <= TOL
    assert relerr(lo.redfield_correlation_4op_3t(-1j * U, 2, rho0, [e_ops[0]] * 4, 'llll'), g['corr4']) <= 1e-12
    # operator form == tensor form
    evals, ev2, A, Lam = lo.redfield_parts(H, a_ops, spectra)
    rho = cases.rand_dm(2, 3)
    G = -1j * np.diag(evals)
    for a, l in zip(A, Lam):
        G = G - a @ l
    k = G @ rho + rho @ G.conj().T
    for a, l in zip(A, Lam):
        k = k + a @ rho @ l.conj().T + l @ rho @ a
    assert relerr(k.reshape(-1), R.dot(rho.reshape(-1))) <= 1e-13


def test_redfield_multilevel():
    g = golden('redfield_multilevel')
    H, a_ops, spectra, rho0 = cases.redfield_multilevel()
    R, evecs = lo.redfield_tensor(H, a_ops, spectra)
    obs, rl = lo.redfield(R, rho0, evecs=evecs, Nt=60, dt=0.02, e_ops=[a_ops[0], g['e1']])
    assert relerr(R.toarray(), g['R']) <= TOL
    assert relerr(obs, g['observables']) <= TOL and relerr(rl, g['rholist']) <= TOL


def test_lindblad_dense_and_superoperator():
    g = golden('lindblad_dense')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense()
    obs, rl = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=80, dt=0.01)
    assert relerr(obs, g['observables']) <= TOL and relerr(rl, g['rholist']) <= TOL
    assert relerr(lo.liouvillian(rho0, H, c_ops), g['rhs']) <= TOL
    L = lo.liouvillian_super(H, c_ops)
    assert relerr(L.toarray(), g['superop']) <= TOL
    assert relerr(L.dot(rho0.flatten()), g['rhs'].flatten()) <= 1e-14
    # invariants the reference has no tests for: trace and hermiticity are conserved
    assert abs(np.trace(rl[-1]) - 1) < 1e-12
    assert relerr(rl[-1], rl[-1].conj().T) < 1e-13


def test_lindblad_jc_and_driven_and_correlations():
    g = golden('lindblad_jc')
    H, c_ops, e_ops, rho0 = cases.jc_point(ncav=8)
    obs, rl = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=100, dt=0.01)
    assert relerr(obs, g['observables']) <= 1e-12
    assert relerr(rl[-1], g['rho_final']) <= 1e-12 and relerr(rl[49], g['rho_mid']) <= 1e-12
    Ho, co, eo = lo.jaynes_cummings(1.0, 1.05, 0.1, 8, 0.05)
    assert relerr(Ho.toarray(), H) == 0 and relerr(co[0].toarray(), c_ops[0]) == 0

    g = golden('lindblad_driven')
    H0, c_ops, e_ops, rho0 = cases.lindblad_dense(n=4, M=1, E=1, seed=77)

    def f1(t):
        return 0.3 * np.exp(-(t - 0.4) ** 2 / 0.02) * np.exp(-1j * 2.0 * t)
    obs, rl = lo.lindblad_driven([H0.copy(), [g['H1'], f1]], rho0, c_ops, e_ops, Nt=60, dt=0.01, t0=0.1,
                                 strict_parity=True)
    assert relerr(obs, g['observables_strict']) <= TOL and relerr(rl[-1], g['rho_final_strict']) <= TOL

    g = golden('lindblad_corr')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=4, M=2, E=1, seed=31)
    ops3 = [g['A'], g['B'], g['C']]
    assert relerr(lo.lindblad_correlation_3op_1t(H, c_ops, rho0, ops3, 0.02, 30), g['c3op1t']) <= TOL
    assert relerr(lo.lindblad_correlation_3op_2t(H, c_ops, rho0, ops3, 0.02, 6, 7), g['c3op2t']) <= TOL


def test_heom_dl_tables_matsubara():
    g = golden('heom_dl')
    H, sz, rho0 = cases.spin_boson_heom()[:3]
    ado, traj = lo.heom_dl(H, rho0, sz, 300.0, 0.002, 0.0005, 12, 0.05, 200)
    assert relerr(ado[:, :, 0], g['rho_final']) <= TOL and relerr(traj, g['traj']) <= TOL

    g = golden('heom_tables')
    for key in g.files:
        dpart, x = key[1:].split('_x')
        dims = [int(v) for v in dpart.split('_')]
        exc = int(x)
        n, s2i, i2s = lo.enr_state_dictionaries(dims, exc)
        ref = g[key]
        assert n == ref.shape[0]
        arr = np.array([i2s[i] for i in range(n)]).reshape(n, len(dims))
        assert np.array_equal(arr, ref)
        assert all(s2i[tuple(r)] == i for i, r in enumerate(arr))
        if exc:
            assert np.array_equal(lo.enr_states_fast(dims, exc), ref)
    assert g['d13_13_x12'].shape[0] == 91
    assert g['d' + '_'.join(['5'] * 14) + '_x4'].shape[0] == 3060
    with pytest.raises(TypeError):
        lo.enr_state_dictionaries([2, 2], None)

    g = golden('heom_matsubara')
    for i in range(4):
        K, lam, gam, T = g['par%d' % i]
        c, nu = lo.calc_matsubara_params(int(K), lam, gam, T)
        assert relerr(c, g['c%d' % i]) <= 1e-15 and relerr(nu, g['nu%d' % i]) <= 1e-15


def test_heom_connectivity_counts():
    # SURVEY.md section 8(a13/a15): 312 directed couplings for [13]*2 depth 12, 19040 for FMO
    st, dn, up = lo.heom_tables([13, 13], 12)
    assert st.shape[0] == 91 and (dn >= 0).sum() + (up >= 0).sum() == 312
    # every down edge is the reverse of an up edge
    for a in range(st.shape[0]):
        for k in range(2):
            if dn[a, k] >= 0:
                assert up[dn[a, k], k] == a


HEOM_RULE_CASES = ['k2_d3_n2', 'k4_d2_n3', 'b2k2_d3_n3', 'b2k2_d2_proj']


def heom_rule_case(name):
    """inputs of one frozen case of tests/golden/heom_rules.npz: the reference's own hierarchy Liouvillian L_helems
    (lime/heom/heom.py:156-216 exec'd by oracle/gen_golden.py) and the bath data it was built from"""
    g = golden('heom_rules')
    n, N_m, N_c, ncoup = [int(x) for x in g[name + '_meta']]
    st, dn, up = lo.heom_tables([N_c + 1] * N_m, N_c)
    return dict(L=g[name + '_L'], Q=g[name + '_Q'], qmap=[int(x) for x in g[name + '_qmap']], c=g[name + '_c'],
                nu=g[name + '_nu'], n=n, N_m=N_m, N_c=N_c, ncoup=ncoup, st=st, dn=dn, up=up)


@pytest.mark.parametrize('name', HEOM_RULE_CASES)
def test_heom_rhs_pinned_on_reference_coupling_rules(name):
    """PINS the multi-index HEOM right-hand side: bath part of lo.heom_rhs == (the matrix the reference's own rule
    loop builds) . vec, for lime's single-Q case and for bath-dependent coupling operators"""
    k = heom_rule_case(name)
    assert k['ncoup'] == int((k['dn'] >= 0).sum() + (k['up'] >= 0).sum())       # ADO connectivity, bit-exact
    rng = np.random.default_rng(3)
    nhe, n = k['st'].shape[0], k['n']
    for _ in range(3):
        ado = rng.standard_normal((nhe, n, n)) + 1j * rng.standard_normal((nhe, n, n))
        ref = (k['L'] @ ado.reshape(-1)).reshape(nhe, n, n)
        got = lo.heom_rhs(ado, np.zeros((n, n), dtype=complex), k['Q'], k['qmap'], k['c'], k['nu'],
                          k['st'], k['dn'], k['up'])
        assert relerr(got, ref) <= TOL
    # block structure: L couples ADO a to b only along table edges (and the diagonal)
    nn = n * n
    blocks = np.abs(k['L']).reshape(nhe, nn, nhe, nn).max(axis=(1, 3)) > 0
    allowed = np.eye(nhe, dtype=bool)
    for a in range(nhe):
        for m in range(k['N_m']):
            for t in (k['dn'][a, m], k['up'][a, m]):
                if t >= 0:
                    allowed[a, t] = True
    assert not np.any(blocks & ~allowed)


def test_heom_rhs_matches_heom_dl_structure():
    """the multi-index RHS with one mode, pref_dn=1, pref_up=-1, c=a+ib reduces to the tier
    equations _heom_dl integrates (lime/oqs.py:1853-1857) -- ties the multi-index restatement
    to the pinned one."""
    H, sz, rho0 = cases.spin_boson_heom()[:3]
    nado = 6
    gamma, a = 0.7, 0.3
    st, dn, up = lo.heom_tables([nado], nado - 1)
    rng = np.random.default_rng(0)
    ado = rng.standard_normal((nado, 2, 2)) + 1j * rng.standard_normal((nado, 2, 2))
    k = lo.heom_rhs(ado, H, sz[None], [0], [a], [gamma], st, dn, up, pref_dn=1.0, pref_up=-1.0)
    for n in range(1, nado - 1):
        ref = -1j * lo.commutator(H, ado[n]) - lo.commutator(sz, ado[n + 1]) - n * gamma * ado[n] \
            + n * a * lo.commutator(sz, ado[n - 1])
        assert relerr(k[n], ref) < 1e-14


def test_round2_api_rows(tmp_path):
    """_correlation_2p_1t (lime/oqs.py:726-800, incl. the cor.dat text) and getG time / frequency (lime/oqs.py:474-526)"""
    g = golden('api_r2')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=6)
    fn = str(tmp_path / 'cor.dat')
    cor = lo.correlation_2p_1t(H, rho0, [g['A'], g['B']], c_ops, 0.01, 30, output=fn)
    assert relerr(cor, g['cor2']) <= TOL
    assert open(fn).read() == str(g['cor_txt'])
    Hj, cj, ej, rj = cases.jc_point(ncav=8)
    assert relerr(lo.correlation_2p_1t(Hj, rj, [ej[0], cj[0] / np.sqrt(0.05)], cj, 0.01, 25), g['cor2_jc']) <= TOL
    Hr, a_ops, spectra, rr = cases.redfield_multilevel()
    R, _ = lo.redfield_tensor(Hr, a_ops, spectra)
    assert relerr(lo.getG(1j * R, g['tg']), g['G_time']) <= 1e-12
    assert relerr(lo.getG(1j * R, g['tg'], w=g['wg'], domain='freq'), g['G_freq']) <= 1e-12


class _GoldPulse:
    def __init__(self, a, w, tc, sig):
        self.a, self.w, self.tc, self.sig = a, w, tc, sig

    def efield(self, t):
        return self.a * np.exp(-(t - self.tc) ** 2 / 2. / self.sig ** 2) * np.exp(-1j * self.w * (t - self.tc))


def test_lindblad_driven_csr_operands():
    """_lindblad_driven with scipy.sparse operands (lime/oqs.py:1691-1800): `Ht += ...` rebinds, no accumulation"""
    g = golden('lindblad_driven_csr')
    Hj, cj, ej, rj = cases.jc_point(ncav=6)
    fdr = lambda t: 0.2 * np.exp(-(t - 0.2) ** 2 / 0.02) * np.cos(3 * t)
    o, rl = lo.lindblad_driven([Hj.copy(), [g['H1'], fdr]], rj, cj, ej, Nt=40, dt=0.01, t0=0.05)
    assert relerr(o, g['obs']) <= TOL and relerr(rl[-1], g['rho_final']) <= TOL and relerr(rl[19], g['rho_mid']) <= TOL


def test_sesolver_driven():
    """SESolver.run(pulse=...) -> driven_dynamics, lime/mol.py:1094-1171, 1473-1560"""
    g = golden('sesolver_driven')
    p1, p2 = [_GoldPulse(*r) for r in g['pulses']]
    e_d = [g['e0'], g['e1']]
    o1, pl1 = lo.driven_dynamics([g['Hd'], [g['mu'], p1.efield]], g['psi0'], dt=0.01, Nt=40, e_ops=e_d, nout=2)
    assert relerr(o1, g['obs1']) <= TOL and relerr(np.array(pl1), g['psi1']) <= TOL
    o2, pl2 = lo.driven_dynamics([g['Hd'], [g['mu'], p1.efield], [g['mu2'], p2.efield]], g['psi0'], dt=0.01, Nt=30,
                                 e_ops=e_d, nout=1)
    assert relerr(o2, g['obs2']) <= TOL and relerr(np.array(pl2), g['psi2']) <= TOL


def test_etpa_double_time_integrals():
    """sos._etpa, lime/signal/sos.py:1171-1223"""
    g = golden('etpa')
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    out = lo.etpa_core(g['wps'], E, dip, g['jta'], g['t1'], g['t2'], list(g_idx), list(e_idx), list(f_idx))
    assert relerr(out, g['etpa']) <= TOL


def test_sos():
    g = golden('sos')
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    w1, w3, w2, w1b, t2 = g['w1'], g['w3'], g['w2'], g['w1b'], float(g['t2'])
    au2ev = 27.211386
    assert relerr(lo.GSB(E, dip, w1, w3, t2, g_idx, e_idx, gamma), g['GSB']) <= TOL
    assert relerr(lo.SE(E, dip, w1, w3, t2, g_idx, e_idx, gamma), g['SE']) <= TOL
    assert relerr(lo.ESA(E, dip, w1, w3, t2, g_idx, e_idx, f_idx, gamma), g['ESA']) <= TOL
    assert relerr(lo.photon_echo_core(E, dip, -w1, w3, t2, g_idx, e_idx, f_idx, gamma), g['PE']) <= TOL
    assert relerr(lo.SE_t3(E, dip, -w1, w3, t2, g_idx, e_idx, gamma, dephasing=0.01 / au2ev), g['SE_t3']) <= TOL
    assert relerr(lo.ESA_t3(E, dip, -w1, w3, t2, g_idx, e_idx, f_idx, gamma, dephasing=0.01 / au2ev), g['ESA_t3']) <= TOL
    kw = dict(g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    assert relerr(lo.DQC_R1(E, dip, omega1=w1b, omega2=w2, tau3=1e-6, **kw), g['R1_t3']) <= TOL
    assert relerr(lo.DQC_R2(E, dip, omega1=w1b, omega2=w2, tau3=1e-6, **kw), g['R2_t3']) <= TOL
    assert relerr(lo.DQC_R1(E, dip, omega2=w2, omega3=w1b, tau1=50.0, **kw), g['R1_t1']) <= TOL
    assert relerr(lo.DQC_R2(E, dip, omega2=w2, omega3=w1b, tau1=50.0, **kw), g['R2_t1']) <= TOL
    assert relerr(lo.TPA2D(E, dip, w2, w1b, g_idx, e_idx, f_idx, gamma), g['TPA2D']) <= TOL
    assert relerr(lo.TPA2D_time_order(E, dip, w2, w1b, g_idx, e_idx, f_idx, gamma), g['TPA2D_to']) <= TOL


def test_time_domain_2des_oracle_matches_reference_functions():
    """golden = the reference's own ESA/GSB/SE source (lime/signal/2DES.py:37-247) exec'd by oracle/gen_golden.py"""
    g = golden('twodes_time')
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    t1, t3, tw = g['t1'][None, :], g['t3'][:, None], float(g['t2'])
    assert np.array_equal(lo.td_ESA(E, gamma, dip, g_idx, e_idx, f_idx, t1, tw, t3), g['ESA'])
    assert np.array_equal(lo.td_GSB(E, gamma, dip, g_idx, e_idx, t1, tw, t3), g['GSB'])
    assert np.array_equal(lo.td_SE(E, gamma, dip, g_idx, e_idx, t1, tw, t3), g['SE'])


def test_liouvillian_eigen_solver_oracle_matches_reference():
    """tests/golden/super_lindblad.npz = outputs of lime.superoperator.Lindblad_solver itself (same host at
    generation time: bit-identical there; another host's LAPACK may round the eigenvectors differently)"""
    g = golden('super_lindblad')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=3, M=2, E=2, seed=61)
    o = lo.SuperLindblad(H, c_ops)
    o.eigenstates()
    A, B, C = g['A'], g['B'], g['C']
    assert relerr(o.evolve(rho0, g['tl'], e_ops), g['evolve']) <= 1e-9
    assert relerr(o.correlation_2op_1t(rho0, [A, B], g['tl']), g['c2_1t']) <= 1e-9
    assert relerr(o.correlation_3op_1w(rho0, [A, B, C], g['wl']), g['c3_1w']) <= 1e-9
    assert relerr(o.correlation_3op_2t(rho0, [A, B, C], g['tl'], g['taul']), g['c3_2t']) <= 1e-9


def test_sesolver_oracle_matches_reference():
    g = golden('sesolver')
    H = cases.rand_herm(5, 81)
    psi0 = cases.rand_cplx(5, 82)[:, 0]
    psi0 = psi0 / np.linalg.norm(psi0)
    e_ops = [cases.rand_herm(5, 83), cases.rand_herm(5, 84)]
    o, pl = lo.quantum_dynamics(H, psi0, dt=0.01, Nt=40, e_ops=e_ops, nout=2)
    assert np.array_equal(o, g['obs']) and np.array_equal(np.array(pl), g['psilist'])
    assert np.array_equal(lo.se_correlation_3op_2t(H, psi0, [g['A'], g['B'], g['C']], 0.01, 5, 6), g['c3_2t'])


def test_tpa_oracle_matches_reference():
    g = golden('sos_mol')
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    assert np.array_equal(np.array([lo.TPA(E, dip, w, g_idx, e_idx, f_idx, gamma) for w in g['wtpa']]), g['TPA'])
    assert np.array_equal(lo.photon_echo_core(E, dip, -g['wp'], g['wp'], 30.0, g_idx, e_idx, f_idx, gamma), g['PE'])


def test_fft_oracle_matches_reference():
    g = golden('fft')
    assert np.array_equal(lo.fft(g['f1'], g['xg'])[0], g['fft'])
    assert np.array_equal(lo.fft2(g['f2'], 0.1, 0.2)[2], g['fft2'])
    assert np.array_equal(lo.dft2(g['xs'], g['ys'], g['fxy'], g['kxs'], g['kys']), g['dft2'])
