"""
TEST INFRASTRUCTURE ONLY -- freeze outputs of the real reference into tests/golden/.

Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py

For every case the reference itself (imported through oracle/ref_shim.py) is run
on the seeded inputs of tests/cases.py, the NumPy restatement oracle/lime_oracle.py
is run on the same inputs, the two are compared (the table is printed and stored
in tests/golden/PINNING.json), and the REFERENCE output is written as a small
.npz fixture.  The reference's own golden files examples/cor.dat + examples/dm.dat
are converted to cavity_cor.npz verbatim (parsed, not recomputed).
"""
import io
import os
import sys
import json
import contextlib
import warnings

import numpy as np
from scipy.sparse import csr_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import ref_shim            # noqa: E402
import lime_oracle as lo   # noqa: E402
import cases               # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
report = {}


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    d = np.max(np.abs(a - b)) if a.size else 0.0
    s = np.max(np.abs(b)) if b.size else 1.0
    return float(d / s) if s > 0 else float(d)


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return f(*a, **k)


def note(name, **errs):
    report[name] = errs
    print('%-28s %s' % (name, '  '.join('%s=%.3g' % kv for kv in errs.items())))


def main():
    lime = ref_shim.load()
    import lime.oqs as oqs
    import lime.phys as phys
    import lime.heom.heom as heom
    import lime.signal.sos as sos
    import lime.superoperator as sop
    os.makedirs(GOLD, exist_ok=True)

    # ---- 1. the reference's own golden: examples/cor.dat, dm.dat ----------------
    ex = os.path.join(ref_shim.REFERENCE_ROOT, 'examples')
    cor = np.genfromtxt(os.path.join(ex, 'cor.dat'), dtype=complex)
    dm = np.genfromtxt(os.path.join(ex, 'dm.dat'), dtype=complex)
    t_gold = cor[:, 0].real
    cor_gold = cor[:, 1]
    dm_gold = dm[:, 1:].reshape(-1, 10, 10)
    H, rho0, ops, c_ops, tlist = cases.thermal_cavity()
    # replay with the reference's own rk4 + liouvillian (CSR operands)
    rho = ops[2].dot(rho0.dot(ops[0]))
    dt = tlist[1] - tlist[0]
    cr, dr = [], []
    for k in range(len(tlist)):
        rho = phys.rk4(rho, phys.liouvillian, dt, H, c_ops)
        cr.append(ops[1].dot(rho).diagonal().sum())
        dr.append(rho.toarray())
    t_o, cor_o, dm_o = lo.correlation_3p_1t(H, rho0, ops, c_ops, tlist)
    note('cavity_cor', ref_replay_vs_file=max(relerr(cr, cor_gold), relerr(dr, dm_gold)),
         oracle_vs_file=max(relerr(cor_o, cor_gold), relerr(dm_o, dm_gold), relerr(t_o, t_gold)))
    np.savez_compressed(os.path.join(GOLD, 'cavity_cor.npz'), t=t_gold, cor=cor_gold, dm=dm_gold)

    # ---- 2. config 1: examples/redfield.py ---------------------------------------
    H, a_ops, spectra, rho0, dt, Nt, e_ops, tlist = cases.redfield_example()
    solver = oqs.Redfield_solver(H, c_ops=a_ops, spectra=spectra)
    R, evecs = quiet(solver.redfield_tensor)
    res = quiet(solver.evolve, rho0, evecs=evecs, dt=dt, Nt=Nt, e_ops=e_ops)
    Ro, evo = lo.redfield_tensor(H, a_ops, spectra)
    obs_o, rl_o = lo.redfield(Ro, rho0, evecs=evo, Nt=Nt, dt=dt, e_ops=e_ops)
    t8 = tlist[:8]
    U_sos = quiet(solver.propagator, t8, 'SOS').copy()
    c4 = quiet(solver.correlation_4op_3t, rho0, [e_ops[0]] * 4, 'llll', t8)
    ex_ref = solver.expect(rho0, e_ops)
    U_eom = quiet(solver.propagator, t8, 'EOM').copy()
    note('redfield_example', R=relerr(Ro.toarray(), R.toarray()), evecs=relerr(evo, evecs),
         obs=relerr(obs_o, res.observables), rholist=relerr(rl_o, res.rholist),
         U_sos=relerr(lo.redfield_propagator(Ro, t8, 'SOS'), U_sos),
         U_eom=relerr(lo.redfield_propagator(Ro, t8, 'EOM'), U_eom),
         expect=relerr(lo.redfield_expect(U_sos, evecs, rho0, e_ops), ex_ref),
         corr4=relerr(lo.redfield_correlation_4op_3t(-1j * U_sos, 2, rho0, [e_ops[0]] * 4, 'llll'), c4))
    np.savez_compressed(os.path.join(GOLD, 'redfield_example.npz'), R=R.toarray(), evecs=evecs,
                        observables=res.observables, rholist=np.array(res.rholist),
                        U_sos=U_sos, U_eom=U_eom, corr4=c4, expect=ex_ref)

    # ---- 2b. multi-level Redfield with two baths -----------------------------------
    H, a_ops, spectra, rho0 = cases.redfield_multilevel()
    solver = oqs.Redfield_solver(H, c_ops=a_ops, spectra=spectra)
    R, evecs = quiet(solver.redfield_tensor)
    e_ops = [a_ops[0], cases.rand_herm(5, 99)]
    res = quiet(solver.evolve, rho0, dt=0.02, Nt=60, e_ops=e_ops)
    Ro, evo = lo.redfield_tensor(H, a_ops, spectra)
    obs_o, rl_o = lo.redfield(Ro, rho0, evecs=evo, Nt=60, dt=0.02, e_ops=e_ops)
    note('redfield_multilevel', R=relerr(Ro.toarray(), R.toarray()),
         obs=relerr(obs_o, res.observables), rholist=relerr(rl_o, res.rholist))
    np.savez_compressed(os.path.join(GOLD, 'redfield_multilevel.npz'), R=R.toarray(), evecs=evecs,
                        observables=res.observables, rholist=np.array(res.rholist), e1=e_ops[1])

    # ---- 3. Lindblad dense -------------------------------------------------------
    H, c_ops, e_ops, rho0 = cases.lindblad_dense()
    res = oqs._lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=80, dt=0.01)
    obs_o, rl_o = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=80, dt=0.01)
    rhs_ref = oqs.liouvillian(rho0, H, c_ops)
    Lsup = sop.liouvillian(H, c_ops)
    note('lindblad_dense', obs=relerr(obs_o, res.observables), rholist=relerr(rl_o, res.rholist),
         rhs=relerr(lo.liouvillian(rho0, H, c_ops), rhs_ref),
         superop=relerr(lo.liouvillian_super(H, c_ops).toarray(), Lsup.toarray()),
         superop_vs_rhs=relerr(Lsup.dot(rho0.flatten()), rhs_ref.flatten()))
    np.savez_compressed(os.path.join(GOLD, 'lindblad_dense.npz'), observables=res.observables,
                        rholist=np.array(res.rholist), rhs=rhs_ref, superop=Lsup.toarray())

    # ---- 3b. Lindblad Jaynes-Cummings, CSR operands (config-2 shape, small cutoff) --
    H, c_ops, e_ops, rho0 = cases.jc_point(ncav=8)
    res = oqs._lindblad(csr_matrix(H), csr_matrix(rho0), [csr_matrix(c) for c in c_ops],
                        e_ops=[csr_matrix(e) for e in e_ops], Nt=100, dt=0.01)
    obs_o, rl_o = lo.lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=100, dt=0.01)
    Ho, co, eo = lo.jaynes_cummings(1.0, 1.05, 0.1, 8, 0.05)
    rl_ref = np.array([r.toarray() for r in res.rholist])
    note('lindblad_jc', obs=relerr(obs_o, res.observables), rholist=relerr(rl_o, rl_ref),
         builder=max(relerr(Ho.toarray(), H), relerr(co[0].toarray(), c_ops[0])))
    np.savez_compressed(os.path.join(GOLD, 'lindblad_jc.npz'), observables=res.observables,
                        rho_final=rl_ref[-1], rho_mid=rl_ref[49])

    # ---- 3c. driven Lindblad ---------------------------------------------------------
    H0, c_ops, e_ops, rho0 = cases.lindblad_dense(n=4, M=1, E=1, seed=77)
    H1 = cases.rand_herm(4, 78)

    def f1(t):
        return 0.3 * np.exp(-(t - 0.4) ** 2 / 0.02) * np.exp(-1j * 2.0 * t)
    Hlist = [H0.copy(), [H1, f1]]
    res = oqs._lindblad_driven(Hlist, rho0, c_ops=c_ops, e_ops=e_ops, Nt=60, dt=0.01, t0=0.1)
    obs_o, rl_o = lo.lindblad_driven([H0.copy(), [H1, f1]], rho0, c_ops, e_ops, Nt=60, dt=0.01,
                                     t0=0.1, strict_parity=True)
    obs_f, rl_f = lo.lindblad_driven([H0.copy(), [H1, f1]], rho0, c_ops, e_ops, Nt=60, dt=0.01, t0=0.1)
    note('lindblad_driven', obs_strict=relerr(obs_o, res.observables),
         rholist_strict=relerr(rl_o, res.rholist), fixed_vs_strict=relerr(obs_f, obs_o))
    np.savez_compressed(os.path.join(GOLD, 'lindblad_driven.npz'), observables_strict=res.observables,
                        rho_final_strict=res.rholist[-1], H1=H1)

    # ---- 3d. Lindblad correlation functions -------------------------------------------
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=4, M=2, E=1, seed=31)
    ops3 = [cases.rand_cplx(4, 32, 0.5), cases.rand_herm(4, 33), cases.rand_cplx(4, 34, 0.5)]
    solver = oqs.Lindblad_solver(H, c_ops=c_ops)
    c1 = solver.correlation_3op_1t(rho0, ops3, dt=0.02, Nt=30)
    c2 = solver.correlation_3op_2t(rho0, ops3, dt=0.02, Nt=6, Ntau=7)
    note('lindblad_corr', c3op1t=relerr(lo.lindblad_correlation_3op_1t(H, c_ops, rho0, ops3, 0.02, 30), c1),
         c3op2t=relerr(lo.lindblad_correlation_3op_2t(H, c_ops, rho0, ops3, 0.02, 6, 7), c2))
    np.savez_compressed(os.path.join(GOLD, 'lindblad_corr.npz'), c3op1t=c1, c3op2t=c2,
                        A=ops3[0], B=ops3[1], C=ops3[2])

    # ---- 4. HEOM ---------------------------------------------------------------------------
    H, sz, rho0 = cases.spin_boson_heom()[:3]
    fn = '/tmp/_heom_dl_golden.dat'
    out = quiet(oqs._heom_dl, H, rho0, sz, None, 300.0, 0.002, 0.0005, 12, 0.05, 200, fn)
    traj = np.genfromtxt(fn, dtype=complex)[:, 1:].reshape(-1, 2, 2)
    ado_o, traj_o = lo.heom_dl(H, rho0, sz, 300.0, 0.002, 0.0005, 12, 0.05, 200)
    note('heom_dl', final=relerr(ado_o[:, :, 0], out), traj=relerr(traj_o, traj))
    np.savez_compressed(os.path.join(GOLD, 'heom_dl.npz'), rho_final=out, traj=traj)

    tables = {}
    worst = 0
    for dims, exc in [([3, 3], 2), ([5, 5], 4), ([13, 13], 12), ([4, 3, 2], 3), ([3] * 4, 0),
                      ([5] * 6, 4), ([3] * 8, 2), ([5] * 14, 4)]:
        n, s2i, i2s = heom.enr_state_dictionaries(dims, exc)
        arr = np.array([np.asarray(i2s[i]) for i in range(n)], dtype=np.int64).reshape(n, len(dims))
        assert all(s2i[tuple(arr[i])] == i for i in range(n))
        no, s2io, i2so = lo.enr_state_dictionaries(dims, exc)
        arro = np.array([np.asarray(i2so[i]) for i in range(no)], dtype=np.int64).reshape(no, len(dims))
        worst = max(worst, 0 if (n == no and np.array_equal(arr, arro)) else 1)
        if exc:
            worst = max(worst, 0 if np.array_equal(lo.enr_states_fast(dims, exc), arr) else 1)
        tables['d' + '_'.join(map(str, dims)) + '_x' + str(exc)] = arr.astype(np.int16)
    # excitations=None: state_number_enumerate yields ndarrays (lime/heom/heom.py:66),
    # which enr_state_dictionaries cannot hash (:104) -> TypeError in the reference
    try:
        heom.enr_state_dictionaries([2, 2], None)
        none_raises = 0
    except TypeError:
        none_raises = 1
    note('heom_tables', none_raises_TypeError=none_raises, mismatches=worst, fmo_nhe=int(tables['d' + '_'.join(['5'] * 14) + '_x4'].shape[0]))
    np.savez_compressed(os.path.join(GOLD, 'heom_tables.npz'), **tables)

    mats = {}
    worst = 0.0
    for i, (K, lam, gam, T) in enumerate([(2, 0.2, 1.0, 1.0), (4, 35 / 219474.6305, 1 / (50 / 2.41888432651e-2), 300 / 315775.13),
                                          (1, 0.1, 0.5, 2.0), (3, 1.0, 2.0, 0.25)]):
        c, nu = heom._calc_matsubara_params(K, lam, gam, T)
        co, nuo = lo.calc_matsubara_params(K, lam, gam, T)
        worst = max(worst, relerr(co, c), relerr(nuo, nu))
        mats['c%d' % i] = np.array(c, dtype=complex)
        mats['nu%d' % i] = np.array(nu, dtype=float)
        mats['par%d' % i] = np.array([K, lam, gam, T])
    note('heom_matsubara', err=worst)
    np.savez_compressed(os.path.join(GOLD, 'heom_matsubara.npz'), **mats)

    # ---- 5. sum-over-states response functions ------------------------------------------------
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    au2ev = 27.211386
    w1 = np.linspace(1.4, 2.1, 12) / au2ev
    w3 = np.linspace(1.3, 2.2, 12) / au2ev
    t2 = 30.0 / 2.41888432651e-2
    g = {}
    g['GSB'] = sos.GSB(E, dip, w1, w3, t2, g_idx, e_idx, gamma)
    g['SE'] = sos.SE(E, dip, w1, w3, t2, g_idx, e_idx, gamma)
    g['ESA'] = sos.ESA(E, dip, w1, w3, t2, g_idx, e_idx, f_idx, gamma)
    g['PE'] = sos._photon_echo(E, dip, -w1, w3, t2, g_idx, e_idx, f_idx, gamma)
    g['SE_t3'] = sos._SE(E, dip, -w1, w3, t2, g_idx, e_idx, gamma, dephasing=0.01 / au2ev)
    g['ESA_t3'] = sos._ESA(E, dip, -w1, w3, t2, g_idx, e_idx, f_idx, gamma, dephasing=0.01 / au2ev)
    w2 = np.linspace(3.0, 4.0, 10) / au2ev
    w1b = np.linspace(1.4, 2.1, 10) / au2ev
    g['R1_t3'] = sos.DQC_R1(E, dip, omega1=w1b, omega2=w2, tau3=1e-6, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    g['R2_t3'] = sos.DQC_R2(E, dip, omega1=w1b, omega2=w2, tau3=1e-6, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    g['R1_t1'] = sos.DQC_R1(E, dip, omega2=w2, omega3=w1b, tau1=50.0, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    g['R2_t1'] = sos.DQC_R2(E, dip, omega2=w2, omega3=w1b, tau1=50.0, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    g['TPA2D'] = sos.TPA2D(E, dip, w2, w1b, g_idx, e_idx, f_idx, gamma)
    g['TPA2D_to'] = sos.TPA2D_time_order(E, dip, w2, w1b, g_idx, e_idx, f_idx, gamma)
    o = {}
    o['GSB'] = lo.GSB(E, dip, w1, w3, t2, g_idx, e_idx, gamma)
    o['SE'] = lo.SE(E, dip, w1, w3, t2, g_idx, e_idx, gamma)
    o['ESA'] = lo.ESA(E, dip, w1, w3, t2, g_idx, e_idx, f_idx, gamma)
    o['PE'] = lo.photon_echo_core(E, dip, -w1, w3, t2, g_idx, e_idx, f_idx, gamma)
    o['SE_t3'] = lo.SE_t3(E, dip, -w1, w3, t2, g_idx, e_idx, gamma, dephasing=0.01 / au2ev)
    o['ESA_t3'] = lo.ESA_t3(E, dip, -w1, w3, t2, g_idx, e_idx, f_idx, gamma, dephasing=0.01 / au2ev)
    o['R1_t3'] = lo.DQC_R1(E, dip, omega1=w1b, omega2=w2, tau3=1e-6, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    o['R2_t3'] = lo.DQC_R2(E, dip, omega1=w1b, omega2=w2, tau3=1e-6, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    o['R1_t1'] = lo.DQC_R1(E, dip, omega2=w2, omega3=w1b, tau1=50.0, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    o['R2_t1'] = lo.DQC_R2(E, dip, omega2=w2, omega3=w1b, tau1=50.0, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx, gamma=gamma)
    o['TPA2D'] = lo.TPA2D(E, dip, w2, w1b, g_idx, e_idx, f_idx, gamma)
    o['TPA2D_to'] = lo.TPA2D_time_order(E, dip, w2, w1b, g_idx, e_idx, f_idx, gamma)
    note('sos', **{k: relerr(o[k], g[k]) for k in g})
    np.savez_compressed(os.path.join(GOLD, 'sos.npz'), w1=w1, w3=w3, w2=w2, w1b=w1b, t2=t2, **g)

    # ---- Mol-based wrappers and the single-frequency TPA, lime/signal/sos.py:199-228, 731-902
    import tempfile
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    mol = cases.DuckMol(E, dip, gamma, dephasing=0.01 / au2ev)
    tmpd = tempfile.mkdtemp()
    wp = np.linspace(1.4, 2.1, 9) / au2ev
    g = {'PE': quiet(sos.photon_echo, mol, wp, wp, t2=30.0, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx,
                     fname=os.path.join(tmpd, 's')),
         'PE_t3': quiet(sos.photon_echo_t3, mol, wp, wp, 20.0, g_idx=g_idx, e_idx=e_idx, f_idx=f_idx,
                        fname=os.path.join(tmpd, 't')),
         'TPA': np.array([sos.TPA(E, dip, w, g_idx, e_idx, f_idx, gamma) for w in np.linspace(3.0, 4.0, 7) / au2ev])}
    o = {'PE': lo.photon_echo_core(E, dip, -wp, wp, 30.0, g_idx, e_idx, f_idx, gamma),
         'PE_t3': lo.SE_t3(E, dip, -wp, wp, 20.0, g_idx, e_idx, gamma, dephasing=0.01 / au2ev) +
         lo.ESA_t3(E, dip, -wp, wp, 20.0, g_idx, e_idx, f_idx, gamma, dephasing=0.01 / au2ev),
         'TPA': np.array([lo.TPA(E, dip, w, g_idx, e_idx, f_idx, gamma) for w in np.linspace(3.0, 4.0, 7) / au2ev])}
    note('sos_mol', **{k: relerr(o[k], g[k]) for k in g})
    np.savez_compressed(os.path.join(GOLD, 'sos_mol.npz'), wp=wp, wtpa=np.linspace(3.0, 4.0, 7) / au2ev, **g)

    # ---- Liouvillian eigen-decomposition solver, lime/superoperator.py:456-773
    import lime.superoperator as lsup
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=3, M=2, E=2, seed=61)
    ref = lsup.Lindblad_solver(H, c_ops)
    ref.eigenstates()
    osv = lo.SuperLindblad(H, c_ops)
    osv.eigenstates()
    tl = np.linspace(0, 2.0, 7)
    taul = np.linspace(0, 1.5, 5)
    wl = np.linspace(-2, 2, 6)
    A3, B3, C3 = cases.rand_cplx(3, 71), cases.rand_cplx(3, 72), cases.rand_cplx(3, 73)
    g = {'evolve': ref.evolve(rho0, tl, e_ops).observables,
         'c2_1t': ref.correlation_2op_1t(rho0, [A3, B3], tl), 'c2_1w': ref.correlation_2op_1w(rho0, [A3, B3], wl),
         'c3_1t': ref.correlation_3op_1t(rho0, [A3, B3, C3], tl), 'c3_1w': ref.correlation_3op_1w(rho0, [A3, B3, C3], wl),
         'c3_2t': ref.correlation_3op_2t(rho0, [A3, B3, C3], tl, taul),
         'c4_2t': ref.correlation_4op_2t(rho0, [A3, B3, C3, A3], tl, taul)}
    o = {'evolve': osv.evolve(rho0, tl, e_ops),
         'c2_1t': osv.correlation_2op_1t(rho0, [A3, B3], tl), 'c2_1w': osv.correlation_2op_1w(rho0, [A3, B3], wl),
         'c3_1t': osv.correlation_3op_1t(rho0, [A3, B3, C3], tl), 'c3_1w': osv.correlation_3op_1w(rho0, [A3, B3, C3], wl),
         'c3_2t': osv.correlation_3op_2t(rho0, [A3, B3, C3], tl, taul),
         'c4_2t': osv.correlation_4op_2t(rho0, [A3, B3, C3, A3], tl, taul)}
    note('super_lindblad', **{k: relerr(o[k], g[k]) for k in g})
    np.savez_compressed(os.path.join(GOLD, 'super_lindblad.npz'), tl=tl, taul=taul, wl=wl, A=A3, B=B3, C=C3, **g)

    # ---- spectral post-processing, lime/fft.py
    import lime.fft as lfft
    rng = np.random.default_rng(91)
    xg = np.linspace(-3.0, 4.0, 48)
    f1 = rng.standard_normal((5, 48)) + 1j * rng.standard_normal((5, 48))
    f2 = rng.standard_normal((24, 24)) + 1j * rng.standard_normal((24, 24))
    xs, ys = np.linspace(0, 2, 11), np.linspace(-1, 1, 9)
    fxy = rng.standard_normal((9, 11)) + 1j * rng.standard_normal((9, 11))
    kxs, kys = np.linspace(-3, 3, 7), np.linspace(-2, 2, 5)
    # lime's dft opens a figure (lime/fft.py:122-123), which the plotting stub cannot unpack: exec its source
    # without the two plotting lines
    import inspect
    dsrc = [l for l in inspect.getsource(lfft.dft).split('\n') if 'plt.' not in l and 'ax.' not in l]
    dns = {'np': np}
    exec('\n'.join(dsrc), dns)
    ref_dft = dns['dft']
    g = {'fft': lfft.fft(f1, xg)[0], 'fft_freq': lfft.fft(f1, xg)[1],
         'ifft': lfft.ifft(f1[0], xg)[0], 'ifft_freq': lfft.ifft(f1[0], xg)[1],
         'fft2': lfft.fft2(f2, 0.1, 0.2)[2], 'fft2_fy': lfft.fft2(f2, 0.1, 0.2)[1],
         'dft': ref_dft(xg, f1[1], kxs), 'dft2': lfft.dft2(xs, ys, fxy, kxs, kys)}
    o = {'fft': lo.fft(f1, xg)[0], 'fft_freq': lo.fft(f1, xg)[1],
         'ifft': lo.ifft(f1[0], xg)[0], 'ifft_freq': lo.ifft(f1[0], xg)[1],
         'fft2': lo.fft2(f2, 0.1, 0.2)[2], 'fft2_fy': lo.fft2(f2, 0.1, 0.2)[1],
         'dft': lo.dft(xg, f1[1], kxs), 'dft2': lo.dft2(xs, ys, fxy, kxs, kys)}
    note('fft', **{k: relerr(o[k], g[k]) for k in g})
    np.savez_compressed(os.path.join(GOLD, 'fft.npz'), xg=xg, f1=f1, f2=f2, xs=xs, ys=ys, fxy=fxy, kxs=kxs, kys=kys, **g)

    # ---- wave-function solver, lime/mol.py:1094-1391
    import lime.mol as lmol
    H = cases.rand_herm(5, 81)
    psi0 = cases.rand_cplx(5, 82)[:, 0]
    psi0 = psi0 / np.linalg.norm(psi0)
    e_ops = [cases.rand_herm(5, 83), cases.rand_herm(5, 84)]
    ops = [cases.rand_cplx(5, 85), cases.rand_cplx(5, 86), cases.rand_cplx(5, 87)]
    ses = lmol.SESolver(H)
    r = quiet(ses.run, psi0=psi0, dt=0.01, Nt=40, e_ops=e_ops, nout=2)
    g = {'obs': r.observables, 'psilist': np.array(r.psilist),
         'c3_1t': quiet(ses.correlation_3op_1t, psi0, ops, 0.01, 12),
         'c3_2t': quiet(ses.correlation_3op_2t, psi0, ops, 0.01, 5, 6),
         'c4_2t': quiet(ses.correlation_4op_2t, psi0, ops + [ops[0]], 0.01, 4, 3)}
    oo, pl = lo.quantum_dynamics(H, psi0, dt=0.01, Nt=40, e_ops=e_ops, nout=2)
    o = {'obs': oo, 'psilist': np.array(pl), 'c3_1t': lo.se_correlation_3op_1t(H, psi0, ops, 0.01, 12),
         'c3_2t': lo.se_correlation_3op_2t(H, psi0, ops, 0.01, 5, 6),
         'c4_2t': lo.se_correlation_3op_2t(H, psi0, [ops[0], ops[1] @ ops[2], ops[0]], 0.01, 4, 3)}
    note('sesolver', **{k: relerr(o[k], g[k]) for k in g})
    np.savez_compressed(os.path.join(GOLD, 'sesolver.npz'), A=ops[0], B=ops[1], C=ops[2], **g)



    # ---- _lindblad_driven with CSR operands (and a CSR rho0, the only way lime runs them), real drive envelope
    Hj, cj, ej, rj = cases.jc_point(ncav=6)
    H1 = np.kron(np.array([[0, 1.], [1, 0]]), np.identity(6))
    fdr = lambda t: 0.2 * np.exp(-(t - 0.2) ** 2 / 0.02) * np.cos(3 * t)
    rdr = quiet(oqs._lindblad_driven, [csr_matrix(Hj), [csr_matrix(H1), fdr]], csr_matrix(rj), c_ops=[csr_matrix(c) for c in cj],
                e_ops=[csr_matrix(e) for e in ej], Nt=40, dt=0.01, t0=0.05)
    odr, rldr = lo.lindblad_driven([Hj.copy(), [H1, fdr]], rj, cj, ej, Nt=40, dt=0.01, t0=0.05)
    note('lindblad_driven_csr', obs=relerr(odr, rdr.observables), rho=relerr(rldr[-1], rdr.rholist[-1].toarray()))
    np.savez_compressed(os.path.join(GOLD, 'lindblad_driven_csr.npz'), H1=H1, obs=rdr.observables,
                        rho_final=rdr.rholist[-1].toarray(), rho_mid=rdr.rholist[19].toarray())

    # ---- laser-driven wave-function dynamics, SESolver.run(pulse=...) -> driven_dynamics, lime/mol.py:1094-1171,1473-1560
    class _Pulse:                       # the two attributes SESolver.run uses of lime.optics.Pulse: .efield(t)
        def __init__(self, a, w, tc, sig):
            self.a, self.w, self.tc, self.sig = a, w, tc, sig

        def efield(self, t):
            return self.a * np.exp(-(t - self.tc) ** 2 / 2. / self.sig ** 2) * np.exp(-1j * self.w * (t - self.tc))
    Hd = cases.rand_herm(5, 181)
    mu = cases.rand_herm(5, 182)
    mu2 = cases.rand_herm(5, 183)
    psi0 = cases.rand_cplx(5, 184)[:, 0]
    psi0 = psi0 / np.linalg.norm(psi0)
    e_sp = [csr_matrix(cases.rand_herm(5, 185)), csr_matrix(cases.rand_herm(5, 186))]
    p1, p2 = _Pulse(0.3, 1.1, 0.15, 0.1), _Pulse(0.2, 0.4, 0.25, 0.2)
    rs = quiet(lmol.SESolver(csr_matrix(Hd)).run, psi0=psi0, dt=0.01, Nt=40, e_ops=e_sp, nout=2, edip=csr_matrix(mu), pulse=p1)
    rl = quiet(lmol.SESolver(csr_matrix(Hd)).run, psi0=psi0, dt=0.01, Nt=30, e_ops=e_sp, nout=1,
               edip=[csr_matrix(mu), csr_matrix(mu2)], pulse=[p1, p2])
    dense_fails = 'ok'
    try:                                    # an ndarray Hamiltonian cannot be combined with lime's sparse psi: ValueError
        quiet(lmol.SESolver(Hd.copy()).run, psi0=psi0, dt=0.01, Nt=20, e_ops=e_sp, nout=1, edip=mu, pulse=p1)
    except Exception as exc:
        dense_fails = type(exc).__name__
    tovec = lambda pl: np.array([np.asarray(p.toarray() if hasattr(p, 'toarray') else p).reshape(-1) for p in pl])
    g = {'obs1': rs.observables, 'psi1': tovec(rs.psilist), 'obs2': rl.observables, 'psi2': tovec(rl.psilist)}
    e_d = [e.toarray() for e in e_sp]
    o1, pl1 = lo.driven_dynamics([Hd, [mu, p1.efield]], psi0, dt=0.01, Nt=40, e_ops=e_d, nout=2)
    o2, pl2 = lo.driven_dynamics([Hd, [mu, p1.efield], [mu2, p2.efield]], psi0, dt=0.01, Nt=30, e_ops=e_d, nout=1)
    o = {'obs1': o1, 'psi1': np.array(pl1), 'obs2': o2, 'psi2': np.array(pl2)}
    note('sesolver_driven', dense_H_raises_ValueError=float(dense_fails != 'ValueError'), **{k: relerr(o[k], g[k]) for k in g})
    np.savez_compressed(os.path.join(GOLD, 'sesolver_driven.npz'), Hd=Hd, mu=mu, mu2=mu2, psi0=psi0, e0=e_d[0], e1=e_d[1],
                        pulses=np.array([[0.3, 1.1, 0.15, 0.1], [0.2, 0.4, 0.25, 0.2]]), **g)

    # ---- time-domain response functions, lime/signal/2DES.py:37-247 (the module cannot be imported:
    # it runs undefined names at :249-263; the function definitions themselves are exec'd verbatim here)
    src = open('/root/reference/lime/signal/2DES.py').read().split('\n')
    ns = {'np': np, 'au2mev': 27211.386, 'au2ev': 27.211386}       # lime/units.py:6,8 (imported at 2DES.py:24)
    exec('\n'.join(src[36:247]), ns)
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    ns['en'] = E
    ns['decay'] = gamma
    t1 = np.linspace(0, 300, 9)[None, :]
    t3 = np.linspace(0, 400, 7)[:, None]
    tw = 50.0
    g = {'ESA': ns['ESA'](E, dip, g_idx, e_idx, f_idx, gamma, t1, tw, t3),
         'GSB': ns['GSB'](E, dip, g_idx, e_idx, gamma, t1, tw, t3),
         'SE': ns['SE'](E, dip, g_idx, e_idx, t1, tw, t3)}
    o = {'ESA': lo.td_ESA(E, gamma, dip, g_idx, e_idx, f_idx, t1, tw, t3),
         'GSB': lo.td_GSB(E, gamma, dip, g_idx, e_idx, t1, tw, t3),
         'SE': lo.td_SE(E, gamma, dip, g_idx, e_idx, t1, tw, t3)}
    note('twodes_time', **{k: relerr(o[k], g[k]) for k in g})
    np.savez_compressed(os.path.join(GOLD, 'twodes_time.npz'), t1=t1[0], t3=t3[:, 0], t2=tw, **g)


    # ---- multi-index HEOM coupling rules, lime/heom/heom.py:156-216.  The excerpt sits in a dead __main__ block that
    # uses names lime never defines (cut from QuTiP 4.5's HSolverDL._configure); the loop itself (:156-216) is exec'd
    # VERBATIM here with those names supplied: cy_pad_csr places a superoperator block at (row ADO, column ADO) of the
    # hierarchy Liouvillian, spreQ / spostQ / commQ are the left / right / commutator superoperators of Q (row-major
    # vec, lime/superoperator.py:131-151 convention), mol.idm the identity superoperator, renorm False.  What comes out
    # is the REFERENCE's own matrix L_helems; the oracle's heom_rhs (without the system term) must equal L_helems . vec.
    hsrc = open('/root/reference/lime/heom/heom.py').read().split('\n')
    loop = '\n'.join(l[4:] if l.startswith('    ') else l for l in hsrc[155:216])     # lines 156-216, un-indented once
    assert loop.startswith('for he_idx in range(N_he):') and 'L_he = cy_pad_csr(op, N_he, N_he, he_idx, he_idx_neigh)' in loop

    def heom_rules_reference(n, N_m, N_c, Qs, qmap, cnu, lam, gam, T):
        """run lime/heom/heom.py:156-216; Qs/qmap: coupling operator of mode k (lime: one Q); cnu: the (c, nu) lists
        `_calc_matsubara_params` returns inside the loop (lime's own function unless a multi-bath list is supplied)"""
        import copy as _copy
        N_he, he2idx, idx2he = heom.enr_state_dictionaries([N_c + 1] * N_m, N_c)
        nn = n * n
        I = np.eye(n)

        def pad(op, nr, nc, i, j):
            out = np.zeros((nr * nn, nc * nn), dtype=complex)
            out[i * nn:(i + 1) * nn, j * nn:(j + 1) * nn] = op
            return out

        class Names(dict):
            # spreQ / spostQ / commQ are looked up inside `for k in range(N_m)`: resolve them for the CURRENT mode k,
            # which is how a bath-dependent coupling operator enters rules that are written for a single Q
            def __missing__(self, key):
                q = Qs[qmap[self['k']]] if 'k' in self else Qs[0]
                if key == 'spreQ':
                    return np.kron(q, I)
                if key == 'spostQ':
                    return np.kron(I, q.T)
                if key == 'commQ':
                    return np.kron(q, I) - np.kron(I, q.T)
                raise KeyError(key)

        class _Mol:
            idm = np.eye(nn)
        ns = Names(np=np, copy=_copy.copy, N_he=N_he, he2idx=he2idx, idx2he=idx2he, N_m=N_m, N_c=N_c,
                   coup_strength=lam, cut_freq=gam, temperature=T, mol=_Mol(), cy_pad_csr=pad, renorm=False,
                   L_helems=np.zeros((N_he * nn, N_he * nn), dtype=complex), N_he_interact=0,
                   _calc_matsubara_params=(heom._calc_matsubara_params if cnu is None else (lambda *a: cnu)))
        quiet(exec, loop, {'__builtins__': __builtins__}, ns)
        return ns['L_helems'], ns['N_he_interact'], N_he

    rng = np.random.default_rng(11)

    def rherm(n):
        a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        return a + a.conj().T

    hr_cases = {}
    worst = 0.0
    # (name, n, N_m, N_c, coupling operators, mode -> operator, (c, nu) override)
    lam, gam, T = 0.3, 0.7, 1.3
    c2a, nu2a = heom._calc_matsubara_params(2, lam, gam, T)
    c2b, nu2b = heom._calc_matsubara_params(2, 0.5 * lam, 1.4 * gam, T)
    Q2 = rherm(2)
    Q3a, Q3b = rherm(3), rherm(3)
    for name, n, N_m, N_c, Qs, qmap, cnu in [
            ('k2_d3_n2', 2, 2, 3, [Q2], [0, 0], None),                                  # lime's case: one bath, one Q
            ('k4_d2_n3', 3, 4, 2, [Q3a], [0, 0, 0, 0], None),
            ('b2k2_d3_n3', 3, 4, 3, [Q3a, Q3b], [0, 0, 1, 1], (c2a + c2b, nu2a + nu2b)),  # two baths, bath-major modes
            ('b2k2_d2_proj', 3, 4, 2, [np.diag([1.0, 0, 0]).astype(complex), np.diag([0, 1.0, 0]).astype(complex)],
             [0, 0, 1, 1], (c2a + c2b, nu2a + nu2b))]:                                   # projector couplings (FMO-like)
        L, ncoup, N_he = heom_rules_reference(n, N_m, N_c, Qs, qmap, cnu, lam, gam, T)
        c, nu = (heom._calc_matsubara_params(N_m, lam, gam, T) if cnu is None else cnu)
        st, dn, up = lo.heom_tables([N_c + 1] * N_m, N_c)
        assert st.shape[0] == N_he
        ado = rng.standard_normal((N_he, n, n)) + 1j * rng.standard_normal((N_he, n, n))
        rhs_o = lo.heom_rhs(ado, np.zeros((n, n), dtype=complex), np.stack(Qs), qmap, c, nu, st, dn, up)
        rhs_r = (L @ ado.reshape(-1)).reshape(N_he, n, n)
        e = relerr(rhs_o, rhs_r)
        worst = max(worst, e)
        assert ncoup == int((dn >= 0).sum() + (up >= 0).sum())
        hr_cases[name + '_L'] = L
        hr_cases[name + '_Q'] = np.stack(Qs)
        hr_cases[name + '_meta'] = np.array([n, N_m, N_c, ncoup])
        hr_cases[name + '_qmap'] = np.array(qmap)
        hr_cases[name + '_c'] = np.array(c, dtype=complex)
        hr_cases[name + '_nu'] = np.array(nu, dtype=float)
    note('heom_rules', oracle_rhs_vs_reference_L=worst, cases=len(hr_cases) // 6)
    np.savez_compressed(os.path.join(GOLD, 'heom_rules.npz'), **hr_cases)


    # ---- round-2 API rows: _correlation_2p_1t (lime/oqs.py:726-800), getG / Redfield_solver.gf (:145-167, 474-526)
    import tempfile
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=6)
    A, B = e_ops[0], cases.rand_cplx(6, 77, 0.5)
    fn = os.path.join(tempfile.mkdtemp(), 'cor.dat')
    cor_ref = quiet(oqs.Lindblad_solver(H, c_ops).correlation_2op_1t, rho0, A, B, 0.01, 30, output=fn)
    cor_txt = open(fn).read()
    fn2 = os.path.join(tempfile.mkdtemp(), 'cor.dat')
    cor_o = lo.correlation_2p_1t(H, rho0, [A, B], c_ops, 0.01, 30, output=fn2)
    Hj, cj, ej, rj = cases.jc_point(ncav=8)
    corj_ref = quiet(oqs._correlation_2p_1t, csr_matrix(Hj), rj, [ej[0], cj[0] / np.sqrt(0.05)], [csr_matrix(c) for c in cj], 0.01, 25,
                     output=os.path.join(tempfile.mkdtemp(), 'cor.dat'))
    corj_o = lo.correlation_2p_1t(Hj, rj, [ej[0], cj[0] / np.sqrt(0.05)], cj, 0.01, 25)
    Hr, a_ops, spectra, rr = cases.redfield_multilevel()
    sol = oqs.Redfield_solver(Hr, c_ops=a_ops, spectra=spectra)
    quiet(sol.redfield_tensor)
    tg = np.linspace(0, 3.0, 7)
    D = Hr.shape[0] ** 2
    wg = np.linspace(-2.0, 2.0, D)                      # lime's frequency-domain einsum needs len(w) == dim(L)
    g = {'cor2': cor_ref, 'cor2_jc': corj_ref,
         'G_time': quiet(sol.gf, tg, method='diag'),
         'G_freq': quiet(oqs.getG, 1j * sol.R, tg, w=wg, domain='freq')}
    o = {'cor2': cor_o, 'cor2_jc': corj_o, 'G_time': lo.getG(1j * sol.R, tg), 'G_freq': lo.getG(1j * sol.R, tg, w=wg, domain='freq')}
    gf_eom = 'ok'
    try:
        quiet(oqs.Redfield_solver(Hr, c_ops=a_ops, spectra=spectra).gf, tg, method='EOM')
    except Exception as exc:                            # lime multiplies a Python list by -1j: always TypeError
        gf_eom = type(exc).__name__
    note('api_r2', cor_file_identical=float(cor_txt != open(fn2).read()), gf_eom_raises_TypeError=float(gf_eom != 'TypeError'),
         **{k: relerr(o[k], g[k]) for k in g})
    np.savez_compressed(os.path.join(GOLD, 'api_r2.npz'), A=A, B=B, tg=tg, wg=wg, cor_txt=np.array(cor_txt), **g)


    # ---- ETPA double time integrals, sos._etpa (lime/signal/sos.py:1171-1223), SURVEY 8f item 4
    E, dip, gamma, g_idx, e_idx, f_idx = cases.sos_system()
    rg = np.random.default_rng(21)
    nt = 24
    t1e = np.linspace(-40.0, 60.0, nt)
    t2e = np.linspace(-40.0, 60.0, nt)
    jta = (rg.standard_normal((nt, nt)) + 1j * rg.standard_normal((nt, nt))) * np.exp(-(t1e[None, :] ** 2 + t2e[:, None] ** 2) / 900.0)
    wps = np.linspace(0.12, 0.16, 5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        et_ref = quiet(sos._etpa, wps, E, dip, jta, t1e, t2e, list(g_idx), list(e_idx), list(f_idx))
        et_o = lo.etpa_core(wps, E, dip, jta, t1e, t2e, list(g_idx), list(e_idx), list(f_idx))
    note('etpa', etpa=relerr(et_o, et_ref))
    np.savez_compressed(os.path.join(GOLD, 'etpa.npz'), wps=wps, t1=t1e, t2=t2e, jta=jta, etpa=et_ref)

    with open(os.path.join(GOLD, 'PINNING.json'), 'w') as f:
        json.dump({'generated_by': 'oracle/gen_golden.py', 'reference': 'binggu56/lime @ /root/reference',
                   'numpy': np.__version__, 'oracle_vs_reference_max_rel_err': report}, f, indent=1)
    print('wrote', GOLD)


if __name__ == '__main__':
    main()
