"""CPU: host logic of the product package -- the C ABI loads and exports every symbol the
header declares, the native HEOM table builder is bit-exact against the reference goldens,
and compute entry points fail loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import cases
from conftest import golden, relerr, ROOT


def test_library_exports_every_declared_symbol():
    from lime_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'lime_b200.h')).read()
    declared = set(re.findall(r'\b(limeb200_\w+)\s*\(', hdr))
    assert declared, 'no declarations found'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    l = _lib.lib()
    for name in declared:
        assert hasattr(l, name), name
    assert l.limeb200_version() == 100


def test_heom_tables_native_bit_exact():
    from lime_b200 import engine
    g = golden('heom_tables')
    for key in g.files:
        dpart, x = key[1:].split('_x')
        dims = [int(v) for v in dpart.split('_')]
        states, dn, up = engine.heom_tables(dims, int(x))
        assert states.dtype == np.int32
        assert np.array_equal(states, g[key].astype(np.int32)), key
    import lime_oracle as lo
    for dims, exc in [([13, 13], 12), ([4, 3, 2], 3), ([5] * 6, 4), ([3, 3], 0)]:
        states, dn, up = engine.heom_tables(dims, exc)
        so, dno, upo = lo.heom_tables(dims, exc)
        assert np.array_equal(states, so) and np.array_equal(dn, dno) and np.array_equal(up, upo)
    states, dn, up = engine.heom_tables([5] * 14, 4)
    assert states.shape == (3060, 14)
    assert (dn >= 0).sum() + (up >= 0).sum() == 19040
    tiers = np.bincount(states.sum(axis=1))
    assert list(tiers) == [1, 14, 105, 560, 2380]


def test_enr_state_dictionaries_drop_in():
    from lime_b200.heom.heom import enr_state_dictionaries, state_number_enumerate, _calc_matsubara_params
    import lime_oracle as lo
    for dims, exc in [([3, 3], 2), ([4, 3, 2], 3), ([3] * 4, 0)]:
        n, s2i, i2s = enr_state_dictionaries(dims, exc)
        no, s2io, i2so = lo.enr_state_dictionaries(dims, exc)
        assert n == no and s2i == s2io and i2s == i2so
        assert list(s2i.keys()) == list(s2io.keys())           # same insertion order
        assert all(isinstance(k, tuple) for k in s2i)
    with pytest.raises(TypeError):
        enr_state_dictionaries([2, 2], None)
    assert [tuple(s) for s in state_number_enumerate([2, 2])] == [(0, 0), (0, 1), (1, 0), (1, 1)]
    g = golden('heom_matsubara')
    for i in range(4):
        K, lam, gam, T = g['par%d' % i]
        c, nu = _calc_matsubara_params(int(K), lam, gam, T)
        assert isinstance(c, list) and isinstance(nu, list)
        assert relerr(c, g['c%d' % i]) == 0 and relerr(nu, g['nu%d' % i]) == 0


def test_redfield_tensor_host_setup_matches_golden():
    import io, contextlib
    from lime_b200.oqs import Redfield_solver, redfield_tensor
    g = golden('redfield_example')
    H, a_ops, spectra, rho0, dt, Nt, e_ops, tlist = cases.redfield_example()
    s = Redfield_solver(H, c_ops=a_ops, spectra=spectra)
    R, evecs = s.redfield_tensor()
    assert relerr(R.toarray(), g['R']) <= 1e-14 and relerr(evecs, g['evecs']) == 0
    R2, _ = redfield_tensor(H, a_ops, spectra)
    assert relerr(R2.toarray(), g['R']) <= 1e-14
    with pytest.raises(TypeError):
        Redfield_solver(H, c_ops=a_ops).redfield_tensor()              # no spectra
    with pytest.raises(TypeError):
        redfield_tensor(H, [np.array([[0, 1.], [0, 0]])], spectra)      # non-Hermitian a_op
    with pytest.raises(TypeError):
        Redfield_solver(H, c_ops=a_ops, spectra=spectra).propagator(tlist)
    with pytest.raises(ValueError):
        s.correlation_4op_3t(rho0, [e_ops[0]] * 3, 'lll', tlist[:4])


def test_superoperator_algebra_matches_golden():
    from lime_b200 import superoperator as sop
    g = golden('lindblad_dense')
    H, c_ops, e_ops, rho0 = cases.lindblad_dense()
    L = sop.liouvillian(H, c_ops)
    assert relerr(L.toarray(), g['superop']) <= 1e-14
    with pytest.raises(ValueError):
        sop.operator_to_superoperator(H, 'x')


def test_result_defaults():
    from lime_b200.mol import Result
    r = Result(dt=0.1, Nt=5, rho0=np.eye(2))
    assert np.allclose(r.times, 0.1 * np.arange(5)) and r.timesteps == 5


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback():
    from lime_b200 import _lib
    from lime_b200.oqs import _lindblad
    H, c_ops, e_ops, rho0 = cases.lindblad_dense()
    with pytest.raises(_lib.LimeB200Error):
        _lindblad(H, rho0, c_ops, e_ops=e_ops, Nt=2, dt=0.01)
    h = ctypes.c_void_p()
    rc = _lib.lib().limeb200_qme_create(ctypes.byref(h), 4, 0)
    assert rc < 0 and b'no CUDA device' in _lib.lib().limeb200_last_error()


def test_qme_plan_analysis_host_logic(monkeypatch):
    """kernel selection, bandwidth-reducing basis order and cluster geometry (limeb200_qme_finalize's host logic)
    through an analysis-only plan -- no GPU involved"""
    from scipy.sparse import csr_matrix
    from lime_b200 import engine
    from lime_b200.oqs import lindblad_generator
    import cases

    def jc(ncav, **kw):
        H, c_ops, e_ops, rho0 = cases.jc_point(ncav=ncav, **kw)
        G, Gr, ls = lindblad_generator(csr_matrix(H), [csr_matrix(c) for c in c_ops])
        return G, [(l, l) for l in ls], e_ops

    G, sw, e_ops = jc(64)                                   # config 2: N = 128
    a = engine.analyze_qme(128, G, sw, e_ops)
    # chain-structured operators: register-patch kernel, two parity paths, cluster of 4, 32 own rows per CTA
    assert a['path'] == 6 and a['cluster'] == 4 and a['rows_per_cta'] == 32 and a['chain'] == 2
    monkeypatch.setenv('LIMEB200_NO_TILE', '1')             # the band kernel it replaced stays selectable
    a = engine.analyze_qme(128, G, sw, e_ops)
    assert a['path'] == 5 and a['cluster'] == 4 and a['rows_per_cta'] == 32
    assert a['noff'] == 2 and a['imag_offdiag'] == 1 and a['real_xz'] == 1 and a['chain'] == 0
    assert sorted(a['perm']) == list(range(128)) and a['permuted'] == 1
    assert 1 <= a['bandwidth'] <= 4
    # every operator entry lies within the band in the chosen order
    inv = np.argsort(a['perm'])
    for op in [G] + [x for x, _ in sw]:
        r, c = op.nonzero()
        assert np.max(np.abs(inv[r] - inv[c])) <= a['bandwidth']
    # opt-in chain variant: two interleaved parity chains, G couples rows 2 apart
    monkeypatch.setenv('LIMEB200_BAND_CHAIN', '1')
    b = engine.analyze_qme(128, G, sw, e_ops)
    assert b['path'] == 5 and b['chain'] == 2 and b['bandwidth'] == 3
    inv = np.argsort(b['perm'])
    r, c = (G - csr_matrix(np.diag(G.diagonal()))).nonzero()
    assert set(np.abs(inv[r] - inv[c])) == {2}
    monkeypatch.delenv('LIMEB200_BAND_CHAIN')
    monkeypatch.delenv('LIMEB200_NO_TILE')
    # too large for a cluster: generic sparse kernel; small / dense operands: dense kernels
    G2, sw2, e2 = jc(128)
    assert engine.analyze_qme(256, G2, sw2, e2)['path'] in (3, 4)
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=6)
    Gd, _, ls = lindblad_generator(H, c_ops)
    assert engine.analyze_qme(6, Gd, [(l, l) for l in ls])['path'] == 1
    H, c_ops, e_ops, rho0 = cases.lindblad_dense(n=100, M=1)
    Gd, _, ls = lindblad_generator(H, c_ops)
    assert engine.analyze_qme(100, Gd, [(l, l) for l in ls])['path'] == 2
    # an analysis-only plan refuses to run
    import ctypes as C
    from lime_b200._lib import lib
    h = C.c_void_p()
    assert lib().limeb200_qme_create(C.byref(h), 4, -1) == 0
    g = np.zeros((4, 4), dtype=complex)
    assert lib().limeb200_qme_set_generator_dense(h, g.ctypes.data_as(C.c_void_p), 1) == 0
    assert lib().limeb200_qme_finalize(h) == 0
    assert lib().limeb200_qme_run(h, None, 1, 0.1, 1, None, None, None, 1, None) < 0
    assert b'no CPU fallback' in lib().limeb200_last_error()
    lib().limeb200_qme_destroy(h)


def test_bench_input_builders_match_the_oracle_builders():
    """the synthetic inputs the GPU arm propagates (lime_b200/builders.py) are the operators the CPU arm builds
    (oracle jaynes_cummings = lime's Composite.getH layout, lime/cavity.py:57-97)"""
    from scipy.sparse import csr_matrix
    from lime_b200 import builders as models
    import lime_oracle as lo
    omega0, kappa, ncav = 1.0, 0.05, 12
    g = np.array([0.01, 0.1, 0.2])
    det = np.array([-0.2, 0.0, 0.15])
    pat, vals, c_ops, e_ops, rho0 = models.jaynes_cummings_batch(omega0, omega0 + det, g, ncav, kappa)
    pat = csr_matrix(pat)
    rows = np.repeat(np.arange(2 * ncav), np.diff(pat.indptr))
    for b in range(3):
        H, c_o, e_o = lo.jaynes_cummings(omega0, omega0 + det[b], g[b], ncav, kappa)
        Hb = csr_matrix((vals[b], (rows, pat.indices)), shape=pat.shape).toarray()
        assert np.allclose(Hb, H.toarray(), atol=0, rtol=1e-15)
        assert np.array_equal(c_ops[0].toarray(), c_o[0].toarray())
        for x, y in zip(e_ops, e_o):
            assert np.array_equal(x.toarray(), y.toarray())
    assert rho0[ncav, ncav] == 1 and np.count_nonzero(rho0) == 1            # |e,0><e,0|, index = i_mol*ncav + n
    Hf, Q, lam, gam, kT = models.fmo_heom_inputs()
    assert Hf.shape == (7, 7) and np.allclose(Hf, Hf.conj().T) and len(Q) == 7
    assert all(np.array_equal(q, np.diag(np.diag(q))) and np.trace(q) == 1 for q in Q)
    assert abs(lam * 219474.6305 - 35.0) < 1e-9 and abs(kT * 315775.13 - 300.0) < 1e-9


class _FakeRkf45Lib:
    """NumPy stand-ins for the four limeb200_rkf45_* kernels (same argument lists), operating on CPU torch storage
    through raw pointers -- a seam for testing the HOST step controller of lime_b200/rkf45.py without a GPU."""

    @staticmethod
    def _a(ptr, n):
        import ctypes as C
        v = ptr.value if hasattr(ptr, 'value') else ptr
        return None if not v else np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_double)), shape=(n,))

    def limeb200_rkf45_stage(self, stage, n, y, yp, f1, f2, f3, f4, f5, h, out, st):
        a = self._a
        y, p, f1, f2, f3, f4, out = a(y, n), a(yp, n), a(f1, n), a(f2, n), a(f3, n), a(f4, n), a(out, n)
        if stage == 1:
            out[:] = y + (h / 4.0) * p
        elif stage == 2:
            out[:] = y + (3.0 * h / 32.0) * (p + 3.0 * f1)
        elif stage == 3:
            out[:] = y + (h / 2197.0) * (1932.0 * p + (7296.0 * f2 - 7200.0 * f1))
        elif stage == 4:
            out[:] = y + (h / 4104.0) * ((8341.0 * p - 845.0 * f3) + (29440.0 * f2 - 32832.0 * f1))
        else:
            out[:] = y + (h / 20520.0) * ((-6080.0 * p + (9295.0 * f3 - 5643.0 * f4)) + (41040.0 * f1 - 28352.0 * f2))
        return 0

    def limeb200_rkf45_error(self, n, y, yp, f2, f3, f4, f5, h, ae, s, res, st):
        a = self._a
        y, p, f2, f3, f4, f5, s, res = a(y, n), a(yp, n), a(f2, n), a(f3, n), a(f4, n), a(f5, n), a(s, n), a(res, 2)
        s[:] = y + (h / 7618050.0) * ((902880.0 * p + (3855735.0 * f3 - 1371249.0 * f4)) + (3953664.0 * f2 + 277020.0 * f5))
        et = np.abs(y) + np.abs(s) + ae
        ee = np.abs((-2090.0 * p + (21970.0 * f3 - 15048.0 * f4)) + (22528.0 * f2 - 27360.0 * f5))
        res[0] = np.max(np.where(et > 0, ee / np.where(et > 0, et, 1.0), 0.0))
        res[1] = np.min(et)
        return 0

    def limeb200_rkf45_hinit(self, n, y, yp, relerr, abserr, h0, res, st):
        y, p, res = self._a(y, n), self._a(yp, n), self._a(res, 2)
        tol = relerr * np.abs(y) + abserr
        res[0] = np.max(np.where(tol > 0, tol, 0.0))
        h = h0
        for k in range(n):
            if tol[k] > 0 and tol[k] < abs(p[k]) * h0 ** 5:
                h = min(h, (tol[k] / abs(p[k])) ** 0.2)
        res[1] = h
        return 0

    def limeb200_rkf45_axpy(self, n, a, x, y, st):
        x, y = self._a(x, n), self._a(y, n)
        y[:] = y + a * x
        return 0


@pytest.fixture
def rkf45_on_host(monkeypatch):
    import torch
    from lime_b200 import rkf45, _dev
    fake = _FakeRkf45Lib()
    cpu = torch.device('cpu')
    monkeypatch.setattr(rkf45, 'lib', lambda: fake)
    monkeypatch.setattr(_dev, 'device', lambda index=None: cpu)
    monkeypatch.setattr(_dev, 'stream_ptr', lambda dev=None: None)
    monkeypatch.setattr(_dev, 'to_dev', lambda a, dtype=np.complex128, dev=None, pinned=False:
                        torch.from_numpy(np.array(a, dtype=dtype, copy=True)))
    monkeypatch.setattr(rkf45._Vec, '__init__', _vec_init_cpu)
    return rkf45


def _vec_init_cpu(self, y, dev):
    import torch
    self.host = True
    a = np.asarray(y)
    self.cplx = np.iscomplexobj(a)
    self.shape = a.shape
    self.t = torch.from_numpy(np.array(a, dtype=np.complex128 if self.cplx else np.float64, copy=True))
    self.real = torch.view_as_real(self.t).reshape(-1) if self.cplx else self.t.reshape(-1)
    self.n = self.real.numel()


def test_rkf45_step_controller_host_logic(rkf45_on_host):
    """the Shampine-Watts controller of lime_b200/rkf45.py on lime's own test problems (examples/rkf45_test.py:72-119,
    153-202, 283-330) with the device kernels replaced by NumPy through a seam: flags, output times, accuracy"""
    rk = rkf45_on_host
    tol = float(np.sqrt(np.finfo(np.double).eps))

    def f1(t, y):
        return np.array([0.25 * y[0] * (1.0 - y[0] / 20.0)])

    def exact1(t):
        return 20.0 / (1.0 + 19.0 * np.exp(-0.25 * t))

    # test04: scalar logistic equation, 5 output intervals on [0, 20]
    S = rk.RKF45State()
    y, flag, t = np.array([1.0]), 1, 0.0
    yp = f1(t, y)
    for i in range(1, 6):
        t, tout = (i - 1) * 4.0, i * 4.0
        y, yp, t, flag = rk.r8_rkf45(f1, 1, y, yp, t, tout, tol, tol, flag, state=S)
        assert flag == 2 and t == tout
        assert abs(y[0] - exact1(t)) < 2e-6
        assert abs(yp[0] - f1(t, y)[0]) < 1e-14
    # test05: harmonic oscillator, 12 intervals on [0, 2 pi]
    def f2(t, y):
        return np.array([y[1], -y[0]])
    S = rk.RKF45State()
    y, flag = np.array([1.0, 0.0]), 1
    yp = f2(0.0, y)
    for i in range(1, 13):
        t, tout = (i - 1) * 2 * np.pi / 12, i * 2 * np.pi / 12
        y, yp, t, flag = rk.r8_rkf45(f2, 2, y, yp, t, tout, tol, tol, flag, state=S)
        assert flag == 2
        assert abs(y[0] - np.cos(t)) < 1e-6 and abs(y[1] + np.sin(t)) < 1e-6
    # test06: single-step mode, flag -1 / -2 until the output point answers +2
    S = rk.RKF45State()
    y, flag, t = np.array([1.0]), -1, 0.0
    yp = f1(t, y)
    nsteps = 0
    for i in range(1, 6):
        tout = i * 4.0
        while flag < 0:
            y, yp, t, flag = rk.r8_rkf45(f1, 1, y, yp, t, tout, tol, tol, flag, state=S)
            nsteps += 1
            assert flag in (-2, 2) and t <= tout
        assert t == tout and abs(y[0] - exact1(t)) < 2e-6
        flag = -2
    assert nsteps > 5 and S.steps_accepted == nsteps
    # input validation and the "relerr too small" answer
    assert rk.r8_rkf45(f1, 0, y, yp, 0.0, 1.0, tol, tol, 1)[3] == 8
    assert rk.r8_rkf45(f1, 1, y, yp, 0.0, 1.0, -1.0, tol, 1)[3] == 8
    assert rk.r8_rkf45(f1, 1, y, yp, 0.0, 1.0, tol, tol, 0)[3] == 8
    assert rk.r8_rkf45(f1, 1, y, yp, 0.0, 1.0, 1e-20, tol, 1, state=rk.RKF45State())[3] == 3
    # complex state: d rho/dt = -i [H, rho] for a two-level system against the exact propagator
    H = np.array([[0.3, 0.2 - 0.1j], [0.2 + 0.1j, -0.4]])
    rho0 = np.array([[0.7, 0.2j], [-0.2j, 0.3]], dtype=complex)
    states, S = rk.integrate(lambda t, r: -1j * (H @ r - r @ H), rho0, np.linspace(0, 3.0, 4), relerr=1e-10, abserr=1e-12)
    w, v = np.linalg.eigh(H)
    for t, r in zip(np.linspace(0, 3.0, 4), states):
        U = v @ np.diag(np.exp(-1j * w * t)) @ v.conj().T
        assert np.max(np.abs(r - U @ rho0 @ U.conj().T)) < 1e-8


def test_obs_every_plumbing_and_device_subsampling(monkeypatch):
    """obs_every of the batch APIs: the sub-sampling helper keeps the samples after steps k, 2k, ... and both solvers
    hand the argument down to the layer that copies the observables (checked with stand-in plans: no GPU needed)"""
    from lime_b200 import engine, oqs
    o = torch.arange(10 * 2 * 3, dtype=torch.float64).reshape(10, 2, 3)
    assert engine.subsample_steps(o, 1) is o and engine.subsample_steps(None, 4) is None
    r = engine.subsample_steps(o, 4)
    assert r.is_contiguous() and torch.equal(r, o[[3, 7]])
    assert engine.subsample_steps(o, 11).shape == (0, 2, 3)
    seen = {}

    class FakePlan:
        dev = None

        def run(self, rho0, dt, nsteps, **kw):
            seen['run'] = kw
            return 'rho', 'obs', 'traj'
    monkeypatch.setattr(oqs, '_lindblad_plan', lambda *a, **k: FakePlan())
    H, c_ops, e_ops, rho0 = cases.jc_point(ncav=4)
    s = oqs.Lindblad_solver(H, c_ops=c_ops)
    assert s.evolve_batch(rho0, 0.01, 8, e_ops=e_ops, obs_every=4) == ('rho', 'obs', 'traj')
    assert seen['run']['obs_every'] == 4 and seen['run']['traj_every'] == 0
    s.evolve_batch(rho0, 0.01, 8, e_ops=e_ops)
    assert seen['run']['obs_every'] == 1
    # Redfield: tensor form -> engine.liouville_rk4, operator form -> the plan's run
    Hs, a_ops, spectra, r0, dt, Nt, e_r, tlist = cases.redfield_example()
    rs = oqs.Redfield_solver(Hs, c_ops=a_ops, spectra=spectra)
    rs.redfield_tensor()

    def fake_rk4(R, v0, dt, nsteps, **kw):
        seen['rk4'] = kw
        return np.asarray(v0), None, None
    monkeypatch.setattr(engine, 'liouville_rk4', fake_rk4)
    rs.evolve_batch(r0, dt, 8, e_ops=e_r, obs_every=2)
    assert seen['rk4']['obs_every'] == 2
    monkeypatch.setattr(rs, 'operator_plan', lambda e: FakePlan())
    rs.evolve_batch(r0, dt, 8, e_ops=e_r, obs_every=2, form='operator')
    assert seen['run']['obs_every'] == 2

