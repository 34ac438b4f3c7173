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
