"""
Thin object layer over the C ABI (include/lime_b200.h): plans own the native handle,
numpy/scipy operands go in, device tensors (torch) carry the state.
"""
import ctypes as C
import numpy as np
import torch
from scipy.sparse import issparse, csr_matrix

from ._lib import lib, check, hptr, LimeB200Error
from . import _dev

(PATH_AUTO, PATH_DENSE_ONCHIP, PATH_DENSE_STAGE, PATH_SPARSE_GLOBAL, PATH_SPARSE_CLUSTER, PATH_SPARSE_BAND,
 PATH_SPARSE_TILE) = range(7)


def _dense_batch(op, N):
    """-> (array [nb][N][N] complex128, nb)"""
    a = _dev.as_c128(op)
    if a.ndim == 2:
        a = a[None]
    if a.shape[1:] != (N, N):
        raise ValueError('operator shape %s does not match N=%d' % (a.shape, N))
    return np.ascontiguousarray(a), a.shape[0]


def _csr_parts(op, data_batch=None):
    """scipy sparse -> (indptr int32, indices int32, data [nb][nnz] complex128, nnz, nb)"""
    m = csr_matrix(op)
    m.sum_duplicates()
    indptr = np.ascontiguousarray(m.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(m.indices, dtype=np.int32)
    if data_batch is None:
        data = np.ascontiguousarray(m.data, dtype=np.complex128)[None]
    else:
        data = np.ascontiguousarray(data_batch, dtype=np.complex128)
        if data.ndim != 2 or data.shape[1] != m.nnz:
            raise ValueError('batched CSR values must be [nb][nnz]')
    return indptr, indices, np.ascontiguousarray(data), int(m.nnz), data.shape[0]


class QmePlan:
    """d rho/dt = G rho + rho G^H + sum_s X_s rho Z_s^H  on the device.

    G, X_s, Z_s: ndarray [N,N] or [nb,N,N], or scipy.sparse (shared pattern; batched values
    through the *_csr_batch helpers)."""

    def __init__(self, N, device_index=None):
        self.N = int(N)
        self.dev = _dev.device(device_index)
        self._h = C.c_void_p()
        check(lib().limeb200_qme_create(C.byref(self._h), self.N, self.dev.index))
        self.E = 0
        self.ndrive = 0
        self.finalized = False

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            try:
                lib().limeb200_qme_destroy(h)
            except Exception:
                pass

    # ---- operators -------------------------------------------------------------
    def set_generator(self, G):
        if issparse(G):
            ip, ix, d, nnz, nb = _csr_parts(G)
            check(lib().limeb200_qme_set_generator_csr(self._h, hptr(ip), hptr(ix), hptr(d), nnz, nb))
        else:
            a, nb = _dense_batch(G, self.N)
            check(lib().limeb200_qme_set_generator_dense(self._h, hptr(a), nb))

    def set_generator_csr_batch(self, pattern, data):
        """pattern: scipy.sparse with the shared sparsity; data [nb][nnz] in pattern.tocsr() order"""
        ip, ix, d, nnz, nb = _csr_parts(pattern, data)
        check(lib().limeb200_qme_set_generator_csr(self._h, hptr(ip), hptr(ix), hptr(d), nnz, nb))

    def set_step_values(self, on=True):
        """the generator batch index is the RK4 step (time-dependent sparse generator)"""
        check(lib().limeb200_qme_set_step_values(self._h, 1 if on else 0))

    def add_sandwich(self, X, Z=None):
        Z = X if Z is None else Z
        if issparse(X) and issparse(Z):
            xp = _csr_parts(X)
            zp = _csr_parts(Z)
            check(lib().limeb200_qme_add_sandwich_csr(self._h, hptr(xp[0]), hptr(xp[1]), hptr(xp[2]), xp[3],
                                                      hptr(zp[0]), hptr(zp[1]), hptr(zp[2]), zp[3], 1))
        else:
            x, nbx = _dense_batch(X, self.N)
            z, nbz = _dense_batch(Z, self.N)
            nb = max(nbx, nbz)
            if nbx != nb:
                x = np.ascontiguousarray(np.broadcast_to(x, (nb, self.N, self.N)))
            if nbz != nb:
                z = np.ascontiguousarray(np.broadcast_to(z, (nb, self.N, self.N)))
            check(lib().limeb200_qme_add_sandwich_dense(self._h, hptr(x), hptr(z), nb))

    def set_right_generator(self, Gr):
        """explicit right generator (rho Gr); needed when H is not Hermitian"""
        a, nb = _dense_batch(Gr, self.N)
        check(lib().limeb200_qme_set_right_generator_dense(self._h, hptr(a), nb))

    def add_drive(self, D, Dr=None):
        """G_k += c_k D ; Gr_k += c_k Dr  (Dr None: Gr_k += conj(c_k) D^H)"""
        a, nb = _dense_batch(D, self.N)
        if nb != 1:
            raise ValueError('drive operators are not batched')
        r = None if Dr is None else _dense_batch(Dr, self.N)[0]
        check(lib().limeb200_qme_add_drive_dense(self._h, hptr(a), hptr(r)))
        self.ndrive += 1

    def set_observables(self, e_ops):
        e_ops = [] if e_ops is None else list(e_ops)
        self.E = len(e_ops)
        if self.E:
            a = np.ascontiguousarray(np.stack([_dev.as_c128(e) for e in e_ops]))
            if a.shape[1:] != (self.N, self.N):
                raise ValueError('observable shape mismatch')
            check(lib().limeb200_qme_set_observables(self._h, hptr(a), self.E))

    def set_path(self, path):
        check(lib().limeb200_qme_set_path(self._h, int(path)))

    def finalize(self):
        check(lib().limeb200_qme_finalize(self._h))
        self.finalized = True
        return self

    @property
    def path(self):
        return lib().limeb200_qme_get_path(self._h)

    @property
    def last_launches(self):
        return int(lib().limeb200_qme_last_launches(self._h))

    # ---- execution ------------------------------------------------------------
    def run_device(self, rho, dt, nsteps, coef=None, want_obs=True, traj_every=0):
        """rho: device tensor [B,N,N] complex128, advanced IN PLACE by nsteps RK4 steps.
        Returns (obs [nsteps,B,E] or None, traj [nsteps//traj_every,B,N,N] or None) on device."""
        if not self.finalized:
            self.finalize()
        assert rho.dtype == torch.complex128 and rho.is_contiguous() and rho.dim() == 3
        B = rho.shape[0]
        obs = _dev.empty((nsteps, B, self.E), dev=self.dev) if (want_obs and self.E) else None
        traj = _dev.empty((nsteps // traj_every, B, self.N, self.N), dev=self.dev) if traj_every else None
        check(lib().limeb200_qme_run(self._h, _dev.ptr(rho), B, float(dt), int(nsteps), _dev.ptr(coef),
                                     _dev.ptr(obs), _dev.ptr(traj), int(traj_every), _dev.stream_ptr(self.dev)))
        return obs, traj

    def run(self, rho0, dt, nsteps, coef=None, traj_every=0, pinned=False, obs_every=1):
        """host in / host out: rho0 [N,N] or [B,N,N] -> (rho_final, obs, traj) as numpy.
        rho0 may be a (pinned) torch CPU tensor; pinned=True returns views of cached
        page-locked buffers (asynchronous copies in both directions).
        obs_every = k > 1: only the observables after steps k, 2k, ... come back to the host ([nsteps // k, B, E];
        sub-sampled on the device, so the device->host stream shrinks by k -- large batches are otherwise PCIe-bound
        on the 16 bytes per step and unit of the observable stream)."""
        r = rho0 if isinstance(rho0, torch.Tensor) else _dev.as_c128(rho0)
        single = r.ndim == 2
        if single:
            r = r[None]
        d = _dev.h2d(r, dev=self.dev)
        c = None if coef is None else _dev.to_dev(np.asarray(coef).reshape(nsteps, -1), dev=self.dev)
        obs, traj = self.run_device(d, dt, nsteps, coef=c, traj_every=traj_every)
        obs = subsample_steps(obs, obs_every)
        out = _dev.d2h(d, pinned)
        obs = _dev.d2h(obs, pinned)
        traj = _dev.d2h(traj, pinned)
        if single:
            out = out[0]
            obs = None if obs is None else obs[:, 0]
            traj = None if traj is None else traj[:, 0]
        return out, obs, traj

    def rhs(self, rho):
        """one right-hand side (numpy in/out)"""
        if not self.finalized:
            self.finalize()
        r = _dev.as_c128(rho)
        single = r.ndim == 2
        if single:
            r = r[None]
        d = _dev.to_dev(r, dev=self.dev)
        o = torch.empty_like(d)
        check(lib().limeb200_qme_rhs(self._h, _dev.ptr(d), _dev.ptr(o), d.shape[0], _dev.stream_ptr(self.dev)))
        out = o.cpu().numpy()
        return out[0] if single else out


def subsample_steps(obs, every):
    """device observables [nsteps, ...] -> the samples after steps every, 2*every, ... (contiguous copy on the device);
    every <= 1 or obs None: unchanged"""
    every = int(every)
    if obs is None or every <= 1:
        return obs
    return obs[every - 1::every].contiguous()


def liouville_rk4(R, v0, dt, nsteps, e_rows=None, traj_every=0, dev=None, obs_every=1):
    """dv/dt = R v (R scipy.sparse / ndarray, D x D); v0 [D] or [B,D].
    Returns (v_final, obs [nsteps,B,E] or None, traj [nsteps//traj_every,B,D] or None), numpy.
    obs_every = k > 1: observables after steps k, 2k, ... only (sub-sampled on the device before the copy)."""
    dev = _dev.device() if dev is None else dev
    m = csr_matrix(R).astype(np.complex128)
    m.sum_duplicates()
    m.sort_indices()
    D = m.shape[0]
    v = np.ascontiguousarray(np.asarray(v0, dtype=np.complex128))
    single = v.ndim == 1
    if single:
        v = v[None]
    B = v.shape[0]
    d_ip = _dev.to_dev(m.indptr, np.int32, dev)
    d_ix = _dev.to_dev(m.indices, np.int32, dev)
    d_da = _dev.to_dev(m.data, np.complex128, dev)
    d_v = _dev.to_dev(v, dev=dev)
    E = 0 if e_rows is None else len(e_rows)
    d_e = _dev.to_dev(np.stack([np.asarray(e, dtype=np.complex128).reshape(D) for e in e_rows]), dev=dev) if E else None
    obs = _dev.empty((nsteps, B, E), dev=dev) if E else None
    traj = _dev.empty((nsteps // traj_every, B, D), dev=dev) if traj_every else None
    check(lib().limeb200_liouville_rk4_csr(_dev.ptr(d_ip), _dev.ptr(d_ix), _dev.ptr(d_da), D, _dev.ptr(d_v), B,
                                           _dev.ptr(d_e), E, _dev.ptr(obs), _dev.ptr(traj), int(traj_every),
                                           float(dt), int(nsteps), _dev.stream_ptr(dev)))
    out = d_v.cpu().numpy()
    obs = subsample_steps(obs, obs_every)
    obs = None if obs is None else obs.cpu().numpy()
    traj = None if traj is None else traj.cpu().numpy()
    if single:
        out = out[0]
        obs = None if obs is None else obs[:, 0]
        traj = None if traj is None else traj[:, 0]
    return out, obs, traj


def analyze_qme(N, G, sandwiches=(), e_ops=None, path=None):
    """What limeb200_qme_finalize would decide for these operators on a B200 -- host only, no GPU needed
    (analysis-only plan, device = -1).  Returns dict(path, permuted, bandwidth, noff, imag_offdiag, real_xz,
    cluster, rows_per_cta, chain, perm)."""
    h = C.c_void_p()
    check(lib().limeb200_qme_create(C.byref(h), int(N), -1))
    try:
        if issparse(G):
            ip, ix, d, nnz, nb = _csr_parts(G)
            check(lib().limeb200_qme_set_generator_csr(h, hptr(ip), hptr(ix), hptr(d), nnz, nb))
        else:
            a, nb = _dense_batch(G, N)
            check(lib().limeb200_qme_set_generator_dense(h, hptr(a), nb))
        for X, Z in sandwiches:
            if issparse(X) and issparse(Z):
                xp, zp = _csr_parts(X), _csr_parts(Z)
                check(lib().limeb200_qme_add_sandwich_csr(h, hptr(xp[0]), hptr(xp[1]), hptr(xp[2]), xp[3],
                                                          hptr(zp[0]), hptr(zp[1]), hptr(zp[2]), zp[3], 1))
            else:
                x, _ = _dense_batch(X, N)
                z, _ = _dense_batch(Z, N)
                check(lib().limeb200_qme_add_sandwich_dense(h, hptr(x), hptr(z), 1))
        if e_ops:
            e = np.ascontiguousarray(np.stack([_dev.as_c128(o) for o in e_ops]))
            check(lib().limeb200_qme_set_observables(h, hptr(e), len(e_ops)))
        if path is not None:
            check(lib().limeb200_qme_set_path(h, int(path)))
        check(lib().limeb200_qme_finalize(h))
        info = np.zeros(9, dtype=np.int32)
        perm = np.zeros(N, dtype=np.int32)
        check(lib().limeb200_qme_get_info(h, hptr(info), hptr(perm)))
    finally:
        lib().limeb200_qme_destroy(h)
    keys = ['path', 'permuted', 'bandwidth', 'noff', 'imag_offdiag', 'real_xz', 'cluster', 'rows_per_cta', 'chain']
    out = {k: int(v) for k, v in zip(keys, info)}
    out['perm'] = perm
    return out


def zgemm(A, Bm):
    """C = A @ B on the FP64 tensor cores.  A [M,K] or [b,M,K], B [K,N] or [b,K,N]: numpy arrays (uploaded) or
    device tensors; returns a device tensor [M,N] / [b,M,N]."""
    dev = _dev.device()
    ta = A if isinstance(A, torch.Tensor) else _dev.to_dev(np.asarray(A), dev=dev)
    tb = Bm if isinstance(Bm, torch.Tensor) else _dev.to_dev(np.asarray(Bm), dev=dev)
    assert ta.dtype == torch.complex128 and tb.dtype == torch.complex128
    ta, tb = ta.contiguous(), tb.contiguous()
    batch = max(ta.shape[0] if ta.dim() == 3 else 1, tb.shape[0] if tb.dim() == 3 else 1)
    M, K = ta.shape[-2], ta.shape[-1]
    K2, N = tb.shape[-2], tb.shape[-1]
    if K != K2:
        raise ValueError('zgemm: inner dimensions %d and %d differ' % (K, K2))
    sA = M * K if ta.dim() == 3 else 0
    sB = K * N if tb.dim() == 3 else 0
    out = _dev.empty((batch, M, N), dev=dev)
    check(lib().limeb200_zgemm(_dev.ptr(ta), _dev.ptr(tb), _dev.ptr(out), M, N, K, batch, sA, sB, M * N,
                               _dev.stream_ptr(dev)))
    return out if (ta.dim() == 3 or tb.dim() == 3) else out[0]


class LiouvillePlan:
    """device-resident form of liouville_rk4: R (CSR) and the observable rows are uploaded once;
    run_device advances a device batch v [B, D] in place"""

    def __init__(self, R, e_rows=None, dev=None):
        self.dev = _dev.device() if dev is None else dev
        m = csr_matrix(R).astype(np.complex128)
        m.sum_duplicates()
        m.sort_indices()
        self.D = m.shape[0]
        self.ip = _dev.to_dev(m.indptr, np.int32, self.dev)
        self.ix = _dev.to_dev(m.indices, np.int32, self.dev)
        self.da = _dev.to_dev(m.data, np.complex128, self.dev)
        self.E = 0 if e_rows is None else len(e_rows)
        self.e = _dev.to_dev(np.stack([np.asarray(e, dtype=np.complex128).reshape(self.D) for e in e_rows]),
                             dev=self.dev) if self.E else None

    def run_device(self, v, dt, nsteps, traj_every=0):
        assert v.dtype == torch.complex128 and v.is_contiguous() and v.shape[1] == self.D
        B = v.shape[0]
        obs = _dev.empty((nsteps, B, self.E), dev=self.dev) if self.E else None
        traj = _dev.empty((nsteps // traj_every, B, self.D), dev=self.dev) if traj_every else None
        check(lib().limeb200_liouville_rk4_csr(_dev.ptr(self.ip), _dev.ptr(self.ix), _dev.ptr(self.da), self.D,
                                               _dev.ptr(v), B, _dev.ptr(self.e), self.E, _dev.ptr(obs), _dev.ptr(traj),
                                               int(traj_every), float(dt), int(nsteps), _dev.stream_ptr(self.dev)))
        return obs, traj


# ---------------------------------------------------------------------------------------
# HEOM
# ---------------------------------------------------------------------------------------
def heom_tables(dims, excitations):
    """(states, dn, up) int32 arrays [N_he, N_m] from the native table builder
    (bit-exact with lime/heom/heom.py:21-108; connectivity :176-216)"""
    dims = np.ascontiguousarray(dims, dtype=np.int32)
    nm = len(dims)
    exc = 0 if not excitations else int(excitations)
    nhe = lib().limeb200_heom_count_states(hptr(dims), nm, exc)
    check(nhe)
    states = np.empty((nhe, nm), dtype=np.int32)
    dn = np.empty((nhe, nm), dtype=np.int32)
    up = np.empty((nhe, nm), dtype=np.int32)
    check(lib().limeb200_heom_build_tables(hptr(dims), nm, exc, nhe, hptr(states), hptr(dn), hptr(up)))
    return states, dn, up


class HeomPlan:
    """multi-index HEOM hierarchy on the device (see include/lime_b200.h)"""

    def __init__(self, H, Q, qmap, c, nu, states, dn, up, pref_dn=-1j, pref_up=-1j,
                 row_range=None, device_index=None):
        self.dev = _dev.device(device_index)
        H = _dev.as_c128(H)
        self.n = H.shape[0]
        Q = np.ascontiguousarray(np.asarray(Q, dtype=np.complex128))
        if Q.ndim == 2:
            Q = Q[None]
        self.nq = Q.shape[0]
        self.states = np.ascontiguousarray(states, dtype=np.int32)
        self.nhe, self.nmodes = self.states.shape
        dn = np.ascontiguousarray(dn, dtype=np.int32)
        up = np.ascontiguousarray(up, dtype=np.int32)
        qmap = np.ascontiguousarray(np.broadcast_to(np.asarray(qmap, dtype=np.int32), (self.nmodes,)))
        c = np.ascontiguousarray(np.asarray(c, dtype=np.complex128))
        nu = np.ascontiguousarray(np.asarray(nu, dtype=np.float64))
        if c.ndim == 1:
            c, nu = c[None], nu[None]
        assert c.shape == nu.shape and c.shape[1] == self.nmodes
        self.npar = c.shape[0]
        lo, hi = (0, self.nhe) if row_range is None else row_range
        self.row_range = (int(lo), int(hi))
        pd = np.array([complex(pref_dn).real, complex(pref_dn).imag])
        pu = np.array([complex(pref_up).real, complex(pref_up).imag])
        self._h = C.c_void_p()
        check(lib().limeb200_heom_create_batched(C.byref(self._h), self.dev.index, self.n, self.nmodes, self.nq,
                                                 self.nhe, hptr(H), hptr(Q), hptr(qmap), hptr(c), hptr(nu), self.npar,
                                                 hptr(pd), hptr(pu), hptr(self.states), hptr(dn), hptr(up),
                                                 self.row_range[0], self.row_range[1]))

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            try:
                lib().limeb200_heom_destroy(h)
            except Exception:
                pass

    def set_path(self, path):
        check(lib().limeb200_heom_set_path(self._h, int(path)))

    @property
    def path(self):
        return lib().limeb200_heom_get_path(self._h)

    @property
    def last_launches(self):
        return int(lib().limeb200_heom_last_launches(self._h))

    def run_device(self, ado, dt, nsteps, eT=None, traj_every=0):
        """ado: device [B,N_he,n,n] advanced in place; eT: device [E,n,n] TRANSPOSED observables"""
        assert ado.dtype == torch.complex128 and ado.is_contiguous() and ado.dim() == 4
        B = ado.shape[0]
        E = 0 if eT is None else eT.shape[0]
        obs = _dev.empty((nsteps, B, E), dev=self.dev) if E else None
        traj = _dev.empty((nsteps // traj_every, B, self.n, self.n), dev=self.dev) if traj_every else None
        check(lib().limeb200_heom_run(self._h, _dev.ptr(ado), B, float(dt), int(nsteps), _dev.ptr(eT), E,
                                      _dev.ptr(obs), _dev.ptr(traj), int(traj_every), _dev.stream_ptr(self.dev)))
        return obs, traj

    def run(self, ado0, dt, nsteps, e_ops=None, traj_every=0):
        a = np.ascontiguousarray(np.asarray(ado0, dtype=np.complex128))
        single = a.ndim == 3
        if single:
            a = a[None]
        d = _dev.to_dev(a, dev=self.dev)
        eT = None
        if e_ops:
            eT = _dev.to_dev(np.stack([_dev.as_c128(e).T for e in e_ops]), dev=self.dev)
        obs, traj = self.run_device(d, dt, nsteps, eT=eT, traj_every=traj_every)
        out = d.cpu().numpy()
        obs = None if obs is None else obs.cpu().numpy()
        traj = None if traj is None else traj.cpu().numpy()
        if single:
            out = out[0]
            obs = None if obs is None else obs[:, 0]
            traj = None if traj is None else traj[:, 0]
        return out, obs, traj

    def rhs(self, ado):
        a = np.ascontiguousarray(np.asarray(ado, dtype=np.complex128))
        single = a.ndim == 3
        if single:
            a = a[None]
        d = _dev.to_dev(a, dev=self.dev)
        o = torch.zeros_like(d)
        check(lib().limeb200_heom_rhs(self._h, _dev.ptr(d), _dev.ptr(o), d.shape[0], _dev.stream_ptr(self.dev)))
        out = o.cpu().numpy()
        return out[0] if single else out

    def stage(self, stage, rho, yin, ynext, acc, dt):
        """one RK4 stage over the owned ADO range (device tensors [B,N_he,n,n])"""
        check(lib().limeb200_heom_stage(self._h, int(stage), _dev.ptr(rho), _dev.ptr(yin), _dev.ptr(ynext),
                                        _dev.ptr(acc), rho.shape[0], float(dt), _dev.stream_ptr(self.dev)))


def heom_dl_euler(H, sz, ado0, par, dt, nt, want_traj=True, dev=None):
    """_heom_dl-exact sweep.  ado0 [B,nado,n,n] (tier-major); par [B,3] = (gamma, a, b).
    Returns (ado_final [B,nado,n,n], traj [nt,B,n,n] or None) as numpy."""
    dev = _dev.device() if dev is None else dev
    H = _dev.as_c128(H)
    sz = _dev.as_c128(sz)
    n = H.shape[0]
    a = np.ascontiguousarray(np.asarray(ado0, dtype=np.complex128))
    B, nado = a.shape[0], a.shape[1]
    d = _dev.to_dev(a, dev=dev)
    p = _dev.to_dev(np.asarray(par, dtype=np.float64).reshape(B, 3), np.float64, dev)
    traj = _dev.empty((nt, B, n, n), dev=dev) if want_traj else None
    check(lib().limeb200_heom_dl_euler(hptr(H), hptr(sz), n, nado, _dev.ptr(d), _dev.ptr(p), B, float(dt), int(nt),
                                       _dev.ptr(traj), _dev.stream_ptr(dev)))
    return d.cpu().numpy(), (None if traj is None else traj.cpu().numpy())


# ---------------------------------------------------------------------------------------
# SOS building blocks
# ---------------------------------------------------------------------------------------
def sos_factor(z, W, p1, p2=None, dev=None):
    """F[t][q][n] = sum_d W[t][q][d] / (z_n - e1 + i g1) [/ (z_n - e2 + i g2)]
    z: device float64 [n]; W: host [T,R,D] complex; p1/p2: host [R,D,2] (e,g).  -> device [T,R,n]"""
    dev = _dev.device() if dev is None else dev
    W = np.ascontiguousarray(np.asarray(W, dtype=np.complex128))
    T, R, D = W.shape
    dW = _dev.to_dev(W, dev=dev)
    dp1 = _dev.to_dev(np.asarray(p1, dtype=np.float64).reshape(R, D, 2), np.float64, dev)
    dp2 = None if p2 is None else _dev.to_dev(np.asarray(p2, dtype=np.float64).reshape(R, D, 2), np.float64, dev)
    n = z.shape[0]
    F = _dev.empty((T, R, n), dev=dev)
    check(lib().limeb200_sos_factor(_dev.ptr(z), n, _dev.ptr(dW), _dev.ptr(dp1), _dev.ptr(dp2), T, R, D,
                                    _dev.ptr(F), _dev.stream_ptr(dev)))
    return F


def sos_factor_dev(z, dW, dp1, dp2=None):
    """same as sos_factor with the weights / poles already on the device: dW [T,R,D] complex128,
    dp1/dp2 [R,D,2] float64 -> device [T,R,n]"""
    T, R, D = dW.shape
    n = z.shape[0]
    F = _dev.empty((T, R, n), dev=z.device)
    check(lib().limeb200_sos_factor(_dev.ptr(z), n, _dev.ptr(dW), _dev.ptr(dp1), _dev.ptr(dp2), T, R, D,
                                    _dev.ptr(F), _dev.stream_ptr(z.device)))
    return F


def sos_factor_time(t, W, p1, dev=None):
    """F[T][q][n] = sum_d W[T][q][d] * (-i) theta(t_n) exp(-i e t_n - g t_n); t: device float64 [n];
    W: host [T,R,D] complex; p1: host [R,D,2] (e, g).  -> device [T,R,n]"""
    dev = _dev.device() if dev is None else dev
    W = np.ascontiguousarray(np.asarray(W, dtype=np.complex128))
    T, R, D = W.shape
    dW = _dev.to_dev(W, dev=dev)
    dp1 = _dev.to_dev(np.asarray(p1, dtype=np.float64).reshape(R, D, 2), np.float64, dev)
    n = t.shape[0]
    F = _dev.empty((T, R, n), dev=dev)
    check(lib().limeb200_sos_factor_time(_dev.ptr(t), n, _dev.ptr(dW), _dev.ptr(dp1), T, R, D,
                                         _dev.ptr(F), _dev.stream_ptr(dev)))
    return F


def sos_outer(A, Bf, T, scale=1.0, out=None, accumulate=False):
    """out[t][r][c] (+)= scale * sum_q A[ta][q][r] * B[tb][q][c]  (device tensors)"""
    TA, R, nrow = A.shape
    TB, R2, ncol = Bf.shape
    assert R == R2
    if out is None:
        out = _dev.empty((T, nrow, ncol), dev=A.device)
        accumulate = False
    check(lib().limeb200_sos_outer(_dev.ptr(A), TA, _dev.ptr(Bf), TB, T, R, nrow, ncol, float(scale),
                                   1 if accumulate else 0, _dev.ptr(out), _dev.stream_ptr(A.device)))
    return out
