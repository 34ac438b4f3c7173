"""ADO-sharded HEOM (lime_b200/heom/sharded.py): host logic under gloo with world_size 2 on CPU
(the CUDA stage kernel is replaced by an oracle-based stage function through the `stage_fn` seam),
and the real CUDA + NCCL path when at least two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import lime_oracle as lo
from conftest import relerr          # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(depth=3):
    n = 3
    H = cases.rand_herm(n, 1)
    Q = [np.diag([1.0, -0.5, 0.25]), np.diag([0.0, 1.0, 0.0])]
    rho0 = cases.rand_dm(n, 4)
    return H, Q, [0.1, 0.2], [0.8, 1.3], 1.5, depth, rho0


def _oracle_stage(sh, stage, rho, yin, ynext, acc, dt):
    """CPU stand-in for limeb200_heom_stage: RK4 stage algebra of lime/phys.py:636-649 on the owned rows"""
    st = sh.states.astype(np.int64)
    y = yin[0, :sh.nhe].numpy()
    k = lo.heom_rhs(y, sh.H, sh.Q, sh.qmap, sh.c, sh.nu, st, sh.dn.astype(np.int64), sh.up.astype(np.int64))
    sl = slice(sh.lo, sh.hi)
    r, a, out = rho[0].numpy(), acc[0].numpy(), ynext[0].numpy()
    if stage == 0:
        a[sl] = k[sl]
        out[sl] = r[sl] + 0.5 * dt * k[sl]
    elif stage == 1:
        a[sl] += 2 * k[sl]
        out[sl] = r[sl] + 0.5 * dt * k[sl]
    elif stage == 2:
        a[sl] += 2 * k[sl]
        out[sl] = r[sl] + dt * k[sl]
    else:
        r[sl] += (a[sl] + k[sl]) / 6.0 * dt
        out[sl] = r[sl]


def _worker(rank, world, port, backend, q, exchange='p2p'):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from lime_b200.heom.sharded import ShardedHEOM
        H, Q, lam, gam, T, depth, rho0 = _problem()
        if backend == 'nccl':
            torch.cuda.set_device(rank)
            sh = ShardedHEOM(H, Q, lam, gam, T, N_exp=2, N_cut=depth, exchange=exchange.replace('_all', ''),
                             halo_only=(exchange != 'p2p_all'))
        else:
            sh = ShardedHEOM(H, Q, lam, gam, T, N_exp=2, N_cut=depth, stage_fn=_oracle_stage)
        ado = torch.from_numpy(sh.initial(rho0)).to(sh.dev)
        sh.run_device(ado, 0.01, 5)
        sh.run_device(ado, 0.01, 7)              # second call: flags / epochs continue
        q.put((rank, ado.cpu().numpy()[0], sh.ranges, sh.nhe))
        if backend == 'nccl':
            sh.close()
    finally:
        dist.destroy_process_group()


def _run(world, backend, exchange='p2p'):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


def _reference():
    from lime_b200.heom.heom import _calc_matsubara_params
    H, Q, lam, gam, T, depth, rho0 = _problem()
    c, nu, qmap = [], [], []
    for b in range(2):
        cb, nub = _calc_matsubara_params(2, lam[b], gam[b], T)
        c += cb
        nu += nub
        qmap += [b, b]
    states, dn, up = lo.heom_tables([depth + 1] * 4, depth)
    ado0 = np.zeros((states.shape[0], 3, 3), dtype=complex)
    ado0[0] = rho0
    out, _, _ = lo.heom_rk4(ado0, H, np.stack(Q).astype(complex), qmap, np.array(c), np.array(nu), states, dn, up,
                            0.01, 12)
    return out


def test_partition():
    from lime_b200.heom.sharded import partition
    chunk, r = partition(35, 2)
    assert chunk == 18 and r == [(0, 18), (18, 35)]
    chunk, r = partition(3060, 8)
    assert chunk == 383 and r[-1] == (2681, 3060) and sum(h - l for l, h in r) == 3060
    chunk, r = partition(3, 8)
    assert chunk == 1 and r[3] == (3, 3)


def test_peer_masks_cover_exactly_the_remote_neighbours():
    from lime_b200 import engine
    from lime_b200.heom.sharded import partition, peer_masks
    states, dn, up = engine.heom_tables([5] * 4, 4)
    nhe = states.shape[0]
    for world in (2, 3, 8):
        chunk, ranges = partition(nhe, world)
        masks = [peer_masks(dn, up, ranges, r) for r in range(world)]
        for me in range(world):
            lo, hi = ranges[me]
            assert not masks[me][:lo].any() and not masks[me][hi:].any()        # only owned ADOs are sent
            others = [r for r in range(world) if r != me]
            for q, r in enumerate(others):
                rlo, rhi = ranges[r]
                need = set(np.concatenate([dn[rlo:rhi].ravel(), up[rlo:rhi].ravel()]).tolist()) - {-1}
                need = {a for a in need if lo <= a < hi}
                got = set(np.nonzero((masks[me] >> q) & 1)[0].tolist())
                assert got == need, (world, me, r)
    # the point of the exercise: far fewer stores than "everything to everyone" at 8 ranks
    chunk, ranges = partition(nhe, 8)
    sent = sum(int(np.unpackbits(peer_masks(dn, up, ranges, r)).sum()) for r in range(8))
    assert sent < 0.6 * 7 * nhe


def test_sharded_heom_gloo_world2():
    res = _run(2, 'gloo')
    ref = _reference()
    assert res[0][3] == ref.shape[0] == 35
    for rank, ado, ranges, nhe in res:
        assert ranges == [(0, 18), (18, 35)]
        assert relerr(ado, ref) <= 1e-12, rank
    assert np.array_equal(res[0][1], res[1][1])          # every rank ends with the same full hierarchy


def test_sharded_heom_gloo_world3_uneven():
    """35 ADOs over 3 ranks: chunks of 12 with a short last one (padding rows are exchanged but never owned)"""
    res = _run(3, 'gloo')
    ref = _reference()
    for rank, ado, ranges, nhe in res:
        assert ranges == [(0, 12), (12, 24), (24, 35)]
        assert relerr(ado, ref) <= 1e-12, rank


@pytest.mark.gpu
def test_sharded_heom_single_rank_persistent_kernel(cuda):
    """world = 1: the sharded entry point (persistent kernel + peer-allocated buffers) on one GPU"""
    from lime_b200.heom.sharded import ShardedHEOM
    H, Q, lam, gam, T, depth, rho0 = _problem()
    sh = ShardedHEOM(H, Q, lam, gam, T, N_exp=2, N_cut=depth)
    ado = torch.from_numpy(sh.initial(rho0)).to(sh.dev)
    sh.run_device(ado, 0.01, 5)
    sh.run_device(ado, 0.01, 7)
    assert relerr(ado.cpu().numpy()[0], _reference()) <= 1e-10
    sh.close()


@pytest.mark.gpu
@pytest.mark.parametrize('exchange', ['flow', 'halo', 'p2p', 'p2p_all', 'nccl'])
def test_sharded_heom_nccl_world2(exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    res = _run(2, 'nccl', exchange)
    ref = _reference()
    for rank, ado, ranges, nhe in res:
        assert relerr(ado, ref) <= 1e-10, rank
    assert np.array_equal(res[0][1], res[1][1])


def _virtual_ranks_one_gpu(world, H, Q, lam, gam, T, depth, ado0, dt, runs, halo_only=True, flow=False, halo=False):
    """`world` ranks of the fused sharded kernel as `world` plans + streams on ONE GPU in one process: every rank's
    persistent kernel is resident at the same time (the grids are small), "peer" stores and the one-hop barrier go
    through ordinary device pointers instead of CUDA-IPC mappings.  Exercises limeb200_heom_run_sharded's cross-rank
    protocol (peer stores, remote arrival counts, epochs across launches) on a single-GPU box."""
    import ctypes as C
    from lime_b200 import engine
    from lime_b200._lib import lib, check
    from lime_b200.heom.heom import _calc_matsubara_params
    from lime_b200.heom.sharded import partition, peer_masks
    dev = torch.device('cuda', 0)
    n = H.shape[0]
    c, nu, qmap = [], [], []
    for b in range(len(Q)):
        cb, nub = _calc_matsubara_params(2, np.broadcast_to(lam, (len(Q),))[b], np.broadcast_to(gam, (len(Q),))[b], T)
        c += cb
        nu += nub
        qmap += [b, b]
    states, dn, up = engine.heom_tables([depth + 1] * len(qmap), depth)
    nhe = states.shape[0]
    chunk, ranges = partition(nhe, world)
    nhe_pad = chunk * world
    plans = [engine.HeomPlan(H, np.stack(Q).astype(complex), qmap, np.array(c), np.array(nu), states, dn, up,
                             row_range=ranges[r], device_index=0) for r in range(world)]
    full = torch.zeros((1, nhe_pad, n, n), dtype=torch.complex128, device=dev)
    full[0, :nhe] = torch.from_numpy(ado0).to(dev)
    if flow or halo:
        return _virtual_flow(world, plans, full, masks_of(dn, up, ranges, world, halo_only, dev), nhe, dt, runs, halo=halo), ranges
    y0 = [full.clone() for _ in range(world)]
    y1 = [torch.zeros_like(full) for _ in range(world)]
    rho = [full.clone() for _ in range(world)]
    flags = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(world)]
    masks = [torch.from_numpy(peer_masks(dn, up, ranges, r)).to(dev) if halo_only else None for r in range(world)]
    grids = (C.c_int * world)(*[check(lib().limeb200_heom_persist_grid(p._h, 1)) for p in plans])
    arr = [(C.c_void_p * world)(*[t.data_ptr() for t in ts]) for ts in (y0, y1, flags)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    epoch = 0
    for nsteps in runs:
        for r in range(world):
            check(lib().limeb200_heom_run_sharded(plans[r]._h, r, world, arr[0], arr[1], arr[2], grids,
                                                  C.c_void_p(rho[r].data_ptr()),
                                                  C.c_void_p(masks[r].data_ptr()) if halo_only else None,
                                                  float(dt), int(nsteps), C.c_uint(epoch),
                                                  C.c_void_p(streams[r].cuda_stream)))
        torch.cuda.synchronize()
        for r in range(world):
            assert lib().limeb200_heom_sharded_error(plans[r]._h, C.c_void_p(streams[r].cuda_stream)) == 0, \
                'rank %d: a peer never reached the stage barrier' % r
        epoch += 4 * nsteps
    return [y[0, :nhe].cpu().numpy() for y in y0], ranges


def masks_of(dn, up, ranges, world, halo_only, dev):
    from lime_b200.heom.sharded import peer_masks
    return [torch.from_numpy(peer_masks(dn, up, ranges, r)).to(dev) if halo_only else None for r in range(world)]


def _virtual_flow(world, plans, full, masks, nhe, dt, runs, halo=False):
    """the dataflow kernel (tagged stage vectors, no barrier) with `world` ranks resident together on one GPU"""
    import ctypes as C
    from lime_b200._lib import lib, check
    dev = full.device
    nwords = full.numel() * 4                                   # 32 bytes per complex element
    T0 = [torch.zeros(nwords, dtype=torch.int64, device=dev) for _ in range(world)]
    T1 = [torch.zeros(nwords, dtype=torch.int64, device=dev) for _ in range(world)]
    rho = [full.clone() for _ in range(world)]
    arr = [(C.c_void_p * world)(*[t.data_ptr() for t in ts]) for ts in (T0, T1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    tag = 1
    state = full.clone()
    for nsteps in runs:
        for r in range(world):
            assert lib().limeb200_heom_flow_supported(plans[r]._h) >= 1
            check(lib().limeb200_heom_flow_pack(plans[r]._h, C.c_void_p(state.data_ptr()), C.c_void_p(T0[r].data_ptr()),
                                                C.c_ulonglong(tag), None))
            rho[r].copy_(state)
        torch.cuda.synchronize()
        run = lib().limeb200_heom_run_sharded_halo if halo else lib().limeb200_heom_flow_run_sharded
        for r in range(world):
            check(run(plans[r]._h, r, world, arr[0], arr[1], C.c_void_p(rho[r].data_ptr()),
                      C.c_void_p(masks[r].data_ptr()) if masks[r] is not None else None,
                      float(dt), int(nsteps), C.c_ulonglong(tag), C.c_void_p(streams[r].cuda_stream)))
        torch.cuda.synchronize()
        outs = []
        for r in range(world):
            o = torch.zeros_like(full)
            check(lib().limeb200_heom_flow_unpack(plans[r]._h, C.c_void_p(T0[r].data_ptr()), C.c_ulonglong(tag + 4 * nsteps),
                                                  C.c_void_p(o.data_ptr()), None))
            torch.cuda.synchronize()
            assert lib().limeb200_heom_sharded_error(plans[r]._h, None) == 0, 'rank %d: timed-out wait or stale tag' % r
            outs.append(o)
        state = outs[0].clone()
        tag += 4 * nsteps + 4
    return [o[0, :nhe].cpu().numpy() for o in outs]


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 3, 4])
@pytest.mark.parametrize('halo_only', [True, False])
def test_sharded_dataflow_virtual_ranks_on_one_gpu(cuda, world, halo_only):
    """the dataflow sharded kernel (tagged entries, peer stores, no barrier), 2-4 ranks resident together on one GPU,
    two consecutive runs (tags continue)"""
    H, Q, lam, gam, T, depth, rho0 = _problem()
    ref = _reference()
    ado0 = np.zeros_like(ref)
    ado0[0] = rho0
    outs, ranges = _virtual_ranks_one_gpu(world, H, Q, lam, gam, T, depth, ado0, 0.01, [5, 7], halo_only=halo_only, flow=True)
    for r, o in enumerate(outs):
        assert relerr(o, ref) <= 1e-10, r
        assert np.array_equal(o, outs[0])


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 3])
@pytest.mark.parametrize('halo_only', [True, False])
def test_sharded_tagged_halo_virtual_ranks_on_one_gpu(cuda, world, halo_only):
    """hybrid sharded propagator (grid barrier inside a rank, tagged halo between ranks), virtual ranks on one GPU"""
    H, Q, lam, gam, T, depth, rho0 = _problem()
    ref = _reference()
    ado0 = np.zeros_like(ref)
    ado0[0] = rho0
    outs, ranges = _virtual_ranks_one_gpu(world, H, Q, lam, gam, T, depth, ado0, 0.01, [5, 7], halo_only=halo_only, halo=True)
    for r, o in enumerate(outs):
        assert relerr(o, ref) <= 1e-10, r
        assert np.array_equal(o, outs[0])


@pytest.mark.gpu
def test_sharded_tagged_halo_fmo_depth3_long(cuda):
    """hybrid propagator, FMO depth 3 over 3 virtual ranks, 200 steps, against the one-GPU barrier kernel"""
    from lime_b200 import builders
    from lime_b200.heom.heom import HEOM
    from lime_b200.units import au2fs
    Hm, Q, lam, gam, kT = builders.fmo_heom_inputs()
    h = HEOM(Hm, Q, lam, gam, kT, N_exp=2, N_cut=3)
    h.plan.set_path(3)
    rho0 = np.zeros((7, 7), dtype=complex)
    rho0[0, 0] = 1.0
    dt = 0.5 / au2fs
    one, _, _ = h.plan.run(h.initial(rho0), dt, 200)
    outs, _ = _virtual_ranks_one_gpu(3, Hm, Q, lam, gam, kT, 3, h.initial(rho0), dt, [200], halo=True)
    for o in outs:
        assert relerr(o, one) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('world', [1, 4])
def test_sharded_dataflow_fmo_depth3_long(cuda, world):
    """FMO depth 3 (680 ADOs of 7x7), 200 steps = 800 unsynchronised stages, against the barrier kernel of the same
    library on one GPU"""
    from lime_b200 import builders
    from lime_b200.heom.heom import HEOM
    from lime_b200.units import au2fs
    Hm, Q, lam, gam, kT = builders.fmo_heom_inputs()
    h = HEOM(Hm, Q, lam, gam, kT, N_exp=2, N_cut=3)
    h.plan.set_path(3)
    rho0 = np.zeros((7, 7), dtype=complex)
    rho0[0, 0] = 1.0
    dt = 0.5 / au2fs
    one, _, _ = h.plan.run(h.initial(rho0), dt, 200)
    assert h.plan.path == 3
    outs, _ = _virtual_ranks_one_gpu(world, Hm, Q, lam, gam, kT, 3, h.initial(rho0), dt, [200], flow=True)
    for o in outs:
        assert relerr(o, one) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 3, 4])
@pytest.mark.parametrize('halo_only', [True, False])
def test_sharded_virtual_ranks_on_one_gpu(cuda, world, halo_only):
    """the cross-rank protocol of the fused sharded kernel (peer stores, one-hop barrier with remote arrival counts,
    epochs continuing over two launches), 2-4 ranks resident together on a single GPU"""
    H, Q, lam, gam, T, depth, rho0 = _problem()
    ref = _reference()
    ado0 = np.zeros_like(ref)
    ado0[0] = rho0
    outs, ranges = _virtual_ranks_one_gpu(world, H, Q, lam, gam, T, depth, ado0, 0.01, [5, 7], halo_only=halo_only)
    for r, o in enumerate(outs):
        assert relerr(o, ref) <= 1e-10, r
        assert np.array_equal(o, outs[0])            # every rank ends with the same full hierarchy


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['0', '1'])
def test_sharded_virtual_ranks_generic_persistent_kernel(cuda, monkeypatch, mode):
    """the GENERIC persistent kernel (what the barrier path runs on shards too large for the cached one; forced here with
    LIMEB200_HEOM_NO_CACHED) across 3 virtual ranks: table walk (mode 0, the default there) and packed neighbour lists
    (mode 1, opt-in: measured slower in this register-limited kernel) behind the same peer stores and one-hop barrier"""
    monkeypatch.setenv('LIMEB200_HEOM_NO_CACHED', '1')
    monkeypatch.setenv('LIMEB200_HEOM_STAGE_MODE', mode)
    H, Q, lam, gam, T, depth, rho0 = _problem()
    ref = _reference()
    ado0 = np.zeros_like(ref)
    ado0[0] = rho0
    outs, ranges = _virtual_ranks_one_gpu(3, H, Q, lam, gam, T, depth, ado0, 0.01, [5, 7], halo_only=True)
    for r, o in enumerate(outs):
        assert relerr(o, ref) <= 1e-10, r
        assert np.array_equal(o, outs[0])


@pytest.mark.gpu
def test_sharded_virtual_ranks_fmo_depth3_long(cuda):
    """a larger hierarchy (FMO, depth 3: 680 ADOs of 7x7) over 4 virtual ranks for 200 steps against the one-GPU
    propagator of the same library: the barrier protocol over 800 stages"""
    from lime_b200 import builders
    from lime_b200.heom.heom import HEOM
    from lime_b200.units import au2fs
    Hm, Q, lam, gam, kT = builders.fmo_heom_inputs()
    h = HEOM(Hm, Q, lam, gam, kT, N_exp=2, N_cut=3)
    rho0 = np.zeros((7, 7), dtype=complex)
    rho0[0, 0] = 1.0
    dt = 0.5 / au2fs
    one, _, _ = h.plan.run(h.initial(rho0), dt, 200)
    outs, _ = _virtual_ranks_one_gpu(4, Hm, Q, lam, gam, kT, 3, h.initial(rho0), dt, [200])
    for o in outs:
        assert relerr(o, one) <= 1e-12
