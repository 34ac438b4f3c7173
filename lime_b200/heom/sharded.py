"""
ADO-sharded HEOM propagation over the ranks of a torch.distributed group (BASELINE config 4:
one large hierarchy partitioned by ADO index across the GPUs of one box).

Partition: contiguous blocks of `chunk = ceil(N_he / world)` ADOs in the lexicographic order of
lime's index tables (lime/heom/heom.py:21-108).  Every rank keeps the full stage vector, computes
the RK4 stage (lime/phys.py:636-649) of the ADOs it OWNS with limeb200_heom_stage, and the new
stage vector is exchanged with ONE all-gather per stage (4 per RK4 step) -- with this ordering
the rows a rank needs from its peers are at least as many as the rows it owns, so a
halo exchange would move about the same bytes.  rho and the RK4 accumulator are only ever
touched on the owned rows, so they need no communication; after stage 3 the gathered stage vector
IS rho_{n+1}.

Two data paths:
  exchange='p2p'  (default on CUDA) fused compute + exchange: ONE persistent kernel per rank for the whole
                  run (limeb200_heom_run_sharded); every new stage-vector element is stored locally and, through
                  CUDA-IPC peer pointers, into the stage vectors of the peers that read it over NVLink; the stages are
                  separated by a one-hop barrier (every CTA counts itself on every peer with one remote reduction and
                  polls only its own memory).  torch.distributed is used only to exchange the IPC handles.
  exchange='flow' (default where it applies: diagonal coupling operators) the dataflow kernel of
                  csrc/heom_flow.cuh: NO barrier between stages.  Every stage-vector entry carries a 64-bit stage tag
                  in the same 16-byte words as its value; owners store new entries locally and into the buffers of the
                  ranks that read them, consumers poll exactly the entries they need in their own memory -- one one-way
                  NVLink traversal per stage instead of store acknowledgement + flag + barrier.
  exchange='halo' (opt-in; measured slower than 'p2p', kept as the tested middle ground) hybrid: the barrier kernel with plain
                  16-byte elements and a grid barrier INSIDE each GPU, the tagged halo of the dataflow kernel BETWEEN the
                  GPUs (rows owned by other ranks are polled in a tagged inbox; no cross-GPU barrier).
  exchange='nccl' CUDA stage kernel -> NCCL all-gather, 4 x per step, captured in a CUDA graph.
`stage_fn` is a seam for the world_size-2 gloo tests of this host logic (they plug a CPU stage
function built from the oracle); the product path always uses the CUDA plan.
"""
import numpy as np
import torch
import torch.distributed as dist

from .. import engine
from .. import _dev
from .heom import _calc_matsubara_params


def peer_masks(dn, up, ranges, rank):
    """uint8 [N_he]: bit q of entry a is set when peer slot q (the q-th rank other than `rank`) owns an ADO that
    couples to ADO a, i.e. that peer reads a in every stage.  Only entries of the ADOs `rank` owns are non-zero."""
    nhe = dn.shape[0]
    mask = np.zeros(nhe, dtype=np.uint8)
    lo, hi = ranges[rank]
    q = 0
    for r, (rlo, rhi) in enumerate(ranges):
        if r == rank:
            continue
        nb = np.concatenate([dn[rlo:rhi].reshape(-1), up[rlo:rhi].reshape(-1)])
        nb = np.unique(nb[(nb >= lo) & (nb < hi)])
        mask[nb] |= np.uint8(1 << q)
        q += 1
    return mask


def partition(nhe, world):
    """(chunk, [(lo, hi)] per rank): contiguous, equal chunks (the last one may be short or empty)"""
    chunk = -(-nhe // world)
    return chunk, [(min(nhe, r * chunk), min(nhe, (r + 1) * chunk)) for r in range(world)]


class ShardedHEOM:
    def __init__(self, H, Q, coup_strength, cut_freq, temperature, N_exp=2, N_cut=4,
                 pref_dn=-1j, pref_up=-1j, group=None, stage_fn=None, device=None, use_graph=True,
                 exchange='auto', halo_only=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.H = _dev.as_c128(H)
        self.n = self.H.shape[0]
        Qs = [_dev.as_c128(q) for q in Q] if isinstance(Q, (list, tuple)) else [_dev.as_c128(Q)]
        nbath = len(Qs)
        lam = np.broadcast_to(np.asarray(coup_strength, dtype=float), (nbath,))
        gam = np.broadcast_to(np.asarray(cut_freq, dtype=float), (nbath,))
        c, nu, qmap = [], [], []
        for b in range(nbath):
            cb, nub = _calc_matsubara_params(N_exp, lam[b], gam[b], temperature)
            c += cb
            nu += nub
            qmap += [b] * N_exp
        self.c = np.array(c, dtype=complex)
        self.nu = np.array(nu, dtype=float)
        self.qmap = np.array(qmap, dtype=np.int32)
        self.Q = np.stack(Qs)
        self.nmodes = nbath * N_exp
        self.states, self.dn, self.up = engine.heom_tables([N_cut + 1] * self.nmodes, N_cut)
        self.nhe = self.states.shape[0]
        self.chunk, self.ranges = partition(self.nhe, self.world)
        if exchange in ('p2p', 'flow', 'halo', 'auto') and stage_fn is None:
            # checked identically on every rank BEFORE any IPC set-up, so that all ranks raise together
            if self.world > 8:
                raise ValueError('exchange="p2p" supports at most 8 ranks (one NVSwitch box); use exchange="nccl"')
            if any(hi <= lo for lo, hi in self.ranges):
                raise ValueError('exchange="p2p": %d ADOs leave a rank of %d without rows; use fewer ranks or '
                                 'exchange="nccl"' % (self.nhe, self.world))
        self.nhe_pad = self.chunk * self.world
        self.lo, self.hi = self.ranges[self.rank]
        self.last_launches = 0
        self._stage_fn = stage_fn
        self.use_graph = use_graph
        self.exchange = exchange if stage_fn is None else 'collective'
        self._tag = 1
        self._peer = None
        self._epoch = 0
        # fused path: new stage values travel only to the peers that own a neighbour of the ADO (plus one full
        # broadcast at the end of a run); halo_only=False stores every value into every peer
        self.peer_mask = peer_masks(self.dn, self.up, self.ranges, self.rank) if halo_only else None
        self._d_mask = None
        if stage_fn is None:
            self.dev = _dev.device() if device is None else device
            self.plan = engine.HeomPlan(self.H, self.Q, self.qmap, self.c, self.nu, self.states, self.dn, self.up,
                                        pref_dn=pref_dn, pref_up=pref_up, row_range=(self.lo, self.hi),
                                        device_index=self.dev.index)
            if self.exchange in ('auto', 'flow', 'halo'):
                from .._lib import lib
                # 0: not supported (dense coupling operators ...), 1: tiled variant, 2: register-resident variant.
                # 'auto' takes the dataflow kernel in the regime it is built for (2: a few thousand elements per SM, where
                # the cross-GPU barrier dominates a stage) and the barrier kernel with peer stores for larger shards, which
                # are throughput-bound (measured at 2 ranks, 38 760 ADOs: 9.1e7 tiled dataflow vs 1.27e8 ADO-steps/s)
                level = int(lib().limeb200_heom_flow_supported(self.plan._h))
                if self.world > 1:                      # every rank must take the same path
                    levels = [None] * self.world
                    dist.all_gather_object(levels, level, group=self.group)
                    level = min(levels)
                if self.exchange in ('flow', 'halo') and level < 1:
                    raise ValueError("exchange='%s' needs diagonal coupling operators with at most 4 modes per matrix "
                                     "element; use exchange='p2p'" % self.exchange)
                if self.exchange == 'auto':
                    # ('halo' is never picked automatically: its per-neighbour polling serialises the gather -- 8.2e7 vs
                    #  1.27e8 ADO-steps/s of the barrier kernel at 2 ranks, 38 760 ADOs)
                    self.exchange = 'flow' if (level == 2 and self.world > 1) else 'p2p'
        else:
            self.dev = torch.device('cpu') if device is None else device
            self.plan = None

    # ---- one RK4 stage of the owned rows, then the exchange --------------------------------
    def _stage(self, stage, rho, yin, ynext, acc, dt):
        if self._stage_fn is not None:
            self._stage_fn(self, stage, rho, yin, ynext, acc, dt)
        else:
            self.plan.stage(stage, rho, yin, ynext, acc, dt)
            self.last_launches += 1

    def _exchange(self, y):
        """all-gather of the stage vector y [1, nhe_pad, n, n]: every rank contributes its chunk"""
        if self.world == 1:
            return
        flat = y.view(self.world, -1)
        if dist.get_backend(self.group) == 'nccl':
            dist.all_gather_into_tensor(y.view(-1), flat[self.rank], group=self.group)     # in place
        else:
            chunks = [torch.empty_like(flat[0]) for _ in range(self.world)]
            dist.all_gather(chunks, flat[self.rank].clone(), group=self.group)
            for r in range(self.world):
                flat[r].copy_(chunks[r])
        self.last_launches += 1

    # ---- fused compute + exchange over peer memory ---------------------------------------------
    def _peer_setup(self):
        import ctypes as C
        from .._lib import lib, check
        nbytes = self.nhe_pad * self.n * self.n * 16
        mine, handles = [], []
        flow = self.exchange in ('flow', 'halo')
        # flow: two TAGGED stage vectors (32 bytes per element) and no flag array (third buffer unused)
        for size in ((2 * nbytes, 2 * nbytes, 256) if flow else (nbytes, nbytes, 256)):
            ptr = C.c_void_p()
            h = (C.c_ubyte * 64)()
            check(lib().limeb200_peer_alloc(self.dev.index, size, C.byref(ptr), h))
            mine.append(ptr.value)
            handles.append(bytes(h))
        allh = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(allh, handles, group=self.group)
        else:
            allh[0] = handles
        ptrs = [[0] * self.world for _ in range(3)]
        opened = []
        for r in range(self.world):
            for k in range(3):
                if r == self.rank:
                    ptrs[k][r] = mine[k]
                else:
                    q = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(allh[r][k])
                    check(lib().limeb200_peer_open(self.dev.index, hb, C.byref(q)))
                    ptrs[k][r] = q.value
                    opened.append(q.value)
        arr = [(C.c_void_p * self.world)(*ptrs[k]) for k in range(3)]
        # arrivals per stage of every rank = CTAs of its persistent kernel (the one-hop barrier counts them)
        g = 1 if flow else check(lib().limeb200_heom_persist_grid(self.plan._h, 1))
        grids = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(grids, int(g), group=self.group)
        else:
            grids[0] = int(g)
        self._peer = dict(mine=mine, ptrs=ptrs, arr=arr, opened=opened, nbytes=nbytes,
                          grids=(C.c_int * self.world)(*grids))

    def close(self):
        if self._peer is not None:
            from .._lib import lib
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)
            for q in self._peer['opened']:
                lib().limeb200_peer_close(self.dev.index, q)
            if self.world > 1:
                dist.barrier(group=self.group)
            for m in self._peer['mine']:
                lib().limeb200_peer_free(self.dev.index, m)
            self._peer = None

    def _run_p2p(self, ado, dt, nsteps):
        import ctypes as C
        from .._lib import lib, check
        if self._peer is None:
            self._peer_setup()
        pr = self._peer
        rho = torch.zeros((1, self.nhe_pad, self.n, self.n), dtype=torch.complex128, device=ado.device)
        rho[:, :self.nhe] = ado
        st = torch.cuda.current_stream()
        # every rank's y0 <- the full state; peers may only start storing into our buffers once we are ready
        sp = C.c_void_p(st.cuda_stream)
        check(lib().limeb200_memcpy_d2d(C.c_void_p(pr['mine'][0]), C.c_void_p(rho.data_ptr()), pr['nbytes'], sp))
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.peer_mask is not None and self._d_mask is None:
            self._d_mask = torch.from_numpy(self.peer_mask).to(ado.device)
        mptr = C.c_void_p(self._d_mask.data_ptr()) if self._d_mask is not None else None
        check(lib().limeb200_heom_run_sharded(self.plan._h, self.rank, self.world, pr['arr'][0], pr['arr'][1],
                                              pr['arr'][2], pr['grids'], C.c_void_p(rho.data_ptr()), mptr, float(dt), int(nsteps),
                                              C.c_uint(self._epoch), C.c_void_p(st.cuda_stream)))
        self._epoch += 4 * nsteps
        err = lib().limeb200_heom_sharded_error(self.plan._h, C.c_void_p(st.cuda_stream))
        if err:
            from .._lib import LimeB200Error
            raise LimeB200Error('sharded HEOM run: a peer GPU never reached the stage barrier')
        check(lib().limeb200_memcpy_d2d(C.c_void_p(rho.data_ptr()), C.c_void_p(pr['mine'][0]), pr['nbytes'], sp))
        ado.copy_(rho[:, :self.nhe])
        torch.cuda.synchronize()
        self.last_launches = 1
        if self.world > 1:
            dist.barrier(group=self.group)       # nobody reuses the buffers before everyone has copied out
        return ado

    def _run_flow(self, ado, dt, nsteps):
        """dataflow kernel: pack the full state into the tagged buffer, synchronise the ranks, ONE launch, synchronise,
        unpack (the last stage delivered every owner's rows to every rank)"""
        import ctypes as C
        from .._lib import lib, check, LimeB200Error
        if self._peer is None:
            self._peer_setup()
        pr = self._peer
        rho = torch.zeros((1, self.nhe_pad, self.n, self.n), dtype=torch.complex128, device=ado.device)
        rho[:, :self.nhe] = ado
        st = torch.cuda.current_stream(ado.device)
        sp = C.c_void_p(st.cuda_stream)
        tag0 = self._tag
        self._tag += 4 * nsteps + 4
        check(lib().limeb200_heom_flow_pack(self.plan._h, C.c_void_p(rho.data_ptr()), C.c_void_p(pr['mine'][0]),
                                            C.c_ulonglong(tag0), sp))
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)        # every rank's buffer holds the tagged state before anyone stores into it
        if self.peer_mask is not None and self._d_mask is None:
            self._d_mask = torch.from_numpy(self.peer_mask).to(ado.device)
        mptr = C.c_void_p(self._d_mask.data_ptr()) if self._d_mask is not None else None
        run = lib().limeb200_heom_flow_run_sharded if self.exchange == 'flow' else lib().limeb200_heom_run_sharded_halo
        check(run(self.plan._h, self.rank, self.world, pr['arr'][0], pr['arr'][1], C.c_void_p(rho.data_ptr()), mptr,
                  float(dt), int(nsteps), C.c_ulonglong(tag0), sp))
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)        # all peers' last-stage stores have landed
        check(lib().limeb200_heom_flow_unpack(self.plan._h, C.c_void_p(pr['mine'][0]), C.c_ulonglong(tag0 + 4 * nsteps),
                                              C.c_void_p(rho.data_ptr()), sp))
        err = lib().limeb200_heom_sharded_error(self.plan._h, sp)
        if err:
            raise LimeB200Error('sharded HEOM run (dataflow kernel): %s' %
                                ('a value never arrived (wait timed out)' if err & 1 else 'stale entries in the final state'))
        ado.copy_(rho[:, :self.nhe])
        torch.cuda.synchronize()
        self.last_launches = 1
        if self.world > 1:
            dist.barrier(group=self.group)
        return ado

    def run_device(self, ado, dt, nsteps):
        """ado: [1, N_he, n, n] complex128 on self.dev, identical on every rank; advanced in place by
        nsteps RK4 steps (every rank ends up with the full hierarchy)."""
        assert ado.shape == (1, self.nhe, self.n, self.n) and ado.dtype == torch.complex128
        if self.exchange in ('flow', 'halo') and self._stage_fn is None and nsteps > 0:
            return self._run_flow(ado, dt, nsteps)
        if self.exchange == 'p2p' and self._stage_fn is None and nsteps > 0:
            return self._run_p2p(ado, dt, nsteps)
        shape = (1, self.nhe_pad, self.n, self.n)
        y = [torch.zeros(shape, dtype=torch.complex128, device=ado.device) for _ in range(2)]
        acc = torch.zeros(shape, dtype=torch.complex128, device=ado.device)
        rho = torch.zeros(shape, dtype=torch.complex128, device=ado.device)
        rho[:, :self.nhe] = ado
        y[0].copy_(rho)
        self.last_launches = 0

        def one_step():
            for stage in range(4):
                yin, ynext = y[stage & 1], y[(stage + 1) & 1]
                self._stage(stage, rho, yin, ynext, acc, dt)
                self._exchange(ynext)

        if self.use_graph and ado.is_cuda and nsteps > 2:
            # one RK4 step (4 stage kernels + 4 all-gathers) captured once, replayed nsteps times: the
            # per-stage host cost (ctypes call + collective launch) would otherwise dominate
            one_step()                                   # warm-up outside capture (NCCL channel set-up)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                one_step()
            per_step = self.last_launches // 2
            for _ in range(nsteps - 1):
                g.replay()
            self.last_launches = per_step * nsteps
        else:
            for _ in range(nsteps):
                one_step()
        ado.copy_(y[0][:, :self.nhe])        # after stage 3 the gathered stage vector is rho_{n+1}
        return ado

    def initial(self, rho0):
        a = np.zeros((1, self.nhe, self.n, self.n), dtype=np.complex128)
        a[0, 0] = rho0
        return a

    def evolve(self, rho0, dt, Nt):
        """host in / host out; returns a lime Result whose rholist holds the final reduced density matrix and
        `ado` the final hierarchy"""
        from ..mol import Result
        d = torch.from_numpy(self.initial(rho0)).to(self.dev)
        self.run_device(d, dt, Nt)
        out = d.cpu().numpy()[0]
        res = Result(dt=dt, Nt=Nt, rho0=rho0)
        res.observables = np.zeros((Nt, 0), dtype=complex)
        res.rholist = [out[0]]
        res.ado = out
        return res
