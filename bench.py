#!/usr/bin/env python
"""
bench.py -- throughput of the density-matrix hot path on B200 (BASELINE.json metric:
Lindblad rho-steps/s and HEOM ADO-steps/s; % of the HBM roof).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

One bench "step" = one pass of the hot path over one batch of synthetic input = ONE launch of
the fused propagator that advances every density matrix (or hierarchy) of the batch by
`--rk-steps` RK4 steps, observables included.  Workloads:

    jc_lindblad   (default) config 2: Jaynes-Cummings, Fock cutoff 64 (N = 128), 4096
                  coupling x detuning points PER GPU, c_ops = [sqrt(kappa) a], e_ops = [a^dag a, s+s-]
    heom_fmo      config 4: FMO 7 sites, 7 Drude baths x 2 exponentials, depth 4 (3060 ADOs);
                  ADO-sharded over the ranks with one exchange per RK4 stage (strong scaling)
    heom_sb       config 3 throughput variant: spin-boson, K = 2, depth 12 (91 ADOs), batch of
                  hierarchies differing in (lambda, beta)
    sos_2des      config 5: photon-echo (GSB+SE+ESA) 256 x 256 grid x 64 waiting times, N = 32
    redfield_batch config 1 throughput variant: the Redfield tensor of examples/redfield.py, 2^20 initial states

The JSON line carries `value` (inputs resident in HBM), `e2e` (same metric through the
lime-compatible public API with pinned HOST buffers, H2D/D2H inside the timed region),
`roofline` (algorithmic bytes / CUDA-event duration of the dominant kernel against
MEASURED_PEAKS.json) and `cpu_baseline` (the NumPy/SciPy port of lime's CPU path, oracle/,
timed on the host cores on a bounded sample).  `--impl reference` times that CPU port alone.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT,):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ---------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_id):
        self.gpu_id = gpu_id
        self.proc = None
        self.path = '/tmp/limeb200_clocks_%d_%s.csv' % (os.getpid(), str(gpu_id).replace('/', '_'))

    def start(self):
        try:
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_id), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(',')]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
                except ValueError:
                    continue
                for n, v in zip(names, c[5:9]):
                    if v.lower() == 'active':
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------
def _jc_oracle_point(job):
    """(final rho, observables[steps, 2]) of one Jaynes-Cummings point with lime's algorithm (checker of JCLindblad.check)"""
    omega0, omegac, g, ncav, kappa, dt, nsteps = job
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import lime_oracle as lo
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(4)
    except Exception:
        pass
    H, c_ops, e_ops = lo.jaynes_cummings(omega0, omegac, g, ncav, kappa)
    N = 2 * ncav
    rho0 = np.zeros((N, N), dtype=complex)
    rho0[ncav, ncav] = 1.0
    obs, rl = lo.lindblad(H.toarray(), rho0, [c.toarray() for c in c_ops], [e.toarray() for e in e_ops], Nt=nsteps, dt=dt)
    return rl[-1], obs


class JCLindblad:
    """config 2: batched Jaynes-Cummings Lindblad RK4, N = 2 x 64."""
    name = 'jc_lindblad'
    metric = 'lindblad_rho_steps_per_s'
    unit = 'rho-steps/s'
    dtype = 'f64'                      # complex128 = interleaved f64 pairs
    scaling = 'weak'
    bound = 'hbm'

    def __init__(self, args, rank, world, need_gpu=True):
        from lime_b200 import builders as models
        self.ncav, self.ng, self.ndet = 64, 64, 64
        if args.batch:
            self.ng = max(1, args.batch // self.ndet) if args.batch >= self.ndet else 1
            self.ndet = min(self.ndet, args.batch)
        self.N = 2 * self.ncav
        self.B = self.ng * self.ndet
        self.rk = args.rk_steps or 1000
        self.dt, self.kappa, self.omega0 = 0.01, 0.05, 1.0
        self.rank, self.world = rank, world
        # weak scaling: the detuning axis is refined to 64*world points, rank r takes block r
        gs = np.linspace(0.01, 0.2, self.ng) * self.omega0
        dets = np.linspace(-0.2, 0.2, self.ndet * world)[rank * self.ndet:(rank + 1) * self.ndet] * self.omega0
        G, Dt = np.meshgrid(gs, dets, indexing='ij')
        self.g_pts, self.det_pts = G.reshape(-1), Dt.reshape(-1)
        self.pat, self.vals, self.c_ops, self.e_ops, self.rho0 = models.jaynes_cummings_batch(
            self.omega0, self.omega0 + self.det_pts, self.g_pts, self.ncav, self.kappa)
        self.alg_bytes_per_unit = 2 * 16 * self.N * self.N          # read rho_n, write rho_{n+1}
        self.units_per_step = self.B * self.rk
        self.launches = 0
        self.spot_check = not getattr(args, 'no_spot_check', False) and rank == 0       # rank 0's block of points
        if need_gpu:
            import torch
            from lime_b200 import oqs
            self.torch = torch
            self.plan, _ = oqs._lindblad_plan_batch((self.pat, self.vals), self.c_ops, self.e_ops)
            self.rho = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(self.rho0, (self.B, self.N, self.N)))).cuda()
            self.kernel = {1: 'qme_dense_onchip', 2: 'qme_dense_stage', 3: 'qme_ell_global',
                           4: 'qme_ell_cluster', 5: 'qme_band_kernel', 6: 'qme_tile_kernel'}[self.plan.path]

    def config(self):
        return {'workload': 'jc_lindblad: Jaynes-Cummings (no RWA) Fock cutoff 64 -> N=128, '
                            '%d coupling x %d detuning points per GPU, kappa=%g, dt=%g, %d RK4 steps per launch, '
                            'E=2 observables per step' % (self.ng, self.ndet, self.kappa, self.dt, self.rk),
                'N': self.N, 'batch_per_gpu': self.B, 'rk4_steps_per_launch': self.rk,
                'state_bytes_per_gpu': self.B * self.N * self.N * 16,
                'l2_policy': 'inputs larger than L2 (state %.2f GiB per GPU)' % (self.B * self.N * self.N * 16 / 2 ** 30),
                'state': 'dynamics continue across bench steps (no re-initialisation, no work skipped)',
                'sharding': 'parameter points split across ranks, no collective'}

    def step(self):
        self.obs, _ = self.plan.run_device(self.rho, self.dt, self.rk)
        self.launches += self.plan.last_launches

    def check(self):
        tr = self.torch.einsum('bii->b', self.rho).cpu().numpy()
        err = float(np.max(np.abs(tr - 1.0)))
        assert err < 1e-9, 'trace drifted: %g' % err
        out = {'max_trace_error': err}
        if self.spot_check:
            # oracle spot check of the launch the bench times: the FULL batch from rho0 for `rk` RK4 steps in one launch,
            # three of its points against lime's algorithm (oracle port, dense np.dot route) on the host
            import multiprocessing as mp
            t = self.torch
            rho = t.from_numpy(np.ascontiguousarray(np.broadcast_to(self.rho0, (self.B, self.N, self.N)))).cuda()
            obs, _ = self.plan.run_device(rho, self.dt, self.rk)
            pts = sorted({0, self.B // 2, self.B - 1})
            got = [(rho[i].cpu().numpy(), obs[:, i].cpu().numpy()) for i in pts]
            with mp.get_context('spawn').Pool(len(pts)) as pool:
                ref = pool.map(_jc_oracle_point, [(self.omega0, self.omega0 + self.det_pts[i], self.g_pts[i], self.ncav,
                                                   self.kappa, self.dt, self.rk) for i in pts])
            worst = 0.0
            for (r_g, o_g), (r_o, o_o) in zip(got, ref):
                worst = max(worst, float(np.max(np.abs(r_g - r_o)) / np.max(np.abs(r_o))),
                            float(np.max(np.abs(o_g - o_o)) / np.max(np.abs(o_o))))
            assert worst <= 1e-10, 'bench launch differs from the oracle: %g' % worst
            out.update(oracle_spot_check_relerr=worst, oracle_spot_check_points=pts, oracle_spot_check_rk4_steps=self.rk,
                       oracle_spot_check_what='final rho and the (steps, 2) observables of these batch points, '
                                              'max|a-b|/max|b|, vs oracle/lime_oracle.lindblad on the host')
            del rho, obs
        return out

    def e2e_setup(self):
        from lime_b200.oqs import Lindblad_solver
        t = self.torch
        self.h_rho = t.from_numpy(np.ascontiguousarray(np.broadcast_to(self.rho0, (self.B, self.N, self.N)))).pin_memory()
        self.solver = Lindblad_solver(None, c_ops=self.c_ops)
        self.h_batch = (self.pat, self.vals)

    def e2e_step(self):
        """public API, host buffers in and out: lime's Lindblad_solver with the batch extension.  The solver object
        is reused, so after the first (warm-up) call the operator plan is cached and a call costs the pinned H2D of
        the states, the launch and the pinned D2H of states + observables"""
        rho_f, obs, _ = self.solver.evolve_batch(self.h_rho, self.dt, self.rk, e_ops=self.e_ops,
                                                 H_batch=self.h_batch, pinned=True)
        self.launches_e2e = 1
        h2d = self.h_rho.numel() * 16          # the operator values travel once, in the warm-up call that builds the plan
        d2h = rho_f.nbytes + obs.nbytes
        return h2d, d2h

    # ---- CPU port of lime's path (oracle/): bounded sample
    def cpu_point(self, idx, variant, nsteps):
        """nsteps RK4 steps of parameter point idx with lime's algorithm; returns seconds"""
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import lime_oracle as lo
        from scipy.sparse import csr_matrix
        H, c_ops, e_ops = lo.jaynes_cummings(self.omega0, self.omega0 + self.det_pts[idx % self.B],
                                             self.g_pts[idx % self.B], self.ncav, self.kappa)
        if variant == 'dense':      # Lindblad_solver.evolve -> _lindblad (np.dot), lime/oqs.py:1590-1688
            Hd, cd, ed = H.toarray(), [c.toarray() for c in c_ops], [e.toarray() for e in e_ops]
            t = time.perf_counter()
            lo.lindblad(Hd, self.rho0, cd, ed, Nt=nsteps, dt=self.dt)
            return time.perf_counter() - t
        # CSR operands and CSR rho through phys.liouvillian (lime/phys.py:561-577), the route of
        # lime/correlation.py; rho starts as the full-pattern mixed state so fill-in is not hidden
        rng = np.random.default_rng(idx)
        a = rng.standard_normal((self.N, self.N)) + 1j * rng.standard_normal((self.N, self.N))
        rho = a @ a.conj().T
        rho = csr_matrix(rho / np.trace(rho))
        t = time.perf_counter()
        for _ in range(nsteps):
            rho = lo.rk4(rho, lo.liouvillian_sp, self.dt, H, c_ops)
            [e.dot(rho).diagonal().sum() for e in e_ops]
        return time.perf_counter() - t


class RedfieldBatch:
    """config 1 throughput variant: the Redfield tensor of examples/redfield.py (N = 2, R 4x4 CSR),
    a batch of random Hermitian trace-1 initial states, RK4 of d vec(rho)/dt = R vec(rho) (lime/oqs.py:453-472)"""
    name = 'redfield_batch'
    metric = 'redfield_rho_steps_per_s'
    unit = 'rho-steps/s'
    dtype = 'f64'                      # complex128 = interleaved f64 pairs
    scaling = 'weak'
    bound = 'hbm'

    def __init__(self, args, rank, world, need_gpu=True):
        self.B = args.batch or (1 << 20)
        self.rk = args.rk_steps or 1000
        self.N = 2
        delta, eps0, gamma1 = 0.2 * 2 * np.pi, 2 * np.pi, 0.5
        sx = np.array([[0., 1.], [1., 0.]])
        sz = np.array([[1., 0.], [0., -1.]])
        self.H = -delta / 2.0 * sx - eps0 / 2.0 * sz
        self.sx = sx
        self.spec = lambda w: gamma1 / 2 * (w / (2 * np.pi))
        self.dt = 20.0 / 199
        rng = np.random.default_rng(rank)
        a = rng.standard_normal((self.B, 2, 2)) + 1j * rng.standard_normal((self.B, 2, 2))
        rho = a @ a.conj().transpose(0, 2, 1)
        self.rho0 = rho / np.einsum('bii->b', rho)[:, None, None]
        self.alg_bytes_per_unit = 2 * 16 * 4
        self.units_per_step = self.B * self.rk
        self.launches = 0
        self.kernel = 'liouville_rk4_kernel'
        if need_gpu:
            import torch
            from lime_b200 import engine
            from lime_b200.oqs import Redfield_solver
            self.torch = torch
            self.solver = Redfield_solver(self.H, c_ops=[sx], spectra=[self.spec])
            self.R, self.evecs = self.solver.redfield_tensor()
            ex = self.evecs.conj().T @ sx @ self.evecs
            self.plan = engine.LiouvillePlan(self.R, e_rows=[ex.T.reshape(-1)])
            v = self.evecs.conj().T[None] @ self.rho0 @ self.evecs[None]
            self.h_v = np.ascontiguousarray(v.reshape(self.B, 4))
            self.v = torch.from_numpy(self.h_v).cuda()

    def config(self):
        return {'workload': 'redfield_batch: Redfield tensor of examples/redfield.py (two-level spin-boson, N=2, R 4x4), '
                            '%d random initial states per GPU, dt=20/199, %d RK4 steps per launch, E=1 observable per step'
                            % (self.B, self.rk), 'batch_per_gpu': self.B, 'rk4_steps_per_launch': self.rk,
                'l2_policy': 'state is 64 B per rho (%.0f MiB per GPU); the %.1f GiB observable stream written per launch '
                             'exceeds L2' % (self.B * 64 / 2 ** 20, self.B * self.rk * 16 / 2 ** 30),
                'sharding': 'initial states split across ranks, no collective'}

    def step(self):
        self.obs, _ = self.plan.run_device(self.v, self.dt, self.rk)
        self.launches += 1

    def check(self):
        tr = (self.v[:, 0] + self.v[:, 3]).cpu().numpy()
        return {'max_trace_error': float(np.max(np.abs(tr - 1.0)))}

    def e2e_setup(self):
        pass

    def e2e_step(self):
        """public API: Redfield_solver.evolve_batch (host in, host out)"""
        out, obs = self.solver.evolve_batch(self.rho0, self.dt, self.rk, e_ops=[self.sx])
        return self.rho0.nbytes, out.nbytes + obs.nbytes

    def cpu_point(self, idx, variant, nsteps):
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import lime_oracle as lo
        R, evecs = lo.redfield_tensor(self.H, [self.sx], [self.spec])
        t = time.perf_counter()
        lo.redfield(R, self.rho0[idx % self.B], evecs=evecs, Nt=nsteps, dt=self.dt, e_ops=[self.sx])
        return time.perf_counter() - t


class LindbladDense:
    """SURVEY.md 8d config 2': random dense Hermitian H (GUE), M = 2 random collapse operators, N = 256, batch 64 --
    the regime where [H, rho] is a real dense contraction (FP64-bound, not HBM-bound)"""
    name = 'lindblad_dense'
    metric = 'lindblad_rho_steps_per_s'
    unit = 'rho-steps/s'
    dtype = 'f64'                      # complex128 = interleaved f64 pairs
    scaling = 'weak'
    bound = 'tensor'

    def __init__(self, args, rank, world, need_gpu=True):
        self.N = args.size or 256
        self.B = args.batch or 64
        self.rk = args.rk_steps or 10
        self.M = 2
        self.dt = 0.002
        rng = np.random.default_rng(1)
        N = self.N
        a = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        self.H = (a + a.conj().T) / (2 * np.sqrt(N))
        self.c_ops = [0.1 * (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))) / np.sqrt(N)
                      for _ in range(self.M)]
        self.e_ops = [self.H]
        r = np.random.default_rng(10 + rank)
        x = r.standard_normal((self.B, N, 8)) + 1j * r.standard_normal((self.B, N, 8))
        rho = x @ x.conj().transpose(0, 2, 1)
        self.rho0 = np.ascontiguousarray(rho / np.einsum('bii->b', rho)[:, None, None])
        self.flops_per_unit = 4 * (2 + 2 * self.M) * 8 * N ** 3
        self.alg_bytes_per_unit = 2 * 16 * N * N
        self.units_per_step = self.B * self.rk
        self.launches = 0
        self.kernel = 'qme_dense_stage'
        if need_gpu:
            import torch
            from lime_b200 import oqs
            self.torch = torch
            self.plan = oqs._lindblad_plan(self.H, self.c_ops, self.e_ops)
            self.rho = torch.from_numpy(self.rho0).cuda()

    def config(self):
        return {'workload': 'lindblad_dense: GUE Hamiltonian N=%d, M=2 dense collapse operators, batch %d per GPU, dt=%g, '
                            '%d RK4 steps per bench step' % (self.N, self.B, self.dt, self.rk),
                'N': self.N, 'batch_per_gpu': self.B, 'flops_per_rho_step': self.flops_per_unit,
                'l2_policy': 'L2 flushed between bench steps (state %.0f MiB)' % (self.B * self.N ** 2 * 16 / 2 ** 20),
                'sharding': 'batch split across ranks, no collective'}

    def step(self):
        self.obs, _ = self.plan.run_device(self.rho, self.dt, self.rk)
        self.launches += self.plan.last_launches

    def check(self):
        tr = self.torch.einsum('bii->b', self.rho).cpu().numpy()
        return {'max_trace_error': float(np.max(np.abs(tr - 1.0)))}

    def e2e_setup(self):
        pass

    def e2e_step(self):
        from lime_b200.oqs import Lindblad_solver
        s = Lindblad_solver(self.H, c_ops=self.c_ops)
        rho_f, obs, _ = s.evolve_batch(self.rho0, self.dt, self.rk, e_ops=self.e_ops)
        return self.rho0.nbytes, rho_f.nbytes + obs.nbytes

    def cpu_point(self, idx, variant, nsteps):
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import lime_oracle as lo
        t = time.perf_counter()
        lo.lindblad(self.H, self.rho0[idx % self.B], self.c_ops, self.e_ops, Nt=nsteps, dt=self.dt)
        return time.perf_counter() - t


class LiouvilleEig:
    """two-time correlation function <A(t)B(t+tau)C(t)> from the eigen-decomposition of the
    Liouvillian (lime/superoperator.py:703-754).  The decomposition (host LAPACK, as in lime) is set-up; a step is
    everything after it: coeff = lhs @ rhs (k^3), cor = tmp1.T @ coeff @ tmp2 on the FP64 tensor cores."""
    name = 'liouville_eig'
    metric = 'two_time_correlation_points_per_s'
    unit = 'points/s'
    dtype = 'f64'                      # complex128 = interleaved f64 pairs
    scaling = 'weak'
    bound = 'tensor'

    def __init__(self, args, rank, world, need_gpu=True):
        self.N = args.size or 16
        self.nt = args.batch or 256
        N = self.N
        rng = np.random.default_rng(5)
        a = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        self.H = (a + a.conj().T) / 2
        self.c_ops = [0.3 * (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))) / np.sqrt(N)]
        self.ops = [rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)) for _ in range(3)]
        x = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        rho = x @ x.conj().T
        self.rho0 = rho / np.trace(rho)
        self.tl = np.linspace(0, 2.0, self.nt)
        self.taul = np.linspace(0, 2.0, self.nt)
        k = N * N
        self.flops_per_unit = 8.0 * (k ** 3 + self.nt * k * k + self.nt * self.nt * k) / (self.nt * self.nt)
        self.alg_bytes_per_unit = 16
        self.units_per_step = self.nt * self.nt
        self.launches = 0
        self.kernel = 'qme_dense_stage_dmma (limeb200_zgemm)'
        if need_gpu:
            import torch
            from lime_b200.superoperator import Lindblad_solver, left, right, operator_to_vector
            from lime_b200 import _dev
            self.torch = torch
            self.solver = Lindblad_solver(self.H, self.c_ops)
            self.solver.eigenstates()
            s = self.solver
            a_, b_, c_ = self.ops
            alpha = (np.conj(s.idv) @ left(b_).dot(s.right_eigvecs)) / s.norm
            beta = (s.left_eigvecs.conj().T @ operator_to_vector(self.rho0)) / s.norm
            X = right(a_).dot(left(c_).dot(s.right_eigvecs))
            self.d = [_dev.to_dev(np.ascontiguousarray(m)) for m in (
                (s.left_eigvecs * np.conj(alpha)[None, :]).conj().T, np.asarray(X) * beta[None, :],
                np.exp(np.outer(self.taul, s.eigvals)), np.exp(np.outer(s.eigvals, self.tl)))]

    def config(self):
        return {'workload': 'liouville_eig: <A(t)B(t+tau)C(t)> on a %d x %d (tau, t) grid from the eigen-decomposition of a '
                            'random N=%d Lindblad generator (k = N^2 = %d); the decomposition itself (host LAPACK) is set-up'
                            % (self.nt, self.nt, self.N, self.N ** 2),
                'l2_policy': 'L2 flushed between bench steps', 'sharding': 'replicas only'}

    def step(self):
        from lime_b200 import engine
        lhs, rhs, t1, t2 = self.d
        coeff = engine.zgemm(lhs, rhs)
        self.out = engine.zgemm(engine.zgemm(t1, coeff), t2)
        self.launches += 3

    def check(self):
        return {}

    def e2e_setup(self):
        pass

    def e2e_step(self):
        out = self.solver.correlation_3op_2t(self.rho0, self.ops, self.tl, self.taul)
        return 4 * self.N ** 4 * 16, out.nbytes

    def cpu_point(self, idx, variant, nsteps):
        """lime's O(k^2) Python loop is hours at k = 256; the CPU sample is the same function at N = 4 (k = 16),
        reported per grid point of ITS grid and therefore only indicative"""
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import lime_oracle as lo
        n = 4
        o = lo.SuperLindblad(self.H[:n, :n], [c[:n, :n] for c in self.c_ops])
        o.eigenstates()
        ops = [x[:n, :n] for x in self.ops]
        r = self.rho0[:n, :n] / np.trace(self.rho0[:n, :n])
        t = time.perf_counter()
        for _ in range(nsteps):
            o.correlation_3op_2t(r, ops, self.tl, self.taul)
        return time.perf_counter() - t


class HeomBase:
    metric = 'heom_ado_steps_per_s'
    unit = 'ADO-steps/s'
    dtype = 'f64'                      # complex128 = interleaved f64 pairs
    bound = 'hbm'


class HeomSpinBoson(HeomBase):
    """config 3 throughput variant: B hierarchies of 91 ADOs (2x2), per-hierarchy (lambda, beta)"""
    name = 'heom_sb'
    scaling = 'weak'

    def __init__(self, args, rank, world, need_gpu=True):
        from lime_b200 import engine
        from lime_b200.heom.heom import _calc_matsubara_params
        self.B = args.batch or 32768
        self.rk = args.rk_steps or 200
        self.dt, self.depth, self.K = 0.01, 12, 2
        sx = np.array([[0., 1.], [1., 0.]], dtype=complex)
        sz = np.array([[1., 0.], [0., -1.]], dtype=complex)
        self.H = 0.5 * 1.0 * sz + 0.5 * 0.5 * sx
        self.sz = sz
        lam = np.linspace(0.05, 0.4, self.B * world)[rank * self.B:(rank + 1) * self.B]
        beta = np.linspace(0.5, 2.0, self.B * world)[::-1][rank * self.B:(rank + 1) * self.B]
        cs, nus = [], []
        for l, b in zip(lam, beta):
            c, nu = _calc_matsubara_params(self.K, l, 1.0, 1.0 / b)
            cs.append(c); nus.append(nu)
        self.c, self.nu = np.array(cs, dtype=complex), np.array(nus, dtype=float)
        self.n = 2
        self.alg_bytes_per_unit = 2 * 16 * self.n * self.n
        self.launches = 0
        if need_gpu:
            import torch
            self.torch = torch
            self.states, self.dn, self.up = engine.heom_tables([self.depth + 1] * self.K, self.depth)
            self.nhe = self.states.shape[0]
            self.plan = engine.HeomPlan(self.H, sz, [0] * self.K, self.c, self.nu, self.states, self.dn, self.up)
            ado = np.zeros((self.B, self.nhe, 2, 2), dtype=complex)
            ado[:, 0, 0, 0] = 1.0
            self.h_ado = ado
            self.ado = torch.from_numpy(ado).cuda()
            self.eT = torch.from_numpy(np.ascontiguousarray(np.stack([sz.T, sx.T]))).cuda()
            self.units_per_step = self.B * self.nhe * self.rk
            self.kernel = 'heom_onchip_kernel'

    def config(self):
        return {'workload': 'heom_sb: spin-boson Drude-Lorentz HEOM, K=2 exponentials, depth 12 (91 ADOs, 312 couplings), '
                            '%d hierarchies per GPU differing in (lambda, beta), dt=%g, %d RK4 steps per launch'
                            % (self.B, self.dt, self.rk),
                'batch_per_gpu': self.B, 'rk4_steps_per_launch': self.rk, 'n_ado': 91,
                'l2_policy': 'inputs larger than L2 (state %.0f MiB per GPU)' % (self.B * 91 * 4 * 16 / 2 ** 20),
                'sharding': 'hierarchies split across ranks, no collective'}

    def step(self):
        self.obs, _ = self.plan.run_device(self.ado, self.dt, self.rk, eT=self.eT)
        self.launches += self.plan.last_launches

    def check(self):
        tr = (self.ado[:, 0, 0, 0] + self.ado[:, 0, 1, 1]).cpu().numpy()
        err = float(np.max(np.abs(tr - 1.0)))
        assert err < 1e-9, 'trace drifted: %g' % err
        return {'max_trace_error': err}

    def e2e_setup(self):
        self.h_pin = self.torch.from_numpy(self.h_ado).pin_memory()

    def e2e_step(self):
        from lime_b200 import _dev
        d = _dev.h2d(self.h_pin)
        obs, _ = self.plan.run_device(d, self.dt, self.rk, eT=self.eT)
        out = _dev.d2h(d, True)
        o = _dev.d2h(obs, True)
        return self.h_pin.numel() * 16, out.nbytes + o.nbytes

    def cpu_point(self, idx, variant, nsteps):
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import lime_oracle as lo
        states, dn, up = lo.heom_tables([self.depth + 1] * self.K, self.depth)
        ado = np.zeros((states.shape[0], 2, 2), dtype=complex)
        ado[0, 0, 0] = 1.0
        t = time.perf_counter()
        lo.heom_rk4(ado, self.H, self.sz[None], [0] * self.K, self.c[idx % self.B], self.nu[idx % self.B],
                    states, dn, up, self.dt, nsteps, e_ops=[self.sz])
        return time.perf_counter() - t


class HeomFMO(HeomBase):
    """config 4: FMO 7 sites x 7 Drude baths x K=2, depth 4 -> 3060 ADOs of 7x7.
    world == 1: `--batch` independent hierarchies (default 1 = the single-hierarchy case).
    world  > 1: ONE hierarchy, ADO-sharded, all-gather of the stage vector after each stage."""
    name = 'heom_fmo'
    scaling = 'strong'

    def __init__(self, args, rank, world, need_gpu=True):
        from lime_b200 import builders as models
        from lime_b200.units import au2fs
        self.B = args.batch or 1
        self.rk = args.rk_steps or 200
        self.depth = args.depth or 4
        self.rank, self.world = rank, world
        self.exchange = args.exchange
        self.Hm, self.Q, self.lam, self.gam, self.kT = models.fmo_heom_inputs()
        self.dt = 0.5 / au2fs
        self.n = 7
        self.alg_bytes_per_unit = 2 * 16 * 49
        self.launches = 0
        from lime_b200 import engine
        self.nhe = engine.heom_tables([self.depth + 1] * 14, self.depth)[0].shape[0]   # host table builder
        if need_gpu:
            import torch
            from lime_b200.heom.heom import HEOM
            self.torch = torch
            if world > 1:
                from lime_b200.heom.sharded import ShardedHEOM
                self.h = ShardedHEOM(self.Hm, self.Q, self.lam, self.gam, self.kT, N_exp=2, N_cut=self.depth,
                                     exchange=args.exchange)
                self.nhe = self.h.nhe
                self.exchange = self.h.exchange
                self.kernel = ('heom_flow_kernel (dataflow: tagged stage vectors, peer stores, no barrier)' if self.h.exchange == 'flow'
                               else 'heom_persist_kernel (grid barrier inside the GPU, tagged halo over NVLink between GPUs)' if self.h.exchange == 'halo'
                               else 'heom_persist_kernel / heom_persist_cached_kernel (fused peer stores + one-hop barrier)' if self.h.exchange == 'p2p'
                               else 'heom_stage_kernel + NCCL all_gather (CUDA graph)')
            else:
                self.h = HEOM(self.Hm, self.Q, self.lam, self.gam, self.kT, N_exp=2, N_cut=self.depth)
                self.nhe = self.h.nhe
                self.kernel = None
            rho0 = np.zeros((7, 7), dtype=complex)
            rho0[0, 0] = 1.0
            ado = np.zeros((self.B, self.nhe, 7, 7), dtype=complex)
            ado[:, 0] = rho0
            self.h_ado = ado
            self.ado = torch.from_numpy(ado).cuda()
            self.eT = torch.from_numpy(np.ascontiguousarray(np.stack([q.T for q in self.Q]))).cuda()
            self.units_per_step = self.B * self.nhe * self.rk / (world if world > 1 else 1)

    def config(self):
        return {'workload': 'heom_fmo: FMO 7-site (Adolphs-Renger), 7 Drude baths (35 cm^-1, 50 fs, 300 K) x K=2, '
                            'depth %d -> %d ADOs of 7x7, dt=0.5 fs, %d RK4 steps per bench step, %d hierarchies'
                            % (self.depth, self.nhe, self.rk, self.B),
                'n_ado': self.nhe, 'rk4_steps_per_launch': self.rk, 'batch': self.B,
                'l2_policy': 'single hierarchy (%.2f MiB) is L2 resident by construction; L2 flushed between bench steps'
                             % (self.nhe * 49 * 16 / 2 ** 20) if self.B * self.nhe * 49 * 16 < 2 ** 27 else 'inputs larger than L2',
                'sharding': ('one hierarchy, ADO ranges per rank, exchange=%s, 4 exchanges per RK4 step' % self.exchange)
                            if self.world > 1 else 'none'}

    def step(self):
        if self.world > 1:
            self.h.run_device(self.ado, self.dt, self.rk)
            self.launches += self.h.last_launches
        else:
            self.obs, _ = self.h.plan.run_device(self.ado, self.dt, self.rk, eT=self.eT)
            self.launches += self.h.plan.last_launches

    def kernel_name(self):
        if self.kernel:
            return self.kernel
        stage = ('heom_stage_fast_kernel (packed neighbour lists, n_k x base coefficients from shared memory; one launch '
                 'per RK4 stage)' if self.B >= 4 else
                 'heom_stage_kernel (packed neighbour lists + coefficient table; one launch per RK4 stage)')
        return {1: 'heom_onchip_kernel', 2: stage,
                3: 'heom_persist_cached_kernel / heom_persist_kernel (one cooperative launch per run)',
                4: 'heom_flow_kernel (dataflow: tagged stage vectors, no barrier; one cooperative launch per run)'
                }.get(self.h.plan.path, 'path %d' % self.h.plan.path)

    def check(self):
        tr = self.torch.einsum('bii->b', self.ado[:, 0]).cpu().numpy()
        err = float(np.max(np.abs(tr - 1.0)))
        assert err < 1e-9, 'trace drifted: %g' % err
        out = {'max_trace_error': err}
        if self.world > 1:
            # sharded vs single-GPU parity, executed in the run: the FULL hierarchy, nchk RK4 steps from rho0, the
            # ADO-sharded propagator over all ranks against the one-GPU propagator of the same library on this rank
            import torch.distributed as dist
            from lime_b200.heom.heom import HEOM
            t = self.torch
            nchk = 200
            a_sh = t.from_numpy(self.h_ado[:1].copy()).cuda()
            self.h.run_device(a_sh, self.dt, nchk)
            single = HEOM(self.Hm, self.Q, self.lam, self.gam, self.kT, N_exp=2, N_cut=self.depth)
            a_1 = t.from_numpy(self.h_ado[:1].copy()).cuda()
            single.plan.run_device(a_1, self.dt, nchk)
            rel = t.max(t.abs(a_sh - a_1)) / t.max(t.abs(a_1))
            rel = rel.to(t.float64).reshape(1)
            dist.all_reduce(rel, op=dist.ReduceOp.MAX)
            rel = float(rel.item())
            assert rel <= 1e-12, 'sharded hierarchy differs from the single-GPU one: %g' % rel
            out.update(sharded_vs_single_relerr=rel, sharded_vs_single_steps=nchk, sharded_vs_single_ados=self.nhe,
                       sharded_vs_single_norm='max|a-b|/max|b| over the full hierarchy, max over ranks')
            del single, a_sh, a_1
        return out

    def close(self):
        if self.world > 1:
            self.h.close()

    def e2e_setup(self):
        pass

    def e2e_step(self):
        """public API: HEOM.evolve(rho0, dt, Nt, e_ops) -> Result (host in, host out)"""
        rho0 = np.zeros((7, 7), dtype=complex)
        rho0[0, 0] = 1.0
        if self.world > 1:
            res = self.h.evolve(rho0, self.dt, self.rk)
            return rho0.nbytes, res.ado.nbytes
        if self.B > 1:          # batch of hierarchies: HeomPlan.run (host batch in, final hierarchies + observables out)
            out, obs, _ = self.h.plan.run(self.h_ado, self.dt, self.rk, e_ops=self.Q)
            return self.h_ado.nbytes, out.nbytes + obs.nbytes
        res = self.h.evolve(rho0, self.dt, self.rk, e_ops=self.Q, store_states=False)
        return self.nhe * 49 * 16, res.ado.nbytes + res.observables.nbytes

    def cpu_point(self, idx, variant, nsteps):
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import lime_oracle as lo
        c, nu, qmap = [], [], []
        for b in range(7):
            cb, nub = lo.calc_matsubara_params(2, self.lam, self.gam, self.kT)
            c += cb; nu += nub; qmap += [b, b]
        states, dn, up = lo.heom_tables([self.depth + 1] * 14, self.depth)
        ado = np.zeros((states.shape[0], 7, 7), dtype=complex)
        ado[0, 0, 0] = 1.0
        t = time.perf_counter()
        lo.heom_rk4(ado, self.Hm, np.stack(self.Q), qmap, np.array(c), np.array(nu), states, dn, up,
                    self.dt, nsteps, e_ops=[self.Q[0]])
        return time.perf_counter() - t


class Sos2DES:
    """config 5: photon echo GSB+SE+ESA on a 256x256 (omega1, omega3) grid x 64 waiting times"""
    name = 'sos_2des'
    metric = 'sos_grid_points_per_s'
    unit = 'grid-points/s'
    dtype = 'f64'                      # complex128 = interleaved f64 pairs
    scaling = 'weak'
    bound = 'hbm'

    def __init__(self, args, rank, world, need_gpu=True):
        au2ev, au2fs = 27.211386, 2.41888432651e-2
        self.n = 256
        self.T = args.batch or 64
        rng = np.random.default_rng(0)
        N, ne = 32, 15
        nf = N - 1 - ne
        E = np.zeros(N)
        E[1:1 + ne] = np.sort(rng.uniform(1.5, 2.0, ne)) / au2ev
        E[1 + ne:] = np.sort(rng.uniform(3.2, 3.8, nf)) / au2ev
        dip = np.zeros((N, N))
        ge = rng.standard_normal(ne)
        dip[0, 1:1 + ne] = ge; dip[1:1 + ne, 0] = ge
        ef = rng.standard_normal((ne, nf))
        dip[1:1 + ne, 1 + ne:] = ef; dip[1 + ne:, 1:1 + ne] = ef.T
        gamma = np.ones(N) * 0.05 / au2ev
        gamma[0] = 0.0
        self.sys = (E, dip, gamma, [0], list(range(1, 1 + ne)), list(range(1 + ne, N)))
        self.w = np.linspace(1.4, 2.1, self.n) / au2ev
        Tall = np.linspace(0, 630, self.T * world) / au2fs
        self.taus = Tall[rank * self.T:(rank + 1) * self.T]
        self.alg_bytes_per_unit = 16
        self.units_per_step = self.T * self.n * self.n
        self.launches = 0
        self.kernel = 'sos_outer_kernel'
        self.grid = None
        if need_gpu:
            import torch
            self.torch = torch

    def config(self):
        return {'workload': 'sos_2des: _photon_echo (GSB+SE+ESA) N=32 (15 e, 16 f), 256x256 grid, %d waiting times per GPU'
                            % self.T, 'l2_policy': 'output (%d MiB) rewritten each step; L2 flushed between steps'
                            % (self.units_per_step * 16 // 2 ** 20),
                'sharding': 'waiting times split across ranks, no collective'}

    def step(self):
        from lime_b200.signal import sos
        if self.grid is None:       # weights / grids resident in HBM; the step is the O(grid) device work
            E, dip, gamma, g, e, f = self.sys
            self.grid = sos.PhotonEchoGrid(E, dip, -self.w, self.w, self.taus, g, e, f, gamma)
        self.out = self.grid.run()
        self.launches += 3          # pole factor, weighted factor, rank-R outer product

    def check(self):
        """one waiting-time plane of the timed output against lime's algorithm (oracle port) on the host"""
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import lime_oracle as lo
        E, dip, gamma, g, e, f = self.sys
        k = self.T // 2
        ref = lo.photon_echo_core(E, dip, -self.w, self.w, self.taus[k], g, e, f, gamma)
        got = self.out[k].cpu().numpy()
        err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
        assert err <= 1e-10, 'SOS plane differs from the oracle: %g' % err
        return {'oracle_plane_relerr': err, 'plane': int(k)}

    def e2e_setup(self):
        pass

    def e2e_step(self):
        from lime_b200.signal import sos
        E, dip, gamma, g, e, f = self.sys
        out = sos._photon_echo(E, dip, -self.w, self.w, self.taus, g, e, f, gamma)
        return 2 * self.w.nbytes + dip.nbytes, out.nbytes

    def cpu_point(self, idx, variant, nsteps):
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import lime_oracle as lo
        E, dip, gamma, g, e, f = self.sys
        t = time.perf_counter()
        for k in range(nsteps):
            lo.photon_echo_core(E, dip, -self.w, self.w, self.taus[(idx + k) % self.T], g, e, f, gamma)
        return time.perf_counter() - t


WORKLOADS = {c.name: c for c in (JCLindblad, HeomFMO, HeomSpinBoson, Sos2DES, RedfieldBatch, LindbladDense, LiouvilleEig)}
# units one cpu_point "step" stands for
CPU_UNITS = {'jc_lindblad': lambda w: 1, 'liouville_eig': lambda w: w.nt * w.nt, 'lindblad_dense': lambda w: 1, 'redfield_batch': lambda w: 1, 'heom_sb': lambda w: 91, 'heom_fmo': lambda w: None,
             'sos_2des': lambda w: w.n * w.n}


def _cpu_worker(job):
    wname, argd, idx, variant, nsteps = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    args = argparse.Namespace(**argd)
    w = WORKLOADS[wname](args, 0, 1, need_gpu=False)
    return w.cpu_point(idx, variant, nsteps)


def cpu_units(w):
    u = CPU_UNITS[w.name](w)
    return w.nhe if u is None else u


def cpu_throughput(w, args, budget_s, cores):
    """lime's CPU algorithm (oracle/ port) on a bounded sample of workload `w`, fanned out over
    `cores` processes (one parameter point / hierarchy / waiting time per process, BLAS
    threads pinned to 1).  Returns (units per second, description)."""
    import multiprocessing as mp
    argd = vars(args)
    variants = ['dense', 'csr'] if w.name == 'jc_lindblad' else ['port']
    per_step = {}
    for v in variants:                           # calibrate on one step, single process
        t1 = _cpu_worker((w.name, argd, 0, v, 1))
        per_step[v] = t1
    best = min(per_step, key=per_step.get)
    nsteps = int(max(1, min(200, budget_s / max(per_step[best], 1e-6))))
    units = cpu_units(w)
    ctx = mp.get_context('spawn')
    jobs = [(w.name, argd, i, best, nsteps) for i in range(cores)]
    # each worker times its own compute section (set-up of the synthetic inputs is excluded); the processes run
    # concurrently, so the sample's duration is the slowest worker's
    if cores > 1:
        with ctx.Pool(cores) as pool:
            pool.map(_cpu_worker, [(w.name, argd, i, best, 1) for i in range(cores)])     # spawn + import warm-up
            wall = max(pool.map(_cpu_worker, jobs))
    else:
        wall = _cpu_worker(jobs[0])
    value = cores * nsteps * units / wall
    desc = ('%d processes x %d RK4 steps/evaluations of one %s each, variant=%s (single-process calibration: %s s/step)'
            % (cores, nsteps, 'point' if w.name != 'sos_2des' else 'waiting time', best,
               ', '.join('%s %.4g' % kv for kv in per_step.items())))
    return value, desc, best


# ---------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    args.workload = args.workload or 'jc_lindblad'
    w = WORKLOADS[args.workload](args, 0, 1, need_gpu=False)
    cores = os.cpu_count() or 1
    vals = []
    desc = ''
    total = args.warmup + args.steps
    budget = max(1.0, min(8.0, 150.0 / max(total, 1) / 2.0))
    t_all = time.perf_counter()
    for i in range(total):
        v, desc, variant = cpu_throughput(w, args, budget, cores)
        if i >= args.warmup:
            vals.append(v)
    wall = time.perf_counter() - t_all
    value = float(np.mean(vals)) if vals else 0.0
    cfg = w.config()
    line = {'impl': 'reference', 'metric': w.metric, 'value': value, 'unit': w.unit, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * wall / max(total, 1), 'higher_is_better': True, 'scaling': w.scaling,
            'vs_baseline': None, 'dtype': w.dtype, 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': value, 'unit': w.unit, 'cores': cores, 'kind': 'port', 'sample': desc},
            'e2e': {'value': value, 'unit': w.unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
            'note': "lime's own NumPy/SciPy algorithm (oracle/lime_oracle.py, a call-for-call port; /root/reference is "
                    "absent on the GPU box) on the host cores; lime itself is single-threaded Python, the fan-out over "
                    "points is the most its CPU path can use"}
    print(json.dumps(line), flush=True)
    return 0


def _fp64_tensor_roof(torch):
    """FP64 roof measured in this run: cuBLAS ZGEMM 4096^3 through torch.matmul, best of 10 (MEASURED_PEAKS.json has
    bf16 only); flops counted as 8 N^3 per complex product"""
    n = 4096
    xa = torch.randn(n, n, dtype=torch.complex128, device='cuda')
    xb = torch.randn(n, n, dtype=torch.complex128, device='cuda')
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(xa, xb); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return 8.0 * n ** 3 / best / 1e12, 'measured in this run: cuBLAS ZGEMM 4096^3 (torch.matmul complex128), best of 10'


def measure(w, args, steps, warmup, rank, world, local, with_cpu):
    """W warm-up steps, K timed steps (CUDA events on the launching stream, barrier + synchronize on both sides, max
    over ranks), the workload's own check, the end-to-end leg through the public API and, on request, the CPU sample.
    Returns the JSON line of workload `w` as a dict (identical on every rank up to rank-local clocks)."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = None
    if 'flushed' in json.dumps(w.config()):
        flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream()
    for _ in range(warmup):
        w.step()
    barrier()
    w.launches = 0
    try:
        gpu_id = 'GPU-' + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_id = local
    sampler = ClockSampler(gpu_id)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall = time.perf_counter()
    for k in range(steps):
        if flush is not None:
            flush.fill_(k & 0xff)
        ev[k][0].record(st)
        w.step()
        ev[k][1].record(st)
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device='cuda')
    launches = torch.tensor([w.launches], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    total_ms = float(total_ms.item())
    chk = w.check()
    units_all = w.units_per_step * world * steps
    value = units_all / (total_ms * 1e-3)
    peak, peak_src = measured_peaks()
    kernel_s = (total_ms / steps) * 1e-3
    # per-launch figures of the rank's own launch: units_per_step units of alg_bytes_per_unit each
    achieved = w.alg_bytes_per_unit * w.units_per_step / kernel_s / 1e9
    roof_unit = 'GB/s'
    if w.bound == 'tensor':
        peak, peak_src = _fp64_tensor_roof(torch)
        achieved = w.flops_per_unit * w.units_per_step / kernel_s / 1e12
        roof_unit = 'TFLOP/s'

    # ---- end to end through the public API with host buffers
    w.e2e_setup()
    w.e2e_step()                                  # warm (pinned pools, plan caches)
    barrier()
    ne2e = max(1, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(ne2e):
        h2d, d2h = w.e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = w.units_per_step * world * ne2e / float(e2e_t.item())

    traffic = TRAFFIC.get((w.name, int(w.units_per_step)), {})
    kernel = w.kernel_name() if hasattr(w, 'kernel_name') else w.kernel
    line = {'metric': w.metric, 'value': value, 'unit': w.unit, 'n_gpus': world, 'steps': steps,
            'warmup': warmup, 'ms_per_step': total_ms / steps, 'higher_is_better': True,
            'scaling': w.scaling, 'vs_baseline': None, 'dtype': w.dtype, 'data': 'synthetic',
            'config': w.config(),
            'e2e': {'value': e2e_val, 'unit': w.unit, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'api': 'lime_b200 public API (solver object reused: operator plan built in the warm-up call), pinned host '
                           'buffers, %d steps, wall clock' % ne2e},
            'gpu_launches': int(launches.item()),
            'clocks': clocks,
            'roofline': {'bound': w.bound, 'achieved': achieved, 'peak': peak, 'unit': roof_unit,
                         'frac': achieved / peak, 'traffic': traffic.get('bytes'),
                         'traffic_source': traffic.get('source', 'not captured for this workload/size'),
                         'kernel': kernel, 'peak_source': peak_src,
                         'algorithmic_bytes_per_unit': w.alg_bytes_per_unit,
                         'units_per_launch': w.units_per_step,
                         'note': getattr(w, 'roof_note', None) or
                                 'state stays on chip for all %s RK4 steps of a launch; achieved = algorithmic bytes '
                                 '(one read + one write of every rho/ADO per RK4 step) / CUDA-event time, so it is a '
                                 'throughput normalised to the HBM roof, not measured DRAM traffic (see traffic)'
                                 % getattr(w, 'rk', 1)},
            'wall_s_timed_region': t_wall, 'check': chk}
    if rank == 0 and world == 1 and with_cpu:
        cores = os.cpu_count() or 1
        v, desc, variant = cpu_throughput(w, args, with_cpu, cores)
        line['cpu_baseline'] = {'value': v, 'unit': w.unit, 'cores': cores, 'kind': 'port', 'sample': desc}
    del flush
    return line


# The HEOM half of the BASELINE metric, measured in the same invocation as the Lindblad headline (default run only).
# (workload, overrides, label).  At N > 1 heom_fmo is ONE hierarchy ADO-sharded over the ranks (strong scaling, fused
# peer-memory kernel), with a sharded-vs-single-GPU parity check executed in the run.
def heom_suite(world):
    if world == 1:
        return [('heom_fmo', dict(depth=4, batch=1, rk_steps=400), 'config 4 on one GPU (persistent kernel)'),
                ('heom_fmo', dict(depth=6, batch=1, rk_steps=50), 'config 4 deepened to N_c = 6 (38 760 ADOs)'),
                ('heom_fmo', dict(depth=4, batch=64, rk_steps=8), 'config 4, batch of 64 hierarchies (HBM-resident)'),
                ('heom_sb', dict(batch=32768, rk_steps=200), 'config 3 throughput variant'),
                ('sos_2des', dict(batch=64, rk_steps=0), 'config 5: 2DES grid, 64 waiting times')]
    return [('heom_fmo', dict(depth=4, batch=1, rk_steps=400), 'config 4, ADO-sharded over %d GPUs' % world),
            ('heom_fmo', dict(depth=6, batch=1, rk_steps=50), 'config 4 deepened to N_c = 6, ADO-sharded over %d GPUs' % world),
            ('heom_sb', dict(batch=32768, rk_steps=200), 'config 3 throughput variant, hierarchies split over the ranks'),
            ('sos_2des', dict(batch=64, rk_steps=0), 'config 5: 2DES grid, 64 waiting times per GPU (grid-sharded, no collective)')]


def run_ours(args):
    import copy
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    assert world == args.gpus or world == 1, 'launch with torchrun --nproc-per-node %d' % args.gpus

    suite = args.workload is None and not args.no_suite
    args.workload = args.workload or 'jc_lindblad'
    w = WORKLOADS[args.workload](args, rank, world)
    line = measure(w, args, args.steps, args.warmup, rank, world, local, 0 if args.no_cpu else args.cpu_seconds)
    del w
    torch.cuda.empty_cache()
    if suite:
        subs = []
        for name, over, label in heom_suite(world):
            a2 = copy.copy(args)
            a2.workload = name
            for k, v in over.items():
                setattr(a2, k, v)
            t0 = time.perf_counter()
            try:
                w2 = WORKLOADS[name](a2, rank, world)
                # CPU sample only where one RK4 step of the port takes well under a second (not the 38 760-ADO hierarchy;
                # the batch-of-64 line shares the per-unit CPU figure of the single hierarchy)
                cpu = 0 if (args.no_cpu or over.get('depth', 4) > 4 or over.get('batch', 1) == 64) else min(args.cpu_seconds, 2.0)
                sub = measure(w2, a2, min(args.steps, 5), 3, rank, world, local, cpu)
                if hasattr(w2, 'close'):
                    w2.close()
                del w2
            except Exception as exc:                       # a failed sub-workload must not lose the headline line
                # (multi-rank runs too: the sharded kernels bound their waits, so a protocol failure raises on EVERY
                #  rank and the ranks stay in step for the next sub-workload; re-raising would only lose the line)
                sub = {'error': '%s: %s' % (type(exc).__name__, str(exc)[:300])}
                print('bench: sub-workload %s failed on rank %d: %s' % (name, rank, sub['error']), file=sys.stderr, flush=True)
            torch.cuda.empty_cache()
            sub['label'] = label
            sub['wall_s_total'] = time.perf_counter() - t0
            for k in ('clocks', 'higher_is_better', 'vs_baseline', 'data'):
                sub.pop(k, None)
            subs.append(sub)
        line['heom'] = subs               # (the last entry is the grid-sharded 2DES workload of config 5)
        line['gpu_launches_heom'] = sum(int(x.get('gpu_launches', 0)) for x in subs)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from an ncu capture of the SAME command
# and sizes (not measured in the bench run itself: ncu replays kernels, so it cannot share a run with the timed region);
# key = (workload, units per launch); workloads / sizes without a capture report null
TRAFFIC = {
    ('jc_lindblad', 4096000): {'bytes': 1178165248 + 1183468544,
                               'source': 'profiles/r02_traffic_jc_lindblad_default.csv (ncu --metrics dram__bytes_read.sum,'
                                         'dram__bytes_write.sum on one qme_tile_kernel launch of `python bench.py`: rho in, rho '
                                         'out, observables = 0.11 % of the algorithmic bytes; the state is on chip for 1000 steps)'},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=None, choices=sorted(WORKLOADS),
                    help='one workload only; default: jc_lindblad as the headline line plus the HEOM suite under "heom"')
    ap.add_argument('--no-suite', action='store_true', help='default run without the HEOM sub-lines')
    ap.add_argument('--rk-steps', type=int, default=0, help='RK4 steps per launch (0 = workload default)')
    ap.add_argument('--batch', type=int, default=0, help='units per GPU (0 = workload default)')
    ap.add_argument('--size', type=int, default=0, help='Hilbert dimension for lindblad_dense (0 = 256)')
    ap.add_argument('--depth', type=int, default=0, help='HEOM depth for heom_fmo (0 = 4)')
    ap.add_argument('--exchange', default='auto', choices=['auto', 'flow', 'halo', 'p2p', 'nccl'],
                    help='sharded heom_fmo: dataflow kernel (auto/flow), barrier kernel with peer stores (p2p) or stage kernel + NCCL all-gather')
    ap.add_argument('--cpu-seconds', type=float, default=8.0, help='per-process budget of the cpu_baseline sample')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-spot-check', action='store_true', help='skip the in-run oracle spot check of jc_lindblad')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        log('note: fewer than 3 warm-up steps requested')
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
