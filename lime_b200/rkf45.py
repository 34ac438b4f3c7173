"""
Runge-Kutta-Fehlberg 4(5) with the call signature lime's examples/rkf45_test.py uses
(`from lime.rkf45 import *` at :7, `y, yp, t, flag = r8_rkf45(f, neqn, y, yp, t, tout, relerr, abserr, flag)`
at :115, :196, :321).  The module `lime/rkf45.py` itself is NOT in the lime tree, so there is no
reference output to match: PARITY UNPINNED, accuracy pinned on the analytic problems of that example file
(tests/test_gpu_parity.py::test_rkf45_*).

What is implemented is the classical algorithm those names refer to -- Fehlberg's fourth-fifth order pair with the
Shampine-Watts step-size control (flag protocol: +1/-1 start, +2/-2 continue, 3 relerr too small, 4 too many
evaluations, 5 pure relative test impossible, 6 accuracy not reachable, 7 too much output, 8 invalid input).

Where it runs: the state, the five slopes and the solution candidate are DEVICE vectors; every Fehlberg stage is one
fused kernel (limeb200_rkf45_stage), the solution + error estimate + max-reduction another (limeb200_rkf45_error),
see lime_b200/csrc/rkf45.cu.  Only the two reduced doubles per attempted step come back to the host, where the step
controller lives.  The right-hand side `f(t, y)` is
  * a device function -- e.g. `QmeRHS(plan)` / `HeomRHS(plan)` below, which wrap limeb200_qme_rhs / limeb200_heom_rhs:
    rho stays resident in HBM for the whole integration, or
  * a host callable on NumPy arrays, as in lime's example (then each stage argument is copied to the host and the
    slope back: this is the drop-in form for small systems, not the fast path).
There is no CPU implementation of the stage arithmetic: without the CUDA library the calls raise.
"""
import ctypes as C
import numpy as np
import torch

from . import _dev
from ._lib import lib, check

_EPS = float(np.finfo(np.float64).eps)
_REMIN = 1.0e-12
_MAXNFE = 3000


class RKF45State:
    """what Burkardt's routine keeps in persistent variables between calls"""

    def __init__(self):
        self.h = -1.0
        self.nfe = -1
        self.kop = -1
        self.init = -1000
        self.kflag = -1000
        self.flag_save = -1000
        self.relerr_save = -1.0
        self.abserr_save = -1.0
        self.work = None          # device workspace (f1..f5, s, result2), reused while n stays the same
        self.steps_accepted = 0
        self.steps_rejected = 0


_default_state = RKF45State()


def _sign(a, b):
    return abs(a) if b >= 0 else -abs(a)


class _Vec:
    """device view of the caller's vector as n real doubles + the conversions at the API boundary"""

    def __init__(self, y, dev):
        self.host = not isinstance(y, torch.Tensor)
        if self.host:
            a = np.asarray(y)
            self.cplx = np.iscomplexobj(a)
            self.shape = a.shape
            self.t = _dev.to_dev(a, np.complex128 if self.cplx else np.float64, dev)
        else:
            assert y.is_cuda and y.is_contiguous() and y.dtype in (torch.float64, torch.complex128)
            self.cplx = y.dtype == torch.complex128
            self.shape = tuple(y.shape)
            self.t = y
        self.real = torch.view_as_real(self.t).reshape(-1) if self.cplx else self.t.reshape(-1)
        self.n = self.real.numel()

    def like(self, real_flat):
        """tensor of the caller's dtype/shape over the storage `real_flat` (n doubles)"""
        if self.cplx:
            return torch.view_as_complex(real_flat.view(-1, 2)).view(self.shape)
        return real_flat.view(self.shape)

    def out(self, real_flat):
        t = self.like(real_flat)
        return t.cpu().numpy() if self.host else t


def _call_f(f, t, vec, arg_real, out_real):
    """out_real <- f(t, arg) with f a device function or a host callable"""
    arg = vec.like(arg_real)
    if vec.host:
        r = np.asarray(f(t, arg.cpu().numpy()))
        out_real.copy_(_Vec(r.astype(np.complex128 if vec.cplx else np.float64, copy=False).reshape(vec.shape),
                            arg_real.device).real)
    else:
        r = f(t, arg)
        r = r.contiguous()
        out_real.copy_(torch.view_as_real(r).reshape(-1) if r.dtype == torch.complex128 else r.reshape(-1))


def _fehl_device(f, vec, y, t, h, yp, w, sp):
    """the five new slopes of one Fehlberg step, each stage argument built by ONE fused kernel; w = workspace dict"""
    L = lib()
    n = vec.n
    p = _dev.ptr
    f1, f2, f3, f4, f5, arg = w['f1'], w['f2'], w['f3'], w['f4'], w['f5'], w['arg']
    check(L.limeb200_rkf45_stage(1, n, p(y), p(yp), None, None, None, None, None, h, p(arg), sp))
    _call_f(f, t + h / 4.0, vec, arg, f1)
    check(L.limeb200_rkf45_stage(2, n, p(y), p(yp), p(f1), None, None, None, None, h, p(arg), sp))
    _call_f(f, t + 3.0 * h / 8.0, vec, arg, f2)
    check(L.limeb200_rkf45_stage(3, n, p(y), p(yp), p(f1), p(f2), None, None, None, h, p(arg), sp))
    _call_f(f, t + 12.0 * h / 13.0, vec, arg, f3)
    check(L.limeb200_rkf45_stage(4, n, p(y), p(yp), p(f1), p(f2), p(f3), None, None, h, p(arg), sp))
    _call_f(f, t + h, vec, arg, f4)
    check(L.limeb200_rkf45_stage(5, n, p(y), p(yp), p(f1), p(f2), p(f3), p(f4), None, h, p(arg), sp))
    _call_f(f, t + h / 2.0, vec, arg, f5)


def _workspace(state, n, dev):
    w = state.work
    if w is None or w['n'] != n or w['dev'] != dev:
        w = {k: torch.empty(n, dtype=torch.float64, device=dev) for k in ('f1', 'f2', 'f3', 'f4', 'f5', 'arg', 's')}
        w['res'] = torch.empty(2, dtype=torch.float64, device=dev)
        w['n'], w['dev'] = n, dev
        state.work = w
    return w


def r8_fehl(f, neqn, y, t, h, yp):
    """one Fehlberg step of size h WITHOUT error control: returns (f1, f2, f3, f4, f5, s), s the new solution"""
    dev = y.device if isinstance(y, torch.Tensor) else _dev.device()
    vy, vyp = _Vec(y, dev), _Vec(yp, dev)
    st = RKF45State()
    w = _workspace(st, vy.n, dev)
    sp = _dev.stream_ptr(dev)
    _fehl_device(f, vy, vy.real, t, h, vyp.real, w, sp)
    check(lib().limeb200_rkf45_error(vy.n, _dev.ptr(vy.real), _dev.ptr(vyp.real), _dev.ptr(w['f2']), _dev.ptr(w['f3']),
                                     _dev.ptr(w['f4']), _dev.ptr(w['f5']), h, 0.0, _dev.ptr(w['s']), _dev.ptr(w['res']), sp))
    return tuple(vy.out(w[k].clone()) for k in ('f1', 'f2', 'f3', 'f4', 'f5', 's'))


def r8_rkf45(f, neqn, y, yp, t, tout, relerr, abserr, flag, state=None):
    """advance y from t towards tout; returns (y, yp, t, flag).  See the module docstring for the flag protocol.
    `state` (an RKF45State) isolates concurrent integrations; by default one module-level state is used, like the
    persistent variables of the classical routine."""
    S = _default_state if state is None else state
    dev = y.device if isinstance(y, torch.Tensor) else _dev.device()
    if neqn < 1 or relerr < 0.0 or abserr < 0.0 or flag == 0 or flag > 8 or flag < -2:
        return y, yp, t, 8
    vy = _Vec(y, dev)
    if int(np.prod(vy.shape)) != neqn:
        return y, yp, t, 8
    mflag = abs(flag)
    # ---- continuation call: consistency checks of the classical routine
    if mflag != 1:
        if t == tout and S.kflag != 3:
            return y, yp, t, 8
        if mflag == 2:
            if S.kflag == 3 or S.init == 0:
                flag = S.flag_save
                mflag = abs(flag)
            elif S.kflag == 4:
                S.nfe = 0
            elif S.kflag == 5 and abserr == 0.0:
                raise RuntimeError('r8_rkf45: KFLAG = 5 and ABSERR = 0 -- a pure relative error test is impossible')
            elif S.kflag == 6 and relerr <= S.relerr_save and abserr <= S.abserr_save:
                raise RuntimeError('r8_rkf45: KFLAG = 6 and the tolerances were not increased')
        else:
            if flag == 3:
                flag = S.flag_save
                if S.kflag == 3:
                    mflag = abs(flag)
            elif flag == 4:
                S.nfe = 0
                flag = S.flag_save
                if S.kflag == 3:
                    mflag = abs(flag)
            elif flag == 5 and abserr > 0.0:
                flag = S.flag_save
                if S.kflag == 3:
                    mflag = abs(flag)
            else:
                raise RuntimeError('r8_rkf45: integration cannot be continued (flag = %d was not reset)' % flag)
    S.flag_save = flag
    S.kflag = 0
    S.relerr_save = relerr
    S.abserr_save = abserr
    relerr_min = 2.0 * _EPS + _REMIN
    if relerr < relerr_min:
        S.kflag = 3
        return y, yp, t, 3
    L = lib()
    p = _dev.ptr
    sp = _dev.stream_ptr(dev)
    n = vy.n
    w = _workspace(S, n, dev)
    yr = vy.real if not vy.host else vy.real.clone()
    vyp = _Vec(yp, dev)
    ypr = vyp.real if not vyp.host else vyp.real.clone()
    dt = tout - t

    def done(tt, fl):
        return vy.out(yr), vyp.out(ypr), tt, fl

    if mflag == 1:
        S.init = 0
        S.kop = 0
        _call_f(f, t, vy, yr, ypr)
        S.nfe = 1
        if t == tout:
            return done(t, 2)
    if S.init == 0:
        S.init = 1
        h0 = abs(dt)
        check(L.limeb200_rkf45_hinit(n, p(yr), p(ypr), relerr, abserr, h0, p(w['res']), sp))
        tolmax, hcand = w['res'].cpu().tolist()
        S.h = hcand if tolmax > 0.0 else 0.0
        S.h = max(S.h, 26.0 * _EPS * max(abs(t), abs(dt)))
        flag = 2 if flag >= 0 else -2
        S.flag_save = flag
    S.h = _sign(S.h, dt)
    if 2.0 * abs(dt) <= abs(S.h):
        S.kop += 1
    if S.kop == 100:
        S.kop = 0
        return done(t, 7)
    if abs(dt) <= 26.0 * _EPS * abs(t):
        # too close to the output point: extrapolate
        check(L.limeb200_rkf45_axpy(n, dt, p(ypr), p(yr), sp))
        t = tout
        _call_f(f, t, vy, yr, ypr)
        S.nfe += 1
        return done(t, 2)
    output = False
    scale = 2.0 / relerr
    ae = scale * abserr
    while True:
        hfaild = False
        hmin = 26.0 * _EPS * abs(t)
        dt = tout - t
        if abs(dt) < 2.0 * abs(S.h):
            if abs(dt) <= abs(S.h):
                output = True
                S.h = dt
            else:
                S.h = 0.5 * dt
        while True:
            if S.nfe > _MAXNFE:
                S.kflag = 4
                return done(t, 4)
            _fehl_device(f, vy, yr, t, S.h, ypr, w, sp)
            S.nfe += 5
            check(L.limeb200_rkf45_error(n, p(yr), p(ypr), p(w['f2']), p(w['f3']), p(w['f4']), p(w['f5']), S.h, ae,
                                         p(w['s']), p(w['res']), sp))
            eeoet, etmin = w['res'].cpu().tolist()           # the only device -> host traffic of a step: 16 bytes
            if etmin <= 0.0:
                return done(t, 5)
            esttol = abs(S.h) * eeoet * scale / 752400.0
            if esttol <= 1.0:
                break
            hfaild = True
            output = False
            S.steps_rejected += 1
            s = 0.9 / esttol ** 0.2 if esttol < 59049.0 else 0.1
            S.h = s * S.h
            if abs(S.h) < hmin:
                S.kflag = 6
                return done(t, 6)
        # accept
        t = t + S.h
        yr.copy_(w['s'])
        _call_f(f, t, vy, yr, ypr)
        S.nfe += 1
        S.steps_accepted += 1
        s = 0.9 / esttol ** 0.2 if esttol > 0.0001889568 else 5.0
        if hfaild:
            s = min(s, 1.0)
        S.h = _sign(max(s * abs(S.h), hmin), S.h)
        if output:
            return done(tout, 2)
        if flag <= 0:
            return done(t, -2)


# ---------------------------------------------------------------------------------------
# device right-hand sides of the density-matrix path
# ---------------------------------------------------------------------------------------
class QmeRHS:
    """f(t, rho) = L rho of a finalized engine.QmePlan (limeb200_qme_rhs): rho [B,N,N] complex128 on the device"""

    def __init__(self, plan):
        self.plan = plan

    def __call__(self, t, rho):
        out = torch.empty_like(rho)
        r = rho.view(-1, self.plan.N, self.plan.N)
        check(lib().limeb200_qme_rhs(self.plan._h, _dev.ptr(r), _dev.ptr(out), r.shape[0], _dev.stream_ptr(rho.device)))
        return out


class HeomRHS:
    """f(t, ado) of an engine.HeomPlan (limeb200_heom_rhs): ado [B,N_he,n,n] complex128 on the device"""

    def __init__(self, plan):
        self.plan = plan

    def __call__(self, t, ado):
        out = torch.empty_like(ado)
        a = ado.view(-1, self.plan.nhe, self.plan.n, self.plan.n)
        check(lib().limeb200_heom_rhs(self.plan._h, _dev.ptr(a), _dev.ptr(out), a.shape[0], _dev.stream_ptr(ado.device)))
        return out


def integrate(f, y0, tlist, relerr=1e-8, abserr=1e-10, state=None):
    """[ext] y(t) at every t of `tlist` (tlist[0] is the initial time).  y0: NumPy array (returned states are NumPy)
    or a device tensor (states are device tensors and f is a device function).  Returns (states list, RKF45State)."""
    S = RKF45State() if state is None else state
    relerr = max(relerr, 2.0 * _EPS + _REMIN)      # below this the routine answers flag 3 ("relerr too small")
    host = not isinstance(y0, torch.Tensor)
    y = np.array(y0, copy=True) if host else y0.clone()
    yp = np.zeros_like(y) if host else torch.zeros_like(y)
    neqn = int(np.prod(y.shape))
    flag = 1
    t = float(tlist[0])
    out = [y.copy() if host else y.clone()]
    for tout in tlist[1:]:
        y, yp, t, flag = r8_rkf45(f, neqn, y, yp, t, float(tout), relerr, abserr, flag, state=S)
        while flag == 4:                           # more than 3000 evaluations since the last reset -- keep going
            y, yp, t, flag = r8_rkf45(f, neqn, y, yp, t, float(tout), relerr, abserr, flag, state=S)
        if flag != 2:
            raise RuntimeError('r8_rkf45 stopped with flag %d at t = %g' % (flag, t))
        out.append(y.copy() if host else y.clone())
    return out, S
