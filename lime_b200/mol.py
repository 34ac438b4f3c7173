"""Result container of the density-matrix solvers (lime/mol.py:78-104).

lime's constructor does arithmetic on nout=None / t0=None for every density-matrix solver
(lime/mol.py:92 called from lime/oqs.py:449,1670,1780) and raises TypeError; here the
defaults are nout=1, t0=0.0 -- the fields and their meaning are unchanged."""
import numpy as np


class Result:
    def __init__(self, description=None, psi0=None, rho0=None, dt=None, Nt=None, times=None,
                 t0=None, nout=None):
        self.description = description
        self.dt = dt
        self.timesteps = Nt
        self.observables = None
        self.rholist = None
        self.psilist = [psi0]
        self.psi = None
        self.rho0 = rho0
        self.psi0 = psi0
        nout = 1 if nout is None else nout
        t0 = 0.0 if t0 is None else t0
        self.nout = nout
        if Nt is None and times is not None:
            Nt = len(times)
        self.times = t0 + np.arange(Nt // nout) * dt * nout if (Nt is not None and dt is not None) else times

    def expect(self):
        return self.observables
