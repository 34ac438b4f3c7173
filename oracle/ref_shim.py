"""
TEST INFRASTRUCTURE ONLY -- import shim for the read-only reference tree.

Loads binggu56/lime from /root/reference *in this container* so that
(a) oracle/lime_oracle.py (the NumPy restatement) can be validated against the
real reference, and (b) oracle/gen_golden.py can freeze reference outputs into
tests/golden/.  /root/reference does not exist on the GPU box: nothing that
runs there (pytest -m gpu, smoke(), bench.py) may import this module.

The shim does not alter any arithmetic of the reference:
  * plotting modules that are not installed (proplot, matplotlib, mpl_toolkits)
    are replaced by permissive stubs (lime/fft.py:6, lime/mol.py:25,
    lime/signal/sos.py:10,14, lime/superoperator.py:16, lime/style.py:1-7);
  * opt_einsum.contract -> numpy.einsum (lime/oqs.py:23,219,366);
  * lime.mol.Result.__init__ gets the defaults nout=1, t0=0.0 (and dt from `times` when only
    `times` is given, lime/superoperator.py:527) because lime/mol.py:92 does arithmetic on None
    for every density-matrix solver (lime/oqs.py:449,1670,1780).
"""
import sys
import types
import importlib

REFERENCE_ROOT = '/root/reference'


class _Anything(types.ModuleType):
    """module stub: any attribute is another stub, callable, iterable-safe"""

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        child = _Anything(self.__name__ + '.' + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + '()')

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _Anything(self.__name__ + '[]')


_STUBS = ['proplot', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.cm',
          'matplotlib.colors', 'matplotlib.collections', 'matplotlib.lines',
          'matplotlib.ticker', 'matplotlib.patches', 'matplotlib.animation',
          'mpl_toolkits', 'mpl_toolkits.mplot3d', 'mpl_toolkits.axes_grid1',
          'mpl_toolkits.axes_grid1.inset_locator']

_loaded = None


def available():
    import os
    return os.path.isdir(REFERENCE_ROOT + '/lime')


def load():
    """return the reference `lime` package (namespace-style, __init__ skipped)"""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    import numpy as np

    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Anything(name)
    if 'opt_einsum' not in sys.modules:
        try:
            importlib.import_module('opt_einsum')
        except Exception:
            oe = types.ModuleType('opt_einsum')
            oe.contract = lambda *a, **k: np.einsum(*a, optimize=True)
            sys.modules['opt_einsum'] = oe

    # synthetic package object: skips lime/__init__.py (which imports plotting)
    pkg = types.ModuleType('lime')
    pkg.__path__ = [REFERENCE_ROOT + '/lime']
    sys.modules['lime'] = pkg

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        mol = importlib.import_module('lime.mol')
        _orig_init = mol.Result.__init__

        def _init(self, description=None, psi0=None, rho0=None, dt=None,
                  Nt=None, times=None, t0=None, nout=None):
            if nout is None:
                nout = 1
            if t0 is None:
                t0 = 0.0
            if Nt is None and times is not None:
                Nt = len(times)
            if dt is None and times is not None:     # superoperator.Lindblad_solver.evolve: Result(times=tlist)
                dt = (times[1] - times[0]) if len(times) > 1 else 0.0
            _orig_init(self, description=description, psi0=psi0, rho0=rho0,
                       dt=dt, Nt=Nt, times=times, t0=t0, nout=nout)

        mol.Result.__init__ = _init
        for sub in ['phys', 'superoperator', 'liouville', 'oqs', 'heom.heom',
                    'cavity', 'signal.sos', 'correlation', 'units']:
            importlib.import_module('lime.' + sub)
    _loaded = pkg
    return pkg
