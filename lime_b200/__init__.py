"""
lime_b200 -- B200-native (sm_100a) replacement for the density-matrix hot path of
binggu56/lime: Redfield/Lindblad RK4 propagation, the HEOM hierarchy, and the
sum-over-states response functions behind 2DES/TPA spectra.

The modules mirror lime's layout and call signatures (lime.oqs -> lime_b200.oqs, ...);
the arithmetic runs in hand-written CUDA kernels behind the C ABI of liblime_b200.so
(include/lime_b200.h).  There is no CPU fallback.
"""
__version__ = '0.1.0'

from . import _lib            # noqa: F401


def library_path():
    return _lib.LIB_PATH
