"""
Spectral post-processing with the signatures of lime/fft.py: `fft`, `ifft`, `fft2`
(lime/fft.py:15-110: NumPy FFT + shift + grid scaling + phase of the first grid point) and the
discrete Fourier transforms at chosen momenta `dft`, `dft2` (lime/fft.py:112-137).

Everything here is a complex GEMM on the FP64 tensor cores (limeb200_zgemm); no FFT library is called.
* `fft`, `ifft`, `fft2`: the grids lime transforms are the t1 / t3 axes of response functions (a few hundred points),
  so the transform is applied as ONE product with the N x N Fourier matrix, and lime's whole epilogue -- fftshift /
  ifftshift of the output index, the dx (dx dy, dx N / 2 pi) scaling and the phase exp(-+ i w x_0) of the first grid
  point -- is folded into that matrix on the host (exact integer arithmetic k n mod N for the twiddle angles), so the
  device does a single pass and nothing is shifted or rescaled afterwards.  O(N^2) per transform instead of
  O(N log N): at N <= 4096 the tensor-core GEMM is launch-bound either way.
* `dft` / `dft2` are O(N K) / O(N_x N_y K_x K_y) Python loops in lime; they are separable, so they are GEMMs too:
  g = E_x f^T E_y^T with E[k, n] = exp(-i k x_n).
lime's `dft` also opens a matplotlib figure (lime/fft.py:122-123); that side effect is not reproduced.
"""
import numpy as np

from . import engine

_MATS = {}


def _fourier_matrix(n, sign, order, col_scale):
    """W[m, k'] = col_scale[k'] * exp(sign * 2 pi i * k(k') * m / n), k(k') = order[k'] (integer output frequencies)"""
    m = np.arange(n, dtype=np.int64)
    kn = (np.asarray(order, dtype=np.int64)[None, :] * m[:, None]) % n
    ang = 2.0 * np.pi * kn / n
    return np.ascontiguousarray((np.cos(ang) + sign * 1j * np.sin(ang)) * np.asarray(col_scale)[None, :])


def _apply_last_axis(f, W):
    """f [..., n] @ W [n, n] on the device"""
    f = np.ascontiguousarray(np.asarray(f, dtype=np.complex128))
    lead = f.shape[:-1]
    out = engine.zgemm(f.reshape(-1, f.shape[-1]), W)
    return out.cpu().numpy().reshape(lead + (W.shape[1],))


def fft(f, x=None, axis=-1, **kwargs):
    """g(w) = int dt f(t) exp(-i w t) on the grid x; returns (g, freq); lime/fft.py:15-58.
    (lime's phase factor only broadcasts for the last axis -- any other `axis` raises there; here the phase is
    applied along `axis`.)"""
    if kwargs:
        raise TypeError('fft: unsupported options %s (lime forwards them to numpy.fft.fft)' % sorted(kwargs))
    f = np.asarray(f)
    nx = f.shape[axis]
    if x is None:
        x = np.arange(nx)
    dx = x[1] - x[0]
    freq = 2. * np.pi * np.fft.fftshift(np.fft.fftfreq(nx, d=dx))
    order = np.fft.fftshift(np.arange(nx))                   # output slot k' holds numpy's frequency index order[k']
    W = _fourier_matrix(nx, -1.0, order, dx * np.exp(-1j * freq * x[0]))
    g = _apply_last_axis(np.moveaxis(f, axis, -1), W)
    return np.moveaxis(g, -1, axis), freq


def ifft(f, x=None, axis=-1):
    """g = int dt f(t) exp(i w t); returns (g, freq); lime/fft.py:61-86 (ifftshift over ALL axes, as in lime)"""
    f = np.asarray(f)
    nx = f.shape[axis]
    if x is None:
        x = np.arange(nx)
    dx = x[1] - x[0]
    freq = 2. * np.pi * np.fft.ifftshift(np.fft.fftfreq(nx, d=dx))
    order = np.fft.ifftshift(np.arange(nx))
    scale = (1.0 / nx) * dx / 2. / np.pi * len(x)
    ax = axis % f.ndim
    phase = np.exp(1j * freq * x[0]) if ax == f.ndim - 1 else np.ones(nx)
    W = _fourier_matrix(nx, +1.0, order, scale * phase)
    g = np.moveaxis(_apply_last_axis(np.moveaxis(f, axis, -1), W), -1, axis)
    other = tuple(a for a in range(f.ndim) if a != ax)
    if other:                                                 # lime shifts every axis; the others are pure permutations
        g = np.fft.ifftshift(g, axes=other)
    if ax != f.ndim - 1:                                      # and its phase broadcasts along the LAST axis
        g = g * np.exp(1j * freq * x[0])
    return g, freq


def fft2(f, dx=1, dy=1):
    """2-D transform; returns (freqx, freqy, g); lime/fft.py:88-110 (freqy is built from nx, as in lime)"""
    f = np.asarray(f)
    nx, ny = f.shape
    Wx = _fourier_matrix(nx, -1.0, np.fft.fftshift(np.arange(nx)), np.full(nx, dx * dy))      # [a, kx']
    Wy = _fourier_matrix(ny, -1.0, np.fft.fftshift(np.arange(ny)), np.ones(ny))               # [b, ky']
    fd = np.ascontiguousarray(f, dtype=np.complex128)
    g = engine.zgemm(np.ascontiguousarray(Wx.T), engine.zgemm(fd, Wy))                       # Wx^T f Wy
    freqx = 2. * np.pi * np.fft.fftshift(np.fft.fftfreq(nx, d=dx))
    freqy = 2. * np.pi * np.fft.fftshift(np.fft.fftfreq(nx, d=dy))
    return freqx, freqy, g.cpu().numpy()


def dft(x, f, k):
    """g[i] = sum_n f[n] exp(-i k_i x_n) dx; lime/fft.py:112-124"""
    x = np.asarray(x)
    dx = (x[1] - x[0]).real
    E = np.exp(-1j * np.outer(np.asarray(k), x))                               # [K, N]
    g = engine.zgemm(np.ascontiguousarray(E), np.ascontiguousarray(np.asarray(f, dtype=complex).reshape(-1, 1)))
    return g.cpu().numpy()[:, 0] * dx


def dft2(x, y, f, kx, ky):
    """g[i, j] = sum_{a,b} f[a, b] exp(-i kx_i x_b - i ky_j y_a) dx dy with X, Y = meshgrid(x, y), i.e. f is indexed
    [y, x]; lime/fft.py:126-137"""
    x, y = np.asarray(x), np.asarray(y)
    dx = x[1] - x[0]
    dy = y[1] - y[0]
    Ex = np.ascontiguousarray(np.exp(-1j * np.outer(np.asarray(kx), x)))       # [Kx, Nx]
    EyT = np.ascontiguousarray(np.exp(-1j * np.outer(y, np.asarray(ky))))      # [Ny, Ky]
    fT = np.ascontiguousarray(np.asarray(f, dtype=complex).T)                  # [Nx, Ny]
    g = engine.zgemm(Ex, engine.zgemm(fT, EyT))                                # [Kx, Ky]
    return g.cpu().numpy() * dx * dy
