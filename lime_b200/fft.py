"""
Spectral post-processing with the signatures of lime/fft.py: `fft`, `ifft`, `fft2`
(lime/fft.py:15-110: NumPy FFT + shift + grid scaling + phase of the first grid point) and the
discrete Fourier transforms at chosen momenta `dft`, `dft2` (lime/fft.py:112-137).

* `dft` / `dft2` are O(N K) / O(N_x N_y K_x K_y) Python loops in lime; they are separable, so here they are
  complex GEMMs on the FP64 tensor cores (limeb200_zgemm): g = E_x f^T E_y^T with E[k, n] = exp(-i k x_n).
* `fft`, `ifft`, `fft2` call the FFT library on the device (torch.fft = cuFFT -- a plain library transform, like a
  library GEMM) and apply lime's shift / scale / phase epilogue on the device; there is no FFT kernel of ours.
lime's `dft` also opens a matplotlib figure (lime/fft.py:122-123); that side effect is not reproduced.
"""
import numpy as np
import torch

from . import engine
from . import _dev


def _to_dev(f):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(f, dtype=np.complex128))).to(_dev.device())


def fft(f, x=None, axis=-1, **kwargs):
    """g(w) = int dt f(t) exp(-i w t) on the grid x; returns (g, freq); lime/fft.py:15-58.
    (lime's phase factor only broadcasts for the last axis -- any other `axis` raises there; here the phase is
    applied along `axis`.)"""
    nx = np.asarray(f).shape[axis]
    if x is None:
        x = np.arange(nx)
    dx = x[1] - x[0]
    g = torch.fft.fft(_to_dev(f), dim=axis, **kwargs)
    g = torch.fft.fftshift(g, dim=(axis,)) * dx
    freq = 2. * np.pi * np.fft.fftshift(np.fft.fftfreq(nx, d=dx))
    phase = torch.from_numpy(np.exp(-1j * freq * x[0])).to(g.device)
    shape = [1] * g.dim()
    shape[axis] = -1
    return (g * phase.reshape(shape)).cpu().numpy(), freq


def ifft(f, x=None, axis=-1):
    """g = int dt f(t) exp(i w t); returns (g, freq); lime/fft.py:61-86 (ifftshift over ALL axes, as in lime)"""
    nx = np.asarray(f).shape[axis]
    if x is None:
        x = np.arange(nx)
    dx = x[1] - x[0]
    g = torch.fft.ifftshift(torch.fft.ifft(_to_dev(f), dim=axis))
    g = g * dx / 2. / np.pi * len(x)
    freq = 2. * np.pi * np.fft.ifftshift(np.fft.fftfreq(nx, d=dx))
    return g.cpu().numpy() * np.exp(1j * freq * x[0]), freq


def fft2(f, dx=1, dy=1):
    """2-D transform; returns (freqx, freqy, g); lime/fft.py:88-110 (freqy is built from nx, as in lime)"""
    nx, ny = np.asarray(f).shape
    g = torch.fft.fftshift(torch.fft.fft2(_to_dev(f))) * dx * dy
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
