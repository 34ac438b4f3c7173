"""
Device plumbing: PyTorch owns device memory and streams, nothing else.  complex128 numpy
arrays go to the device as torch.complex128 tensors (same interleaved bytes as double2).
"""
import ctypes as C
import numpy as np
import torch

from ._lib import LimeB200Error


def device(index=None):
    if not torch.cuda.is_available():
        raise LimeB200Error('no CUDA device visible: lime_b200 has no CPU fallback')
    if index is None:
        index = torch.cuda.current_device()
    return torch.device('cuda', index)


def to_dev(a, dtype=np.complex128, dev=None, pinned=False):
    """numpy -> contiguous device tensor of the given numpy dtype"""
    a = np.ascontiguousarray(a, dtype=dtype)
    t = torch.from_numpy(a)
    if pinned:
        t = t.pin_memory()
    return t.to(device() if dev is None else dev, non_blocking=pinned)


def h2d(a, dtype=np.complex128, dev=None):
    """host -> device.  A torch CPU tensor is copied as it is (asynchronously when it is
    pinned); anything else goes through numpy (pageable)."""
    if isinstance(a, torch.Tensor):
        assert a.device.type == 'cpu' and a.is_contiguous()
        return a.to(device() if dev is None else dev, non_blocking=a.is_pinned())
    return to_dev(a, dtype, dev)


_pinned_pool = {}


def d2h(t, pinned=False):
    """device -> host numpy array; pinned=True stages through a cached page-locked buffer
    (the returned array is a view of it and is overwritten by the next call of that shape)"""
    if t is None:
        return None
    if not pinned:
        return t.cpu().numpy()
    key = (tuple(t.shape), t.dtype)
    buf = _pinned_pool.get(key)
    if buf is None:
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        _pinned_pool[key] = buf
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()      # the copy runs on the SOURCE device's current stream
    return buf.numpy()


def d2h_owned(t):
    """device -> host numpy array that the CALLER owns: staged through a page-locked buffer from torch's caching host
    allocator (a block is recycled only after the returned array -- which keeps the buffer alive -- is dropped), so the
    copy runs at pinned PCIe speed without the aliasing of d2h(pinned=True)'s shape-keyed pool"""
    if t is None:
        return None
    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return buf.numpy()


def empty(shape, dtype=torch.complex128, dev=None):
    return torch.empty(shape, dtype=dtype, device=device() if dev is None else dev)


def ptr(t):
    """raw device pointer of a tensor (or None)"""
    if t is None:
        return None
    assert t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def stream_ptr(dev=None):
    """the current stream of device `dev` (default: the current device) as a cudaStream_t"""
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def as_c128(a):
    """dense C-contiguous complex128 numpy array from ndarray / scipy.sparse / np.matrix"""
    if hasattr(a, 'toarray'):
        a = a.toarray()
    return np.ascontiguousarray(np.asarray(a), dtype=np.complex128)
