"""Device plumbing: PyTorch provides device memory (caching allocator), the
current CUDA stream and torch.distributed; everything numerical goes through
the C ABI in _lib.py.  No CPU fallback: constructing anything without a CUDA
device raises RuntimeError, as the reference's `arch='gpu!'` does
(dense_matrix.py:21-23).
"""
import ctypes

import numpy
import torch

from . import _lib

_checked = False


def require_cuda():
    global _checked
    if _checked:
        return
    if not torch.cuda.is_available():
        raise RuntimeError('raleigh_b200: no CUDA device available (cannot use GPU; there is no CPU fallback)')
    major, minor = torch.cuda.get_device_capability()
    if major != 10:
        raise RuntimeError('raleigh_b200 is built for sm_100a (B200) only; found compute capability %d.%d'
                           % (major, minor))
    _checked = True


_device_index = None


def stream():
    """cudaStream_t of torch's current stream, as an int for ctypes.  Called once per C-ABI call (about 80
    times per solver iteration): the raw query (0.3 us) instead of torch.cuda.current_stream() (15 us, profiled
    at 15 ms per config-2 solve).  One process drives one device (dist.py), so its index is looked up once."""
    global _device_index
    if _device_index is None:
        _device_index = torch.cuda.current_device()
    try:
        return torch._C._cuda_getCurrentRawStream(_device_index)
    except AttributeError:          # private API moved: fall back to the public one
        return torch.cuda.current_stream().cuda_stream


def synchronize():
    """cuda_wrap.synchronize (cuda_wrap.py:141); called by the reference's
    interfaces before reading timers (partial_svd.py:288-289)."""
    _lib.check(_lib.lib.rl_sync_device())
    return 0


class Buffer:
    """Owner of one device allocation; shared by shallow copies / references
    (the reference's _Data, dense_cublas.py:801-811, minus the cudaFree that
    raises at interpreter exit)."""

    __slots__ = ('tensor', 'ptr', 'nbytes', 'version')

    def __init__(self, nbytes, zero=False):
        require_cuda()
        nbytes = max(int(nbytes), 16)
        if zero:
            self.tensor = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
        else:
            self.tensor = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
        self.ptr = self.tensor.data_ptr()
        self.nbytes = nbytes
        self.version = 0      # bumped by every write through a Vectors/Matrix view


def host_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def padded_ld(n, itemsize):
    """Leading dimension: rows start on 128-byte boundaries so every kernel can
    use 128-bit accesses and TMA (16-byte stride rule) on any window."""
    q = 128 // itemsize
    return max(q, (int(n) + q - 1) // q * q)


def upload_2d(dst_ptr, ld_bytes, a):
    """H2D of a 2-D C-contiguous host array into a pitched device block."""
    m, n = a.shape
    if m == 0 or n == 0:
        return
    row = n * a.itemsize
    if ld_bytes == row:
        _lib.check(_lib.lib.rl_h2d(dst_ptr, host_ptr(a), m * row, stream()))
    else:
        _lib.check(_lib.lib.rl_h2d_2d(dst_ptr, ld_bytes, host_ptr(a), row, row, m, stream()))
    # pageable source: make sure the DMA has consumed it before the caller mutates it
    _lib.check(_lib.lib.rl_sync_stream(stream()))


def download_2d(src_ptr, ld_bytes, m, n, dtype):
    out = numpy.empty((m, n), dtype=dtype)
    if m == 0 or n == 0:
        return out
    row = n * out.itemsize
    if ld_bytes == row:
        _lib.check(_lib.lib.rl_d2h(host_ptr(out), src_ptr, m * row, stream()))
    else:
        _lib.check(_lib.lib.rl_d2h_2d(host_ptr(out), row, src_ptr, ld_bytes, row, m, stream()))
    return out
