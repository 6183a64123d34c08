"""Stand-in for raleigh/algebra/cuda_wrap.py: the handful of runtime calls the
reference's interfaces make on the object returned by ``AMatrix.gpu()``
(``synchronize()`` at partial_svd.py:288-289, tests_algebra.py:126)."""
import ctypes

from ._lib import lib, check
from .device import synchronize, require_cuda  # noqa: F401

numDevices = ctypes.c_int(0)


def getDeviceCount(ptr=None):
    n = ctypes.c_int(0)
    rc = lib.rl_device_count(ctypes.byref(n))
    numDevices.value = n.value
    if ptr is not None:
        ptr._obj.value = n.value
    return rc


def device_info(device=0):
    sm, maj, mnr = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    l2, mem = ctypes.c_size_t(0), ctypes.c_size_t(0)
    check(lib.rl_device_info(device, ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr), ctypes.byref(l2),
                             ctypes.byref(mem)))
    return dict(sm_count=sm.value, cc=(maj.value, mnr.value), l2_bytes=l2.value, total_mem=mem.value)


def launch_count():
    return int(lib.rl_launch_count())
