"""Per-kernel device timing (CUDA events recorded inside the C library around
each entry point's device work).  The reference has only wall-clock timers
(partial_svd.py:261,290-291); bench.py uses this to measure
roofline.achieved live over the timed region."""
import ctypes

from ._lib import lib


def enable(on=True):
    lib.rl_profile_enable(1 if on else 0)


def reset():
    lib.rl_profile_reset()


def report():
    """{kind: {count, ms, bytes, flops, GBps, TFLOPs}} since the last reset()."""
    out = {}
    for kind in range(lib.rl_profile_kinds()):
        cnt, ms = ctypes.c_int64(0), ctypes.c_double(0)
        by, fl = ctypes.c_double(0), ctypes.c_double(0)
        lib.rl_profile_get(kind, ctypes.byref(cnt), ctypes.byref(ms), ctypes.byref(by), ctypes.byref(fl))
        if cnt.value == 0:
            continue
        name = lib.rl_profile_name(kind).decode()
        out[name] = {
            'count': cnt.value, 'ms': ms.value, 'bytes': by.value, 'flops': fl.value,
            'GBps': by.value / ms.value / 1e6 if ms.value > 0 else 0.0,
            'TFLOPs': fl.value / ms.value / 1e9 if ms.value > 0 else 0.0,
        }
    return out
