"""Host <-> device copies of big PAGEABLE arrays: staged path (pinned double buffer + host threads) against the
driver's own pageable path (knob 7), and pinned memory for reference.
    python tools/time_copies.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raleigh_b200._lib import lib, check
from raleigh_b200 import device as dev

rows, cols = 65536, 4096                      # one config-5 chunk, 1.07 GB
a = np.random.rand(rows, cols).astype(np.float32)
ld = cols * 4 + 128                           # pitched destination
d = torch.empty(rows * ld, dtype=torch.uint8, device='cuda')
out = np.empty_like(a)
pinned = torch.empty((rows, cols), dtype=torch.float32, pin_memory=True)
pinned.numpy()[...] = a


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


res = {'bytes': a.nbytes}
for knob, name in ((0, 'staged'), (1, 'driver_pageable')):
    lib.rl_debug_set_knob(7, knob)
    s = t(lambda: (check(lib.rl_h2d_2d(d.data_ptr(), ld, dev.host_ptr(a), cols * 4, cols * 4, rows, dev.stream())),
                   check(lib.rl_sync_stream(dev.stream()))))
    res['h2d_%s_GBps' % name] = round(a.nbytes / s / 1e9, 2)
    s = t(lambda: check(lib.rl_d2h_2d(dev.host_ptr(out), cols * 4, d.data_ptr(), ld, cols * 4, rows, dev.stream())))
    res['d2h_%s_GBps' % name] = round(a.nbytes / s / 1e9, 2)
    assert np.array_equal(out, a), name
lib.rl_debug_set_knob(7, 0)
s = t(lambda: (check(lib.rl_h2d_2d(d.data_ptr(), ld, pinned.data_ptr(), cols * 4, cols * 4, rows, dev.stream())),
               check(lib.rl_sync_stream(dev.stream()))))
res['h2d_pinned_GBps'] = round(a.nbytes / s / 1e9, 2)
# contiguous 1-D path
flat = np.random.rand(300_000_000 // 8).astype(np.float64)
d1 = torch.empty(flat.nbytes, dtype=torch.uint8, device='cuda')
o1 = np.empty_like(flat)
s = t(lambda: (check(lib.rl_h2d(d1.data_ptr(), dev.host_ptr(flat), flat.nbytes, dev.stream())), check(lib.rl_sync_stream(dev.stream()))))
res['h2d_1d_staged_GBps'] = round(flat.nbytes / s / 1e9, 2)
s = t(lambda: check(lib.rl_d2h(dev.host_ptr(o1), d1.data_ptr(), flat.nbytes, dev.stream())))
res['d2h_1d_staged_GBps'] = round(flat.nbytes / s / 1e9, 2)
assert np.array_equal(o1, flat)
res['copy_threads'] = int(os.environ.get('RALEIGH_B200_COPY_THREADS', 0)) or 'auto'
print(json.dumps(res))
