"""Gram / block update at block sizes beyond one 32 x 32 tile (config 4 runs with m = k = 120), fp64.
    python tools/time_gram120.py [--rows 2097152]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raleigh_b200._lib import lib, check
from raleigh_b200 import device as dev

ap = argparse.ArgumentParser()
ap.add_argument('--rows', type=int, default=2097152)
args = ap.parse_args()
n = args.rows
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for m in (48, 64, 96, 120):
    ld = (n + 31) // 32 * 32
    X = torch.randn(m, ld, dtype=torch.float64, device='cuda')
    Y = torch.randn(m, ld, dtype=torch.float64, device='cuda')
    G = torch.zeros(m, m, dtype=torch.float64, device='cuda')
    wsb = lib.rl_gram_ws_bytes(1, m, m, n)
    ws = torch.zeros(wsb // 8 + 8, dtype=torch.float64, device='cuda')
    st = dev.stream()
    out = {'rows': n, 'block': m}
    for knob, name in ((0, 'tile_major'), (1, 'chunk_major')):
        lib.rl_debug_set_knob(6, knob)
        for same, tag in ((False, 'xy'), (True, 'xx')):
            O = X if same else Y
            ms = timeit(lambda: check(lib.rl_gram(1, X.data_ptr(), ld, m, O.data_ptr(), ld, m, n, G.data_ptr(), ws.data_ptr(), wsb, st)))
            out['gram_%s_%s_ms' % (tag, name)] = round(ms, 4)
            out['gram_%s_%s_TFLOPs' % (tag, name)] = round(2.0 * n * m * m / ms / 1e9, 2)
    lib.rl_debug_set_knob(6, 0)
    for tma in (1, 2):                       # TMA-fed ring kernel forced on the multi-tile product
        lib.rl_debug_set_knob(0, tma)
        wsb2 = lib.rl_gram_ws_bytes(1, m, m, n)
        ws2 = torch.zeros(wsb2 // 8 + 8, dtype=torch.float64, device='cuda')
        for same, tag in ((False, 'xy'), (True, 'xx')):
            O = X if same else Y
            ms = timeit(lambda: check(lib.rl_gram(1, X.data_ptr(), ld, m, O.data_ptr(), ld, m, n, G.data_ptr(), ws2.data_ptr(), wsb2, st)))
            out['gram_%s_tma%d_TFLOPs' % (tag, tma)] = round(2.0 * n * m * m / ms / 1e9, 2)
    lib.rl_debug_set_knob(0, 0)
    ref = (Y[:, :n] @ X[:, :n].T)
    check(lib.rl_gram(1, X.data_ptr(), ld, m, Y.data_ptr(), ld, m, n, G.data_ptr(), ws.data_ptr(), wsb, st))
    torch.cuda.synchronize()
    out['max_rel_err'] = float((G - ref).abs().max() / ref.abs().max())
    print(json.dumps(out), flush=True)
