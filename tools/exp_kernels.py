"""One-off A/B experiments (SpMM CTA width, Gram L2 prefetch distance)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import raleigh_b200 as rb
from raleigh_b200._lib import lib, check
from raleigh_b200 import device as dev
from microbench import timeit
from run_c4 import lap3d_slab
n, m = 2097152, 32
X, Y = rb.Vectors(n, m), rb.Vectors(n, m); X.fill_random_device(1); Y.fill_random_device(2)
g = dev.Buffer(m * m * 8); wsb = lib.rl_gram_ws_bytes(1, m, m, n); ws = dev.Buffer(wsb)
f = lambda: check(lib.rl_gram(1, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.ptr, ws.ptr, wsb, dev.stream()))
for w8 in (0, 1, 0, 1):
    lib.rl_debug_set_gram_simt(w8 << 20)
    wsb = lib.rl_gram_ws_bytes(1, m, m, n); ws = dev.Buffer(wsb)
    f = lambda: check(lib.rl_gram(1, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.ptr, ws.ptr, wsb, dev.stream()))
    ms, best = timeit(f)
    print(json.dumps({'exp': 'gram 8 warps/CTA = %d' % w8, 'ms': round(ms, 4), 'GBps': round(2.0 * n * m * 8 / ms / 1e6)}), flush=True)
import oracle
xs = X.data()[:, :100000]; ys = Y.data()[:, :100000]
Xs, Ys = rb.Vectors(xs.copy()), rb.Vectors(ys.copy())
lib.rl_debug_set_gram_simt(1 << 20)
err = abs(Xs.dot(Ys) - ys @ xs.T).max()
print(json.dumps({'exp': 'gram 8-warp correctness', 'max_abs_err': float(err)}))
lib.rl_debug_set_gram_simt(0)
sys.exit(0)
for N in (128, 256):
    nn = N ** 3
    A = rb.SparseSymmetricMatrix(lap3d_slab(N, 0, nn))
    Xs, Ys = rb.Vectors(nn, m), rb.Vectors(nn, m); Xs.fill_random_device(3)
    by = A.nnz() * 12.0 + (nn + 1) * 8.0 + 2.0 * nn * m * 8
    for wps in (4, 16):
        lib.rl_debug_set_spmm_warps(wps)
        ms, best = timeit(lambda: A.apply(Xs, Ys), reps=5)
        print(json.dumps({'exp': 'spmm lap3d %d^3 m=%d warps/CTA=%d' % (N, m, wps), 'ms': round(ms, 4), 'GBps': round(by / ms / 1e6)}), flush=True)
    del A, Xs, Ys
