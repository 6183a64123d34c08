"""Per-kernel timing of the C-ABI entry points with CUDA events (L2 flushed between
repetitions).  Prints one JSON line per (kernel, shape): ms, algorithmic GB/s
(SURVEY.md section 8d formulas), fraction of the measured HBM peak.

    python tools/microbench.py [--quick] [--only gram,update,...] [--out file.jsonl]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raleigh_b200 as rb  # noqa: E402
from raleigh_b200._lib import lib, check  # noqa: E402
from raleigh_b200 import device as dev  # noqa: E402


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        return 6650.0


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
    _flush.zero_()


REPS = [10]


def timeit(fn, reps=None, warm=3, flush=True):
    reps = reps or REPS[0]
    if REPS[0] < 3:
        warm = 1
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def emit(out, name, shape, ms, best, byts, flops=0.0, note=''):
    rec = {'kernel': name, 'shape': shape, 'ms_median': round(ms, 5), 'ms_best': round(best, 5),
           'GBps': round(byts / ms / 1e6, 1), 'frac_hbm': round(byts / ms / 1e6 / peak_gbs(), 3),
           'TFLOPs': round(flops / ms / 1e9, 2), 'note': note}
    print(json.dumps(rec), flush=True)
    if out:
        out.write(json.dumps(rec) + '\n')
        out.flush()


def vec_suite(out, n, m, dtype, only):
    w = np.dtype(dtype).itemsize
    code = 1 if w == 8 else 0
    X, Y, W = rb.Vectors(n, m, dtype), rb.Vectors(n, m, dtype), rb.Vectors(n, m, dtype)
    X.fill_random_device(1)
    Y.fill_random_device(2)
    blk = n * m * w
    st = dev.stream
    shape = 'n=%d,m=%d,%s' % (n, m, np.dtype(dtype).name)
    q = torch.randn(m, m, dtype=torch.float64 if w == 8 else torch.float32, device='cuda') * 0.1
    s = torch.rand(m, dtype=q.dtype, device='cuda') + 0.5
    g = torch.empty(m * m, dtype=q.dtype, device='cuda')
    if 'gram' in only:
        wsb = lib.rl_gram_ws_bytes(code, m, m, n)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device='cuda')
        f = lambda: check(lib.rl_gram(code, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.data_ptr(), ws.data_ptr(), wsb, st()))
        ms, best = timeit(f)
        emit(out, 'gram_xy', shape, ms, best, 2 * blk, 2.0 * n * m * m)
        f = lambda: check(lib.rl_gram(code, X._wptr(), X._ld, m, X._wptr(), X._ld, m, n, g.data_ptr(), ws.data_ptr(), wsb, st()))
        ms, best = timeit(f)
        emit(out, 'gram_xx', shape, ms, best, blk, 2.0 * n * m * m)
        if w == 8 and m > 16:
            lib.rl_debug_set_gram_simt(2)
            wsb3 = lib.rl_gram_ws_bytes(code, m, m, n)
            ws3 = torch.empty(max(wsb3, 16), dtype=torch.uint8, device='cuda')
            f = lambda: check(lib.rl_gram(code, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.data_ptr(), ws3.data_ptr(), wsb3, st()))
            ms, best = timeit(f)
            lib.rl_debug_set_gram_simt(0)
            emit(out, 'gram_xy_wide', shape, ms, best, 2 * blk, 2.0 * n * m * m, 'one warp per 32x32 tile (A/B)')
        if w == 8:
            lib.rl_debug_set_gram_simt(1)
            wsb2 = lib.rl_gram_ws_bytes(code, m, m, n)
            ws2 = torch.empty(max(wsb2, 16), dtype=torch.uint8, device='cuda')
            f = lambda: check(lib.rl_gram(code, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.data_ptr(), ws2.data_ptr(), wsb2, st()))
            ms, best = timeit(f)
            lib.rl_debug_set_gram_simt(0)
            emit(out, 'gram_xy_simt', shape, ms, best, 2 * blk, 2.0 * n * m * m, 'FMA-pipe variant (A/B)')
    if 'update' in only:
        f = lambda: check(lib.rl_update(code, W._wptr(), W._ld, m, X._wptr(), X._ld, m, q.data_ptr(), m, 1, 1.0, 0.0, n, st()))
        ms, best = timeit(f)
        emit(out, 'update_beta0', shape, ms, best, 2 * blk, 2.0 * n * m * m)
        f = lambda: check(lib.rl_update(code, W._wptr(), W._ld, m, X._wptr(), X._ld, m, q.data_ptr(), m, 1, -1.0, 1.0, n, st()))
        ms, best = timeit(f)
        emit(out, 'update_beta1', shape, ms, best, 3 * blk, 2.0 * n * m * m)
        if w == 8:
            lib.rl_debug_set_update_fma(1)
            f = lambda: check(lib.rl_update(code, W._wptr(), W._ld, m, X._wptr(), X._ld, m, q.data_ptr(), m, 1, 1.0, 0.0, n, st()))
            ms, best = timeit(f)
            lib.rl_debug_set_update_fma(0)
            emit(out, 'update_beta0_fma', shape, ms, best, 2 * blk, 2.0 * n * m * m, 'FMA-pipe variant (A/B)')
    if 'blas1' in only:
        f = lambda: check(lib.rl_axpy(code, W._wptr(), W._ld, X._wptr(), X._ld, m, n, 0.5, st()))
        ms, best = timeit(f)
        emit(out, 'axpy', shape, ms, best, 3 * blk)
        f = lambda: check(lib.rl_axpy_diag(code, W._wptr(), W._ld, X._wptr(), X._ld, m, n, s.data_ptr(), st()))
        ms, best = timeit(f)
        emit(out, 'axpy_diag', shape, ms, best, 3 * blk)
        f = lambda: check(lib.rl_scale(code, W._wptr(), W._ld, m, n, s.data_ptr(), 0, st()))
        ms, best = timeit(f)
        emit(out, 'scale', shape, ms, best, 2 * blk)
        f = lambda: check(lib.rl_copy(code, W._wptr(), W._ld, X._wptr(), X._ld, m, n, st()))
        ms, best = timeit(f)
        emit(out, 'copy', shape, ms, best, 2 * blk)
        wsb = lib.rl_dots_ws_bytes(code, m, n)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device='cuda')
        f = lambda: check(lib.rl_dots(code, X._wptr(), X._ld, Y._wptr(), Y._ld, m, n, g.data_ptr(), ws.data_ptr(), wsb, st()))
        ms, best = timeit(f)
        emit(out, 'dots', shape, ms, best, 2 * blk)
        d = torch.rand(n, dtype=q.dtype, device='cuda')
        f = lambda: check(lib.rl_diag_mul(code, W._wptr(), W._ld, X._wptr(), X._ld, m, n, d.data_ptr(), st()))
        ms, best = timeit(f)
        emit(out, 'jacobi', shape, ms, best, 2 * blk + n * w)
        a = torch.empty(n * m, dtype=q.dtype, device='cuda')
        b = torch.empty(n * m, dtype=q.dtype, device='cuda')
        ms, best = timeit(lambda: b.copy_(a))
        emit(out, 'torch_copy(ref)', shape, ms, best, 2 * blk, note='driver peak recipe: b.copy_(a)')


def spmm_suite(out, name, A, m, dtype):
    w = np.dtype(dtype).itemsize
    n = A.shape[0]
    op = rb.SparseSymmetricMatrix(A.astype(dtype))
    X, Y = rb.Vectors(n, m, dtype), rb.Vectors(n, m, dtype)
    X.fill_random_device(3)
    nnz = op.nnz()
    byts = nnz * (w + 4) + (n + 1) * 8 + 2 * n * m * w
    ms, best = timeit(lambda: op.apply(X, Y))
    emit(out, 'spmm', '%s,n=%d,nnz=%d,m=%d,%s' % (name, n, nnz, m, np.dtype(dtype).name), ms, best, byts,
         2.0 * nnz * m)


def gemm_suite(out, M, N, k, reps=5):
    a = (torch.randn(M, N, dtype=torch.float32, device='cuda')).cpu().numpy()
    A = rb.Matrix(a)
    x = rb.Vectors(N, k, np.float32)
    y = rb.Vectors(M, k, np.float32)
    x.fill_random_device(4)
    fl = 2.0 * M * N * k
    by = M * N * 4 + k * (M + N) * 4
    shape = 'A=%dx%d,k=%d,f32' % (M, N, k)
    A.apply(x, y)
    ms, best = timeit(lambda: A.apply(x, y), reps=reps, warm=1)
    emit(out, 'dense_apply_tc', shape, ms, best, by, fl, 'lo parts split in shared memory: A streamed once')
    ms, best = timeit(lambda: A.apply(y, x, transp=True), reps=reps, warm=1)
    emit(out, 'dense_apply_tc_T', shape, ms, best, by, fl, 'lo parts split in shared memory: A streamed once')
    f = lambda: check(lib.rl_dense_apply(0, A._aptr(), A._ld, M, N, x._wptr(), x._ld, y._wptr(), y._ld, k, 0, 1.0, 0.0, dev.stream()))
    ms, best = timeit(f, reps=2, warm=1)
    emit(out, 'dense_apply_simt', shape, ms, best, by, fl)
    at = torch.as_tensor(a, device='cuda')
    xt = torch.randn(k, N, dtype=torch.float32, device='cuda')
    ms, best = timeit(lambda: torch.matmul(xt, at.T), reps=reps, warm=1)
    emit(out, 'torch_matmul_fp32(ref)', shape, ms, best, by, fl, 'cuBLAS SGEMM incumbent')
    # accuracy of the 3xTF32 split against an fp64 product
    A.apply(x, y)
    ref = (torch.as_tensor(x.data(), device='cuda').double() @ at.double().T)
    err = ((torch.as_tensor(y.data(), device='cuda').double() - ref).abs().max() / ref.abs().max()).item()
    cub = ((torch.matmul(torch.as_tensor(x.data(), device='cuda'), at.T).double() - ref).abs().max() / ref.abs().max()).item()
    print(json.dumps({'kernel': 'dense_apply_tc', 'shape': shape, 'max_rel_err_vs_fp64': err, 'cublas_sgemm_err': cub}), flush=True)
    for kk in (1, 8):
        xs = rb.Vectors(N, kk, np.float32); ys = rb.Vectors(M, kk, np.float32); xs.fill_random_device(5)
        ms, best = timeit(lambda: A.apply(xs, ys), reps=3, warm=1)
        emit(out, 'dense_apply(k=%d)' % kk, shape, ms, best, M * N * 4.0, 2.0 * M * N * kk)


def eig_suite(out):
    for p in (32, 64, 128, 240):
        rng = np.random.RandomState(p)
        b = rng.randn(p, p)
        a = torch.tensor(b @ b.T, dtype=torch.float64, device='cuda')
        wsb = lib.rl_syevj_ws_bytes(p)
        ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
        wv = torch.empty(p, dtype=torch.float64, device='cuda')
        work = a.clone()

        def f():
            work.copy_(a)
            check(lib.rl_syevj(work.data_ptr(), p, wv.data_ptr(), ws.data_ptr(), wsb, None, dev.stream()))
        ms, best = timeit(f, reps=3, warm=1, flush=False)
        ref = np.linalg.eigvalsh(b @ b.T)
        err = float(np.max(np.abs(wv.cpu().numpy() - ref)) / ref[-1])
        emit(out, 'syevj', 'p=%d' % p, ms, best, 0.0, note='rel eigenvalue error %.1e' % err)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--only', default='gram,update,blas1,spmm,gemm,eig')
    ap.add_argument('--out', default='')
    ap.add_argument('--shape', default='', help='n,m: run the vector suite on this single fp64 shape only')
    ap.add_argument('--reps', type=int, default=10)
    args = ap.parse_args()
    only = set(args.only.split(','))
    REPS[0] = args.reps
    out = open(args.out, 'w') if args.out else None
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from oracle import algebra_np as K
    shapes = [(32768, 16), (140874, 32), (2097152, 32)]
    if args.shape:
        n_, m_ = (int(t) for t in args.shape.split(','))
        shapes = [(n_, m_)]
        args.quick = True
    if not args.quick:
        shapes += [(2097152, 16), (2097152, 64), (2097152, 120), (8388608, 32)]
    for (n, m) in shapes:
        vec_suite(out, n, m, np.float64, only)
    if not args.quick:
        vec_suite(out, 2097152, 32, np.float32, only)
    if 'spmm' in only:
        spmm_suite(out, 'lap3d_32', K.lap3d_csr(32, 32, 32), 16, np.float64)
        spmm_suite(out, 'lap3d_128', K.lap3d_csr(128, 128, 128), 32, np.float64)
        from tests_common import spd_c3_like
        offs = tuple(sorted(set([1, 2, 3, 4, 5, 6, 440, 441, 442, 443, 444, 445, 446, 2656, 2657, 2658, 2659, 2660,
                                 2661, 2662, 2214, 2215, 2216, 2217, 3100, 3101, 3102])))
        spmm_suite(out, 'c3like_55nnz', spd_c3_like(140874, offsets=offs), 32, np.float64)
    if 'gemm' in only:
        gemm_suite(out, 12000, 39375, 128)
    if 'eig' in only:
        eig_suite(out)
