"""r1e A/B sweep on one B200: TMA-fed Gram modes and footprint-clustered SpMM orders.

Every variant is first checked against the default kernel on the same inputs (Gram: relative
difference of the k x m result; SpMM: Y must be bit-identical, the per-row arithmetic does not
change), then timed with CUDA events, L2 flushed between repetitions.

    python tools/sweep_r1e.py [--out gpurun_out/sweep_r1e.jsonl] [--N 128] [--only edge,gram,spmm] [--window]

ncu captures of the variants (one launch each; `gram_reduce` included so that the C-ABI call can be
split into its two kernels):

    ncu --set full --clock-control none --import-source on \
        -k regex:"gram_tma|gram_dmma|gram_reduce|spmm_kernel|spmm_win" -o gpurun_out/prof -f \
        python tools/sweep_r1e.py --only ncu
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --title "..." > profiles/xxx.md
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import raleigh_b200 as rb  # noqa: E402
from raleigh_b200._lib import lib, check  # noqa: E402
from raleigh_b200 import device as dev  # noqa: E402
from raleigh_b200 import sparse as rsp  # noqa: E402
from microbench import timeit, peak_gbs  # noqa: E402
from run_c4 import lap3d_slab  # noqa: E402

KNOB_GRAM_TMA, KNOB_SPMM_CARVEOUT, KNOB_SPMM_WPS, KNOB_SPMM_PF, KNOB_GRAM_INTERLEAVE, KNOB_GRAM_WAVES = 0, 1, 2, 3, 4, 5
GRAM_MODES = [(-1, 0), (0, 0), (3, 4), (3, 8), (3, 16), (3, 32)]      # (TMA mode, CTAs per SM slot)
OUT = [None]


def emit(**rec):
    line = json.dumps(rec)
    print(line, flush=True)
    if OUT[0]:
        OUT[0].write(line + '\n')
        OUT[0].flush()


def gram_sweep(n, shapes, reps):
    mmax = max(max(s) for s in shapes)
    X, Y = rb.Vectors(n, mmax), rb.Vectors(n, mmax)
    X.fill_random_device(1)
    Y.fill_random_device(2)
    st = dev.stream
    for (m, k) in shapes:
        g = torch.empty(m * k, dtype=torch.float64, device='cuda')
        ref = {}
        for same in (False, True):
            if same and m != k:
                continue
            O = X if same else Y
            for (mode, il) in GRAM_MODES:
                lib.rl_debug_set_knob(KNOB_GRAM_TMA, mode)
                lib.rl_debug_set_knob(KNOB_GRAM_WAVES, il)
                wsb = lib.rl_gram_ws_bytes(1, m, k, n)
                ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device='cuda')
                f = lambda: check(lib.rl_gram(1, X._wptr(), X._ld, m, O._wptr(), O._ld, k, n, g.data_ptr(),
                                              ws.data_ptr(), wsb, st()))
                g.zero_()
                f()
                torch.cuda.synchronize()
                res = g.clone()
                if mode == -1:
                    ref[same] = res
                    err = 0.0
                else:
                    err = float((res - ref[same]).abs().max() / ref[same].abs().max())
                ms, best = timeit(f, reps=reps)
                byts = (m if same else m + k) * n * 8.0
                emit(exp='gram', n=n, m=m, k=k, same=same, mode=mode, waves=il, ms=round(ms, 5), ms_best=round(best, 5),
                     GBps=round(byts / ms / 1e6, 1), frac_hbm=round(byts / ms / 1e6 / peak_gbs(), 3),
                     TFLOPs=round(2.0 * n * m * k / ms / 1e9, 2), rel_diff_vs_default=err)
        lib.rl_debug_set_knob(KNOB_GRAM_TMA, 0)
        lib.rl_debug_set_knob(KNOB_GRAM_WAVES, 0)


def ncu_pass(N):
    """One launch of every variant worth an ncu capture (run under `ncu -k regex:gram_|spmm_`)."""
    n, m = 2097152, 32
    X, Y = rb.Vectors(n, m), rb.Vectors(n, m)
    X.fill_random_device(1)
    Y.fill_random_device(2)
    g = torch.empty(m * m, dtype=torch.float64, device='cuda')
    wsb = lib.rl_gram_ws_bytes(1, m, m, n)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device='cuda')
    for (mode, il) in ((3, 8),):
        lib.rl_debug_set_knob(KNOB_GRAM_TMA, mode)
        lib.rl_debug_set_knob(KNOB_GRAM_WAVES, il)
        wsb = lib.rl_gram_ws_bytes(1, m, m, n)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device='cuda')
        check(lib.rl_gram(1, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.data_ptr(), ws.data_ptr(), wsb, dev.stream()))
    lib.rl_debug_set_knob(KNOB_GRAM_TMA, 0)
    torch.cuda.synchronize()
    A = lap3d_slab(N, 0, N ** 3)
    rsp.SPMM_CLUSTER_WARPS = 0
    op = rb.SparseSymmetricMatrix(A)
    n = A.shape[0]
    indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    ip, ix, va = _priv(op, 'indptr'), _priv(op, 'indices'), _priv(op, 'values')
    X, Y = rb.Vectors(n, m), rb.Vectors(n, m)
    X.fill_random_device(3)
    for group, wps, pf in ((0, 24, 1), (0, 24, 3), (0, 24, 7)):
        ob = None
        lib.rl_debug_set_knob(KNOB_SPMM_WPS, wps)
        lib.rl_debug_set_knob(KNOB_SPMM_PF, pf)
        if group:
            ob = rsp._to_device(rsp.cluster_runs(indptr, indices, n, group)[0])
        check(lib.rl_csr_spmm_ex(1, n, op.nnz(), ip.ptr, ix.ptr, va.ptr, X._wptr(), X._ld, Y._wptr(), Y._ld, m, 0, None,
                                 ob.ptr if ob else None, group if group else 4, dev.stream()))
    torch.cuda.synchronize()


def gemm_sweep(M, N, k, reps):
    """tcgen05 dense apply: default (lo copies streamed from HBM) vs the experimental in-kernel lo split
    (knob 8), both orientations; the split variant must reproduce the default to fp32 rounding."""
    KNOB_GEMM_INSPLIT = 8
    a = torch.randn(M, N, dtype=torch.float32, device='cuda').cpu().numpy()
    A = rb.Matrix(a)
    x, y = rb.Vectors(N, k, np.float32), rb.Vectors(M, k, np.float32)
    x.fill_random_device(4)
    y2, x2 = rb.Vectors(M, k, np.float32), rb.Vectors(N, k, np.float32)
    fl = 2.0 * M * N * k
    for transp in (False, True):
        src, dst, dst2 = (y, x, x2) if transp else (x, y, y2)
        lib.rl_debug_set_knob(KNOB_GEMM_INSPLIT, 0)
        A.apply(src, dst, transp=transp)
        ms0, _ = timeit(lambda: A.apply(src, dst, transp=transp), reps=reps, warm=1)
        lib.rl_debug_set_knob(KNOB_GEMM_INSPLIT, 1)
        A.apply(src, dst2, transp=transp)
        ms1, _ = timeit(lambda: A.apply(src, dst2, transp=transp), reps=reps, warm=1)
        lib.rl_debug_set_knob(KNOB_GEMM_INSPLIT, 0)
        d0, d1 = dst.data(), dst2.data()
        err = float(np.abs(d0 - d1).max() / np.abs(d0).max())
        emit(exp='gemm_insplit', shape='A=%dx%d,k=%d' % (M, N, k), transp=transp, ms_default=round(ms0, 4),
             ms_insplit=round(ms1, 4), TFLOPs_default=round(fl / ms0 / 1e9, 1), TFLOPs_insplit=round(fl / ms1 / 1e9, 1),
             max_rel_diff=err, ok=bool(err < 1e-5))


def gram_edge_checks():
    """Ragged n, windows with an offset, m != k, against float64 NumPy on the host."""
    rng = np.random.RandomState(5)
    for n, m, k in ((8192, 32, 32), (10007, 17, 32), (65539, 9, 5), (20000, 40, 33), (12345, 8, 8), (9000, 1, 32)):
        x, y = rng.randn(m + 3, n), rng.randn(k + 2, n)
        Xv, Yv = rb.Vectors(x.copy()), rb.Vectors(y.copy())
        Xv.select(m, 3)
        Yv.select(k, 2)
        want = y[2:] @ x[3:].T
        for (mode, il) in GRAM_MODES:
            lib.rl_debug_set_knob(KNOB_GRAM_TMA, mode)
            lib.rl_debug_set_knob(KNOB_GRAM_WAVES, il)
            got = Xv.dot(Yv)
            err = float(np.abs(got - want).max() / np.abs(want).max())
            emit(exp='gram_edge', n=n, m=m, k=k, mode=mode, waves=il, rel_err_vs_numpy=err, ok=bool(err < 1e-12))
            if m == k:
                got = Xv.dot(Xv)
                w2 = x[3:] @ x[3:].T
                err = float(np.abs(got - w2).max() / np.abs(w2).max())
                emit(exp='gram_edge_same', n=n, m=m, mode=mode, waves=il, rel_err_vs_numpy=err, ok=bool(err < 1e-12))
    lib.rl_debug_set_knob(KNOB_GRAM_TMA, 0)
    lib.rl_debug_set_knob(KNOB_GRAM_WAVES, 0)


def _priv(op, name):
    return getattr(op, '_SparseSymmetricMatrix__' + name)


def spmm_sweep(name, A_full, plans, reps):
    """plans: list of (m, groups, carveouts)."""
    t0 = time.time()
    saved = rsp.SPMM_CLUSTER_WARPS
    rsp.SPMM_CLUSTER_WARPS = 0
    op = rb.SparseSymmetricMatrix(A_full)
    rsp.SPMM_CLUSTER_WARPS = saved
    n = A_full.shape[0]
    nnz = op.nnz()
    full = A_full.tocsr()
    indptr = np.ascontiguousarray(full.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(full.indices, dtype=np.int32)
    setup_s = time.time() - t0
    ip, ix, va = _priv(op, 'indptr'), _priv(op, 'indices'), _priv(op, 'values')
    orders = {}
    for (m, groups, carveouts) in plans:
      X, Y, Y0 = rb.Vectors(n, m), rb.Vectors(n, m), rb.Vectors(n, m)
      X.fill_random_device(3)
      byts = nnz * 12.0 + (n + 1) * 8.0 + 2.0 * n * m * 8
      y0 = None
      for group in groups:
        order_buf, ratio, cl_s = None, None, 0.0
        if group:
            if group not in orders:
                t1 = time.time()
                order, ratio = rsp.cluster_runs(indptr, indices, n, group)
                orders[group] = (rsp._to_device(order), ratio, time.time() - t1)
            order_buf, ratio, cl_s = orders[group]
        warps = group if group else 4
        for (carve, pf) in carveouts:
            lib.rl_debug_set_knob(KNOB_SPMM_WPS, carve)
            lib.rl_debug_set_knob(KNOB_SPMM_PF, pf)
            f = lambda: check(lib.rl_csr_spmm_ex(1, n, nnz, ip.ptr, ix.ptr, va.ptr, X._wptr(), X._ld, Y._wptr(), Y._ld, m,
                                                 0, None, order_buf.ptr if order_buf else None, warps, dev.stream()))
            Y.zero()
            f()
            torch.cuda.synchronize()
            if y0 is None:
                check(lib.rl_copy(1, Y0._wptr(), Y0._ld, Y._wptr(), Y._ld, m, n, dev.stream()))
                y0 = True
                diff = 0.0
            else:
                check(lib.rl_axpy(1, Y._wptr(), Y._ld, Y0._wptr(), Y0._ld, m, n, -1.0, dev.stream()))
                d = Y.dots(Y)
                diff = float(np.abs(d).max())
            ms, best = timeit(f, reps=reps)
            emit(exp='spmm', matrix=name, n=n, nnz=nnz, m=m, group=group, warps=warps, wps=carve, pf=pf,
                 footprint_ratio=ratio, cluster_setup_s=round(cl_s, 3), ms=round(ms, 5), ms_best=round(best, 5),
                 GBps=round(byts / ms / 1e6, 1), frac_hbm=round(byts / ms / 1e6 / peak_gbs(), 3),
                 sq_diff_vs_default=diff)
      lib.rl_debug_set_knob(KNOB_SPMM_WPS, 0)
      lib.rl_debug_set_knob(KNOB_SPMM_PF, 0)
      if op.layout() == 'sell32':
        ms, best = timeit(lambda: op.apply(X, Y), reps=reps)
        emit(exp='spmm', matrix=name, n=n, nnz=nnz, m=m, group='sell32', ms=round(ms, 5), GBps=round(byts / ms / 1e6, 1),
             frac_hbm=round(byts / ms / 1e6 / peak_gbs(), 3))
    emit(exp='spmm_setup', matrix=name, operator_setup_s=round(setup_s, 2))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='')
    ap.add_argument('--N', type=int, default=128)
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--only', default='edge,gram,spmm')
    ap.add_argument('--window', action='store_true', help='also time the experimental band-window SpMM kernel')
    args = ap.parse_args()
    only = set(args.only.split(','))
    if args.out:
        OUT[0] = open(args.out, 'w')
    if 'ncu' in only:
        ncu_pass(args.N)
    if 'edge' in only:
        gram_edge_checks()
    if 'gram' in only:
        gram_sweep(2097152, [(32, 32), (16, 16), (8, 8), (32, 16), (24, 24)], args.reps)
        gram_sweep(140874, [(32, 32)], args.reps)
        gram_sweep(32768, [(16, 16)], args.reps)
    if 'gemm' in only:
        gemm_sweep(12000, 39375, 128, 5)
        gemm_sweep(3000, 2000, 128, 5)
        gemm_sweep(1000, 777, 40, 3)
    if 'spmm' in only:
        N = args.N
        spmm_sweep('lap3d_%d' % N, lap3d_slab(N, 0, N ** 3),
                   [(32, (0,), ((24, -1), (24, 1), (24, 3), (24, 7), (24, 5))), (16, (0,), ((24, -1), (24, 1), (24, 3), (24, 7))), (8, (0,), ((24, -1), (24, 1), (24, 3)))], args.reps)
        spmm_sweep('lap3d_32', lap3d_slab(32, 0, 32 ** 3), [(16, (0,), ((24, -1), (24, 1), (24, 3)))], args.reps)
        from tests_common import spd_c3_like
        offs = tuple(sorted(set([1, 2, 3, 4, 5, 6, 440, 441, 442, 443, 444, 445, 446, 2656, 2657, 2658, 2659, 2660,
                                 2661, 2662, 2214, 2215, 2216, 2217, 3100, 3101, 3102])))
        spmm_sweep('c3like_55nnz', spd_c3_like(140874, offsets=offs), [(32, (0,), ((16, -1), (16, 1), (16, 3)))], args.reps)
