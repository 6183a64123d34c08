"""Device time of the small-matrix kernels of the Rayleigh-Ritz step (CUDA events, warm cache):
pivoted Cholesky, cluster Jacobi, the whole rl_rr_solve, at the block sizes of the BASELINE configs.

    python tools/time_rr.py [--out gpurun_out/time_rr.jsonl]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raleigh_b200  # noqa: E402,F401
from raleigh_b200._lib import lib, check  # noqa: E402
from raleigh_b200 import device as dev  # noqa: E402


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def up(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--blocks', default='16,32,64,120,128,160')
    ap.add_argument('--no-big', action='store_true')
    args = ap.parse_args()
    out = open(args.out, 'w') if args.out else None

    def emit(**rec):
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + '\n')

    st = dev.stream
    for m in [int(t) for t in args.blocks.split(',') if t]:
        n = 2 * m
        rng = np.random.RandomState(m)
        N = 6 * n
        V = rng.randn(N, n)
        V[:, :m], _ = np.linalg.qr(V[:, :m])
        V[:, m:] -= V[:, :m] @ (V[:, :m].T @ V[:, m:])
        V[:, m:] /= np.linalg.norm(V[:, m:], axis=0)
        A = np.diag(np.linspace(1.0, 1000.0, N))
        GB, GA = V.T @ V, V.T @ A @ V
        dGB0, dGA = up(GB), up(GA)
        dGB, dA0 = torch.zeros_like(dGB0), torch.zeros_like(dGB0)
        ind = torch.zeros(n, dtype=torch.int32, device='cuda')
        info = torch.zeros(8, dtype=torch.int32, device='cuda')

        def chol():
            dGB.copy_(dGB0)
            check(lib.rl_rr_piv_chol(dGB.data_ptr(), dA0.data_ptr(), n, n, m, 1e-8, ind.data_ptr(), info.data_ptr(), st()))
        t_chol = timeit(chol)
        assert int(info[0]) == 0
        lib.rl_debug_set_knob(13, 1)           # the same without the condition estimates of the drop rule
        t_chol_noest = timeit(chol)
        lib.rl_debug_set_knob(13, 0)
        wsb = lib.rl_rr_solve_ws_bytes(n)
        ws = torch.zeros(wsb // 8 + 8, dtype=torch.float64, device='cuda')
        cx, cz = torch.zeros(n, n, dtype=torch.float64, device='cuda'), torch.zeros(n, n, dtype=torch.float64, device='cuda')
        lx, lz, est = (torch.zeros(2 * n, dtype=torch.float64, device='cuda') for _ in range(3))

        def rr():
            check(lib.rl_rr_solve(dGA.data_ptr(), dGB.data_ptr(), n, m, m, m, 0, m, 0, cx.data_ptr(), n, cz.data_ptr(), n,
                                  lx.data_ptr(), lz.data_ptr(), est.data_ptr(), n, 0.0, ws.data_ptr(), wsb, info.data_ptr(), st()))
        t_rr = timeit(rr)
        sweeps_rr = int(info[0])
        w, Q = torch.zeros(n, dtype=torch.float64, device='cuda'), torch.zeros(n, n, dtype=torch.float64, device='cuda')
        ewsb = lib.rl_small_eigh_ws_bytes(n)
        ews = torch.zeros(ewsb // 8 + 8, dtype=torch.float64, device='cuda')
        G = rng.randn(n, n)
        G = up(G + G.T)
        res = {}
        for p in (m, n):
            def eig():
                check(lib.rl_small_eigh(G.data_ptr(), n, p, 0.0, w.data_ptr(), Q.data_ptr(), n, ews.data_ptr(), ewsb, info.data_ptr(), st()))
            res['eigh_%d_ms' % p] = round(timeit(eig, 10), 4)
            res['eigh_%d_sweeps' % p] = int(info[0])
            if p <= 320:         # cycle counts of CTA 0 by phase (csrc/jacobi.cu writes them behind the parameters)
                cyc = ews[4:8].view(torch.int64).cpu().numpy()
                res['eigh_%d_kcycles_intra_norms_cross_exchange' % p] = [int(c // 1000) for c in cyc]
        B = up(rng.randn(n, n))

        def trsm():
            check(lib.rl_small_trsm(0, dGB.data_ptr(), n, n, B.data_ptr(), n, n, st()))
        t_trsm = timeit(trsm)
        emit(block=m, nxy=n, piv_chol_ms=round(t_chol, 4), piv_chol_no_estimates_ms=round(t_chol_noest, 4), rr_solve_ms=round(t_rr, 4), rr_final_eigh_sweeps=sweeps_rr,
             trsm_ms=round(t_trsm, 4), **res)
    if not args.no_big:
        big(emit)


def big(emit):
    """Orders of the partial-SVD post-processing (nsv = 1000 at config 2): blocked Cholesky and Jacobi on the
    factor for a graded, nearly diagonal Gram matrix; symmetric mode for comparison."""
    st = dev.stream
    for n in (500, 1000):
        rng = np.random.RandomState(n)
        sig = np.arange(1, n + 1) ** -0.75
        E = rng.randn(n, n) * 3e-4
        S = np.eye(n) + E + E.T
        G = up(sig[:, None] * S * sig[None, :])
        U = torch.zeros_like(G)
        info = torch.zeros(8, dtype=torch.int32, device='cuda')
        w, Q = torch.zeros(n, dtype=torch.float64, device='cuda'), torch.zeros(n, n, dtype=torch.float64, device='cuda')
        ewsb = lib.rl_small_eigh_ws_bytes(n)
        ews = torch.zeros(ewsb // 8 + 8, dtype=torch.float64, device='cuda')

        def potrf():
            U.copy_(G)
            check(lib.rl_small_potrf(U.data_ptr(), n, n, info.data_ptr(), st()))
        t_potrf = timeit(potrf, 5)

        def fac():
            check(lib.rl_small_eigh_factor(U.data_ptr(), n, n, 0.0, w.data_ptr(), Q.data_ptr(), n, ews.data_ptr(), ewsb, info.data_ptr(), st()))
        t_fac = timeit(fac, 3)
        sw_fac = int(info[0])

        def sym():
            check(lib.rl_small_eigh(G.data_ptr(), n, n, 0.0, w.data_ptr(), Q.data_ptr(), n, ews.data_ptr(), ewsb, info.data_ptr(), st()))
        t_sym = timeit(sym, 3)
        sw_sym = int(info[0])
        lib.rl_debug_set_knob(14, 1)          # flat grid kernel (one grid barrier per round), for comparison
        t_fac_flat = timeit(fac, 3)
        sw_flat = int(info[0])
        lib.rl_debug_set_knob(14, 0)
        dbg = {}
        for mode, name in ((1, 'ring6_exchange_only_ms'), (2, 'ring6_compute_only_ms'), (3, 'ring6_load_store_only_ms')):
            lib.rl_debug_set_knob(15, mode)       # six sweeps with parts switched off (results meaningless)
            dbg[name] = round(timeit(fac, 3), 3)
        lib.rl_debug_set_knob(15, 0)
        emit(order=n, **dbg)
        emit(order=n, potrf_ms=round(t_potrf, 3), eigh_factor_ms=round(t_fac, 3), eigh_factor_sweeps=sw_fac,
             eigh_sym_ms=round(t_sym, 3), eigh_sym_sweeps=sw_sym, eigh_factor_flat_ms=round(t_fac_flat, 3),
             eigh_factor_flat_sweeps=sw_flat)


if __name__ == '__main__':
    main()
