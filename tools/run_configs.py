"""Time-to-k-eigenpairs on the sparse BASELINE configs, GPU backend next to the CPU path,
same inputs / seeds / solver options:

  C1: 3D 7-point Laplacian 32^3 (32,768 rows), 10 smallest eigenpairs, block 16 (auto)
  C3: synthetic shipsec1-shaped SPD CSR (140,874 rows, ~55 nnz/row, fp64), 10 smallest,
      Jacobi preconditioning, block 32 -- through partial_hevp(A, T=...)

Three arms on the same inputs: the device-resident driver (raleigh_b200/jcg.py, default), the
reference's UNMODIFIED main loop on the same GPU backend (verbatim path), and the reference solver on
the oracle port of its NumPy algebra on the host (MKL is absent).  Prints one JSON line per config:
times, iteration counts side by side, max relative eigenvalue differences, residuals.

    python tools/run_configs.py [c1] [c3] [--tol 1e-6] [--no-cpu]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402
from threadpoolctl import threadpool_limits  # noqa: E402
import raleigh_b200 as rb  # noqa: E402
import oracle  # noqa: E402
from oracle import algebra_np as K  # noqa: E402
from tests_common import spd_c3_like  # noqa: E402

rb.install()
import raleigh.core.solver as rs  # noqa: E402

C3_OFFSETS = tuple(sorted({1, 2, 3, 4, 5, 6, 440, 441, 442, 443, 444, 445, 446, 2656, 2657, 2658, 2659, 2660, 2661,
                           2662, 2214, 2215, 2216, 2217, 3100, 3101, 3102}))


def solve(Vectors, op, n, dtype, nev, tol, block, T=None, max_iter=2000):
    np.random.seed(1)
    opt = rs.Options()
    opt.block_size = block
    opt.max_iter = max_iter
    opt.convergence_criteria = rs.DefaultConvergenceCriteria()
    opt.convergence_criteria.set_error_tolerance('k eigenvector error', tol)
    v = Vectors(n, data_type=dtype)
    solver = rs.Solver(rs.Problem(v, op))
    if T is not None:
        solver.set_preconditioner(T)
    t0 = time.perf_counter()
    status = solver.solve(v, opt, which=(nev, 0))
    if Vectors is rb.Vectors:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dict(status=int(status), iterations=int(solver.iteration), seconds=dt,
                lmd=np.sort(np.array(solver.eigenvalues)), x=v)


def both(fn):
    """(device-resident driver, verbatim main loop) on the GPU backend."""
    rb.use_device_solver(False)
    try:
        verb = fn()
    finally:
        rb.use_device_solver(True)
    return fn(), verb


def report(name, A, gpu, cpu, extra, verb=None):
    lmd = gpu['lmd']
    line = {'config': name, 'gpu_s': round(gpu['seconds'], 4), 'gpu_iterations': gpu['iterations'],
            'gpu_status': gpu['status'], 'eigenvalues': [float(t) for t in lmd[:10]]}
    if verb is not None:
        k = min(len(lmd), len(verb['lmd']))
        line.update({'gpu_verbatim_s': round(verb['seconds'], 4), 'gpu_verbatim_iterations': verb['iterations'],
                     'max_rel_eigenvalue_diff_device_vs_verbatim':
                     float(np.max(np.abs(lmd[:k] - verb['lmd'][:k]) / np.abs(verb['lmd'][:k])))})
    line.update(extra)
    if cpu is not None:
        k = min(len(lmd), len(cpu['lmd']))
        line.update({'cpu_s': round(cpu['seconds'], 3), 'cpu_iterations': cpu['iterations'],
                     'cpu_status': cpu['status'], 'cpu_threads': 1,
                     'max_rel_eigenvalue_diff': float(np.max(np.abs(lmd[:k] - cpu['lmd'][:k]) / np.abs(cpu['lmd'][:k]))),
                     'speedup': round(cpu['seconds'] / gpu['seconds'], 2)})
    print(json.dumps(line), flush=True)
    return line


def residuals(A, sol):
    x = sol['x'].data()
    lam = np.array([float((xi @ (A @ xi)) / (xi @ xi)) for xi in x])
    r = A @ x.T - x.T * lam[None, :]
    return float(np.max(np.linalg.norm(r, axis=0) / np.maximum(np.abs(lam), 1e-300)))


def main():
    args = sys.argv[1:]
    tol = 1e-6
    if '--tol' in args:
        tol = float(args[args.index('--tol') + 1])
    do_cpu = '--no-cpu' not in args
    which = [a for a in args if a in ('c1', 'c3')] or ['c1', 'c3']
    out = []
    with threadpool_limits(limits=1):       # the survey found 1 BLAS thread fastest for the CPU path too
        if 'c1' in which:
            L = K.lap3d_csr(32, 32, 32)
            n = L.shape[0]
            op = rb.SparseSymmetricMatrix(L)
            solve(rb.Vectors, op, n, np.float64, 10, tol, -1)            # warm-up (allocator, module load)
            gpu, verb = both(lambda: solve(rb.Vectors, op, n, np.float64, 10, tol, -1))
            cpu = solve(oracle.Vectors, oracle.SparseSymmetricMatrix(L), n, np.float64, 10, tol, -1) if do_cpu else None
            exact = K.lap3d_eigenvalues(32, 32, 32)[:10]
            out.append(report('C1 lap3d 32^3, 10 smallest, tol %g, block auto' % tol, L, gpu, cpu, {
                'n': n, 'nnz': int(L.nnz), 'max_rel_err_vs_analytic': float(np.max(np.abs(gpu['lmd'] - exact) / exact)),
                'max_rel_residual': residuals(L, gpu), 'spmm_layout': op.layout()}, verb))
        if 'c3' in which:
            from raleigh.interfaces.partial_hevp import partial_hevp
            n = 140874
            A = spd_c3_like(n, offsets=C3_OFFSETS)
            op = rb.SparseSymmetricMatrix(A)
            T = rb.Operator(rb.DiagonalPreconditioner(A))
            solve(rb.Vectors, op, n, np.float64, 10, 1e-2, 32, T=T)     # warm-up
            gpu, verb = both(lambda: solve(rb.Vectors, op, n, np.float64, 10, tol, 32, T=T))
            cpu = None
            if do_cpu:
                cpu = solve(oracle.Vectors, oracle.SparseSymmetricMatrix(A), n, np.float64, 10, tol, 32,
                            T=oracle.Operator(oracle.Jacobi(A)))
            line = report('C3 synthetic SPD n=140874 (~%d nnz/row), 10 smallest, Jacobi, block 32, tol %g'
                          % (A.nnz // n, tol), A, gpu, cpu, {
                              'n': n, 'nnz': int(A.nnz), 'max_rel_residual': residuals(A, gpu), 'spmm_layout': op.layout()},
                          verb)
            # the same through the reference's partial_hevp entry point (preconditioned branch)
            np.random.seed(1)
            opt = rs.Options()
            opt.block_size = 32
            opt.max_iter = 2000
            t0 = time.perf_counter()
            lmd, x, status = partial_hevp(A, T=rb.DiagonalPreconditioner(A), which=10, tol=tol, verb=-1, opt=opt)
            dt = time.perf_counter() - t0
            print(json.dumps({'config': 'C3 via partial_hevp(A, T=DiagonalPreconditioner)', 'gpu_s_incl_setup': round(dt, 3),
                              'status': int(status), 'max_rel_diff_vs_core_solver_run':
                              float(np.max(np.abs(np.sort(lmd) - gpu['lmd']) / gpu['lmd']))}), flush=True)
            out.append(line)
    return out


if __name__ == '__main__':
    main()
