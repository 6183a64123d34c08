"""BASELINE config 4 (bounded): 3D 7-point Laplacian N^3, rows sharded over the ranks,
smallest eigenpairs with the reference's unmodified core solver on the raleigh_b200 backend.

    python tools/run_c4.py --N 256 --block 32 --nev 20 --iters 12            # 1 GPU
    torchrun --nproc-per-node 8 ... tools/run_c4.py --N 256 --block 120 --nev 100 --iters 12

No rank ever forms the global matrix (each builds the CSR of its own z-slab).  A full
solve at 256^3 needs >~1000 iterations; this harness runs a FIXED
number of iterations and reports seconds per iteration, the per-kernel device time and
achieved GB/s (CUDA events inside the library), NCCL traffic -- flagged as a bounded run.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lap3d_slab(N, row0, nloc):
    """Rows [row0, row0+nloc) of the 7-point Laplacian on an N^3 grid (h = 1/(N+1), x fastest),
    same operator as examples/laplace.py:23-27, built without the global matrix."""
    h2 = float(N + 1) ** 2
    r = np.arange(row0, row0 + nloc, dtype=np.int64)
    x, y, z = r % N, (r // N) % N, r // (N * N)
    cols = [r]
    vals = [np.full(nloc, 6.0 * h2)]
    rows = [r - row0]
    for ok, off in ((x > 0, -1), (x < N - 1, 1), (y > 0, -N), (y < N - 1, N), (z > 0, -N * N), (z < N - 1, N * N)):
        rows.append((r - row0)[ok])
        cols.append(r[ok] + off)
        vals.append(np.full(int(ok.sum()), -h2))
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nloc, N ** 3))
    A.sort_indices()
    return A


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--N', type=int, default=256)
    ap.add_argument('--block', type=int, default=32)
    ap.add_argument('--nev', type=int, default=20)
    ap.add_argument('--iters', type=int, default=12)
    ap.add_argument('--dtype', default='float64')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    import torch
    torch.cuda.set_device(local)
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=1)
    import raleigh_b200 as rb
    from raleigh_b200 import dist, profile
    ctx = None
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group('nccl', device_id=torch.device('cuda', local))
        ctx = dist.enable()
    rb.install()
    import raleigh.core.solver as rs
    dtype = np.dtype(args.dtype).type
    n = args.N ** 3
    t0 = time.perf_counter()
    if world > 1:
        row0, nloc = dist.partition(n, world, rank)
        op = rb.SparseSymmetricMatrix(lap3d_slab(args.N, row0, nloc).astype(dtype), local_rows=(row0, n))
    else:
        row0, nloc = 0, n
        op = rb.SparseSymmetricMatrix(lap3d_slab(args.N, 0, n).astype(dtype))
    setup = time.perf_counter() - t0

    def run(iters):
        np.random.seed(1)
        opt = rs.Options()
        opt.block_size = args.block
        opt.max_iter = iters
        opt.verbosity = -1
        opt.convergence_criteria = rs.DefaultConvergenceCriteria()
        opt.convergence_criteria.set_error_tolerance('k eigenvector error', 1e-6)
        v = rb.Vectors(n, data_type=dtype)
        solver = rs.Solver(rs.Problem(v, op))
        torch.cuda.synchronize()
        t = time.perf_counter()
        status = solver.solve(v, opt, which=(args.nev, 0))
        torch.cuda.synchronize()
        return time.perf_counter() - t, solver, status

    run(3)                                     # warm-up: allocator, NCCL channels
    profile.reset()
    profile.enable(True)
    dt, solver, status = run(args.iters)
    profile.enable(False)
    prof = profile.report()
    if world > 1:
        t = torch.tensor([dt], device='cuda')
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        dt = t.item()
    if rank == 0:
        its = max(int(solver.iteration), 1)
        line = {'config': 'C4 (bounded): lap3d %d^3 = %d rows, %s, block %d, %d smallest wanted, %d iterations run '
                          '(status %d = iteration limit)' % (args.N, n, args.dtype, args.block, args.nev, its, status),
                'n_gpus': world, 'rows_per_gpu': nloc, 'seconds_total': round(dt, 4),
                'seconds_per_iteration': round(dt / its, 5), 'setup_s': round(setup, 2),
                'device_ms_per_iteration': round(sum(v['ms'] for v in prof.values()) / its, 3),
                'kernels': {k: {'count': v['count'], 'ms': round(v['ms'], 2), 'GBps': round(v['GBps'], 1),
                                'TFLOPs': round(v['TFLOPs'], 2)} for k, v in prof.items()},
                'spmm_layout': op.layout()}
        if ctx is not None:
            line['nccl'] = {'allreduce_calls': ctx.allreduce_calls, 'allreduce_MB': round(ctx.allreduce_bytes / 1e6, 2),
                            'halo_MB': round(getattr(op, 'halo_bytes', 0) / 1e6, 1)}
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == '__main__':
    main()
