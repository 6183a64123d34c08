"""BASELINE config 5 (reduced): incremental chunked PCA, pca(A, tol=..., batch_size=...) on the
GPU backend (lra.icompute / lra.update run verbatim: chunk-as-vectors dots, mean update,
orthogonalize against the components, deflated solve, svd-based re-orthogonalisation).

    python tools/run_c5.py [--rows 262144] [--cols 4096] [--chunk 65536] [--tol 0.05] [--cpu]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from threadpoolctl import threadpool_limits
import raleigh_b200 as rb
from raleigh_b200 import profile
rb.install()
from raleigh.interfaces.pca import pca
from raleigh.core.solver import Options

ap = argparse.ArgumentParser()
ap.add_argument('--rows', type=int, default=262144)
ap.add_argument('--cols', type=int, default=4096)
ap.add_argument('--chunk', type=int, default=65536)
ap.add_argument('--rank', type=int, default=256)
ap.add_argument('--tol', type=float, default=0.05)
ap.add_argument('--cpu', action='store_true')
ap.add_argument('--no-warmup', action='store_true', help='time the very first pass (includes one-time set-up costs)')
ap.add_argument('--host-profile', action='store_true', help='cProfile of a second run (rank 0): where the host time goes')
args = ap.parse_args()

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
if world > 1:
    # sample-partitioned run (torchrun, one process per GPU): every process generates and keeps rows/world rows of
    # the matrix -- the same right factor everywhere, its own left factor -- and feeds its share of every chunk
    import torch.distributed as tdist
    from raleigh_b200 import dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    tdist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', rank))))
    dist.enable()
    assert args.rows % world == 0 and args.chunk % world == 0
g = torch.Generator(device='cuda'); g.manual_seed(1)
r = args.rank
sigma = torch.arange(1, r + 1, device='cuda', dtype=torch.float32) ** (-0.75)
v, _ = torch.linalg.qr(torch.randn(args.cols, r, generator=g, device='cuda'))
g.manual_seed(100 + rank)
rows_here = args.rows // world
u = torch.randn(rows_here, r, generator=g, device='cuda') / args.rows ** 0.5
a = (u * sigma[None, :]) @ v.T
a += 1e-3 * sigma[-1] * torch.randn(rows_here, args.cols, generator=g, device='cuda') / args.cols ** 0.5
A = a.cpu().numpy()
args.chunk //= world
with threadpool_limits(limits=1):
    if not args.no_warmup:
        # one untimed pass: NCCL communicators, pinned staging buffers, allocator growth, lazy kernel loading
        np.random.seed(1)
        pca(A, tol=args.tol, batch_size=args.chunk, arch='gpu!', opt=Options())
        torch.cuda.synchronize()
    np.random.seed(1)
    profile.reset(); profile.enable(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mean, trans, comps = pca(A, tol=args.tol, batch_size=args.chunk, arch='gpu!', opt=Options())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    profile.enable(False)
prof = profile.report()
t = torch.as_tensor(trans[rank * rows_here:(rank + 1) * rows_here], device='cuda'); c = torch.as_tensor(comps, device='cuda')
ds = a - torch.as_tensor(mean, device='cuda').reshape(1, -1)
num, den = torch.linalg.norm(t @ c - ds) ** 2, torch.linalg.norm(ds) ** 2
if world > 1:
    both = torch.stack((num, den)); tdist.all_reduce(both); num, den = both[0], both[1]
    tmax = torch.tensor([dt], device='cuda'); tdist.all_reduce(tmax, op=tdist.ReduceOp.MAX); dt = float(tmax.item())
ef = float(torch.sqrt(num / den).item())
line = {'config': 'C5 (%s): incremental PCA of %dx%d fp32 in %d-row chunks, tol %.2g' % ('single GPU' if world == 1 else '%d GPUs, samples partitioned' % world, args.rows, args.cols, args.chunk * world, args.tol),
        'gpu_s': round(dt, 3), 'warmup_passes': 0 if args.no_warmup else 1, 'components': int(comps.shape[0]), 'pca_error_frobenius': ef,
        'device_ms': round(sum(v_['ms'] for v_ in prof.values()), 1),
        'kernels': {k: {'count': v_['count'], 'ms': round(v_['ms'], 1)} for k, v_ in prof.items()}}
if args.cpu:
    np.random.seed(1)
    t0 = time.perf_counter()
    mean2, trans2, comps2 = pca(A, tol=args.tol, batch_size=args.chunk, arch='cpu', opt=Options())
    line['cpu_s'] = round(time.perf_counter() - t0, 2)
    line['cpu_components'] = int(comps2.shape[0])
if rank == 0:
    print(json.dumps(line), flush=True)
if args.host_profile:
    import cProfile, pstats
    with threadpool_limits(limits=1):
        np.random.seed(1)
        pr = cProfile.Profile(); pr.enable()
        pca(A, tol=args.tol, batch_size=args.chunk, arch='gpu!', opt=Options())
        torch.cuda.synchronize(); pr.disable()
    if rank == 0:
        pstats.Stats(pr).sort_stats('tottime').print_stats(30)
    if world > 1:
        tdist.barrier()
if world > 1:
    tdist.barrier()
    tdist.destroy_process_group()
