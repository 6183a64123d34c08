#!/usr/bin/env python
"""bench.py -- time to k eigenpairs on BASELINE.json's quoted configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload = "C2"): PCA / truncated SVD of a synthetic
12,000 x 39,375 fp32 low-rank + noise matrix (LFW 175x225 shape), 1,000 principal
components -- BASELINE.json configs[1], the configuration the reference's README
quotes (27 s CPU / 12 s GPU).  One "step" = one complete solve: the reference's
unmodified LowerRankApproximation.compute -> PartialSVD -> core block-JCG solver
running on the raleigh_b200 Vectors/Matrix backend.

  value : seconds per solve with the data matrix already resident in HBM
  e2e   : seconds per `pca(A_host, npc=1000, arch='gpu!')` call: host matrix in
          pinned memory -> H2D inside the timed region, components/scores D2H
  roofline : the dominant kernel (dense operator application), measured live
          with CUDA events inside the C library over the timed region
  cpu_baseline : the reference's own CPU path (dense_numpy on NumPy/OpenBLAS; MKL
          is not installable here) on the box's host cores

--impl reference times only that CPU path and prints the same line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_ROWS, N_COLS, RANK_GEN, NPC = 12000, 39375, 2000, 1000
METRIC = 'time to k eigenpairs (s)'
BASELINE_PUBLISHED_S = 12.0      # README.md:33 "raleigh GPU" column, 1000 components (GPU model not stated)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--rows', type=int, default=M_ROWS, help='debug: shrink the workload')
    ap.add_argument('--cols', type=int, default=N_COLS)
    ap.add_argument('--npc', type=int, default=NPC)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


# --------------------------------------------------------------------------- workload
def generate_c2(rows, cols, rank, device, seed=1, world=1, shard=0):
    """Synthetic data matrix in the style of the reference's generator
    (examples/pca/generate_matrix.py:55-77 with pca=True, alpha=0.75, plus the
    --ptb noise, :122-125): A = U diag(k^-0.75) V^T + noise, U[:, 0] = const.
    With world > 1 this returns rows [shard*rows, (shard+1)*rows) of the
    (world*rows) x cols matrix: V is common to all shards, U is drawn per shard
    and scaled so that its columns stay (nearly) orthonormal globally."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rank = min(rank, rows, cols)
    sigma = torch.arange(1, rank + 1, device=device, dtype=torch.float32) ** (-0.75)
    v = torch.randn(cols, rank, generator=g, device=device, dtype=torch.float32)
    v, _ = torch.linalg.qr(v)
    g.manual_seed(seed + 7919 * (shard + 1))
    u = torch.randn(rows, rank, generator=g, device=device, dtype=torch.float32)
    u[:, 0] = 1.0
    u, _ = torch.linalg.qr(u)
    if world > 1:
        u = u / float(world) ** 0.5
    a = (u * sigma[None, :]) @ v.T
    del u, v
    noise = 2 * torch.rand(rows, cols, generator=g, device=device, dtype=torch.float32) - 1
    scale = sigma[-1] / (10 * torch.linalg.norm(noise, dim=1))
    a += scale[:, None] * noise
    del noise
    return a.contiguous()


def pca_error_gpu(a_dev, mean, trans, comps):
    """pca_error (interfaces/pca.py:165-174) evaluated with torch on the device
    (verification only, outside every timed region)."""
    import torch
    t = torch.as_tensor(trans, device=a_dev.device)
    c = torch.as_tensor(comps, device=a_dev.device)
    mu = torch.as_tensor(mean, device=a_dev.device).reshape(1, -1)
    data_s = a_dev - mu
    err = t @ c - data_s
    em = (torch.linalg.norm(err, dim=1).max() / torch.linalg.norm(data_s, dim=1).max()).item()
    ef = (torch.linalg.norm(err) / torch.linalg.norm(data_s)).item()
    return em, ef


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                 '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1:
                continue
            parts = [p.strip() for p in line.split(',')]
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except (ValueError, IndexError):
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                 parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smax, 'samples': len(sm),
                'reasons': sorted(reasons)}


# --------------------------------------------------------------------------- reference CPU path
def load_reference_cpu():
    """The reference package with its own CPU algebra (dense_cpu -> dense_numpy,
    MKL being absent); only the SciPy `turbo` shim is applied."""
    from raleigh_b200.compat import find_reference, shim_scipy, unshim_host_hotspots
    path = find_reference()
    if path is None:
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    shim_scipy()
    unshim_host_hotspots()      # the GPU arm's vectorised `_norm` must not speed up the reference arm
    return path


def cpu_pca_seconds(a_host, npc, budget_s=60.0):
    """Time the reference's CPU pca on the box's host cores.  Runs the FULL
    workload when a calibration GEMM says it fits the budget, otherwise a
    row/component-proportional sample, extrapolated (and labelled so)."""
    import numpy as np
    from raleigh.interfaces.pca import pca
    from raleigh.core.solver import Options
    rows, cols = a_host.shape
    x = np.random.rand(128, cols).astype(np.float32)
    t0 = time.perf_counter()
    y = x @ a_host.T
    _ = y @ a_host
    t_pair = time.perf_counter() - t0
    est = 22 * t_pair * (npc / 1000.0) + 3.0
    frac = 1.0
    if est > budget_s:
        frac = max(0.1, (budget_s / est) ** 0.5)
    r, k = int(rows * frac), max(8, int(npc * frac))
    sample = a_host if frac == 1.0 else np.ascontiguousarray(a_host[:r])
    from raleigh.interfaces.lra import LowerRankApproximation
    from raleigh.algebra.dense_matrix import AMatrix
    t0 = time.perf_counter()
    # exactly what pca(sample, npc=k, arch='cpu') does (pca.py:142-153), spelled out so that
    # the solver's iteration count can be reported next to the GPU run's
    lra = LowerRankApproximation()
    lra.ortho = 1e-3 if sample.shape[0] < sample.shape[1] else 0
    lra.compute(AMatrix(sample, arch='cpu'), opt=Options(), rank=k, tol=0, norm='f', max_rank=-1, svtol=1e-3,
                shift=True, verb=0)
    trans, comps, mean = lra.left(), lra.right(), lra.mean()
    t = time.perf_counter() - t0
    cpu_pca_seconds.last_iterations = int(lra.iterations)
    if frac == 1.0:
        desc = 'full C2 workload: pca(%dx%d fp32, npc=%d), reference dense_numpy on NumPy/OpenBLAS (not MKL)' % (
            rows, cols, npc)
        return t, desc, comps.shape[0]
    scale = (rows * npc) / float(r * k)
    desc = ('rows 0..%d of the C2 matrix, npc=%d (%.1f s measured), scaled x%.2f by rows*components to the '
            'full workload; reference dense_numpy on NumPy/OpenBLAS (not MKL)' % (r, k, t, scale))
    return t * scale, desc, comps.shape[0]


# --------------------------------------------------------------------------- main arms
def run_reference(args, rank, world):
    if rank != 0:
        return
    import numpy as np
    ref = load_reference_cpu()
    if ref is None:
        print(json.dumps({'impl': 'reference', 'unavailable': 'reference package not on this box'}))
        return
    import torch
    dev = 'cuda' if torch.cuda.is_available() else 'cpu'
    a_host = generate_c2(args.rows, args.cols, RANK_GEN, dev).cpu().numpy()
    np.random.seed(1)
    budget = 240.0
    times, desc = [], ''
    warm = min(args.warmup, 1)
    t_first, desc, ncomp = cpu_pca_seconds(a_host, args.npc)
    if warm == 0:
        times.append(t_first)
    per = max(t_first, 1e-3)
    while len(times) < args.steps and (len(times) + 1) * per < budget:
        t, desc, ncomp = cpu_pca_seconds(a_host, args.npc)
        times.append(t)
    if not times:
        times.append(t_first)
    val = sum(times) / len(times)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 's', 'n_gpus': args.gpus,
        'steps': args.steps, 'steps_executed': len(times), 'warmup': args.warmup, 'ms_per_step': val * 1e3,
        'higher_is_better': False, 'scaling': 'weak', 'vs_baseline': val / 27.0, 'dtype': 'f32',
        'data': 'synthetic', 'config': dict(workload_config(args, 1), parallelism='host CPU, %d cores' % os.cpu_count(),
                                                 solver='reference core solver + lra/partial_svd on the reference\'s own '
                                                        'dense_numpy algebra (NumPy/OpenBLAS; MKL not installable)'),
        'cpu_baseline': {'value': val, 'unit': 's', 'cores': os.cpu_count(), 'kind': 'reference', 'sample': desc},
        'e2e': {'value': val, 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'components': int(ncomp), 'solver_iterations': getattr(cpu_pca_seconds, 'last_iterations', None),
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {'workload': 'C2: PCA of synthetic %dx%d fp32 low-rank+noise (LFW 175x225 shape), %d components, '
                        'block 128, svtol 1e-3%s' % (args.rows * world, args.cols, args.npc,
                                                     '' if world == 1 else ' (%d rows per GPU)' % args.rows),
            'l2_policy': 'inputs_exceed_l2 (data matrix %.2f GB streamed every operator application)'
                         % (args.rows * args.cols * 4 / 1e9),
            'parallelism': 'single GPU' if world == 1 else
            'sample-partitioned data matrix over %d GPUs: row-sharded block vectors, NCCL all-reduce of Gram '
            'matrices and of the k x n_features partial products' % world,
            'solver': 'reference core solver + lra/partial_svd, unmodified, on raleigh_b200 backend'}


def run_b200(args, rank, world, local_rank):
    """GPU arm.  The host keeps only the solver's k x k algebra (<= 256 x 256), for
    which multi-threaded BLAS/LAPACK is slower than one thread (SURVEY.md section 6:
    8 threads were 3x slower than 1 on config 1; torchrun pins OMP_NUM_THREADS=1
    anyway), so host BLAS is limited to one thread here; the CPU baseline below
    gets all cores back."""
    import numpy as np
    import torch
    from threadpoolctl import threadpool_limits
    host_limit = threadpool_limits(limits=1)
    torch.cuda.set_device(local_rank)
    import raleigh_b200 as rb
    from raleigh_b200 import profile, cuda
    if rb.find_reference() is None:
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'value': None, 'unit': 's', 'n_gpus': world,
                              'error': 'reference solver package (baseline/_ref) not on this box'}))
        return
    rb.install()
    from raleigh.interfaces.pca import pca
    from raleigh.interfaces.lra import LowerRankApproximation
    from raleigh.algebra.dense_matrix import AMatrix
    from raleigh.core.solver import Options
    import torch.distributed as dist
    ctx = None
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG asks for it; keep stdout
        # for the single JSON line by pointing fd 1 at stderr while the communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            warm = torch.zeros(1, device='cuda')
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        from raleigh_b200 import dist as rdist
        ctx = rdist.enable()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    a_dev = generate_c2(args.rows, args.cols, RANK_GEN, 'cuda', seed=1, world=world, shard=rank)
    a_pinned = torch.empty((args.rows, args.cols), dtype=torch.float32, pin_memory=True)
    a_pinned.copy_(a_dev)
    a_host = a_pinned.numpy()
    torch.cuda.synchronize()

    def solve_resident(matrix):
        np.random.seed(1)
        lra = LowerRankApproximation()
        lra.ortho = 1e-3                      # pca(): m < n  =>  ortho = svtol (pca.py:147-148)
        lra.compute(matrix, opt=Options(), rank=args.npc, tol=0, norm='f', max_rank=-1, svtol=1e-3,
                    shift=True, verb=0)
        return lra

    def solve_e2e():
        np.random.seed(1)
        return pca(a_host, npc=args.npc, arch='gpu!', opt=Options())

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), out, t0, t1

    # ---- resident arm ("value")
    matrix = AMatrix(a_host, arch='gpu!')
    for _ in range(args.warmup):
        lra = solve_resident(matrix)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    profile.reset()
    profile.enable(True)
    launches0 = cuda.launch_count()
    ms_val, lra, t0, t1 = timed(lambda: solve_resident(matrix), args.steps)
    launches = cuda.launch_count() - launches0
    profile.enable(False)
    prof = profile.report()
    clocks = sampler.stop(t0, t1) if sampler else None
    iterations = int(lra.iterations)
    ncomp = int(lra.left_v().nvec())
    del matrix

    # ---- end-to-end arm
    for _ in range(min(args.warmup, 3)):
        res = solve_e2e()
    ms_e2e, res, _, _ = timed(solve_e2e, args.steps)
    mean, trans, comps = res
    if world > 1:      # `trans` is gathered over ranks: check this rank's rows against its slab
        trans = trans[rank * args.rows:(rank + 1) * args.rows]
    em, ef = pca_error_gpu(a_dev, mean, trans, comps)
    h2d = a_host.nbytes
    d2h = trans.nbytes + comps.nbytes + mean.nbytes

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    roof = roofline_from(prof, peaks, args.steps)
    try:        # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        cap = json.load(open(os.path.join(ROOT, 'profiles', 'dense_apply_tc_ncu.json')))
        if roof and roof['kernel'] == 'dense_apply_tc' and (args.rows, args.cols) == (M_ROWS, N_COLS):
            roof['traffic'] = cap['dram_bytes_per_launch']
            roof['traffic_source'] = cap['source']
    except Exception:
        pass
    val_s = ms_val / 1e3 / args.steps
    e2e_s = ms_e2e / 1e3 / args.steps
    line = {
        'metric': METRIC, 'value': val_s, 'unit': 's', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_val / args.steps, 'higher_is_better': False,
        'scaling': 'weak', 'vs_baseline': val_s / BASELINE_PUBLISHED_S, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, world),
        'e2e': {'value': e2e_s, 'unit': 's', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
        'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof,
        'solver_iterations': iterations, 'components': ncomp,
        'pca_error': {'max_2norm': em, 'frobenius': ef},
        'kernels': {k: {'count': v['count'], 'ms': round(v['ms'], 3), 'GBps': round(v['GBps'], 1),
                        'TFLOPs': round(v['TFLOPs'], 2)} for k, v in prof.items()},
        'device_busy_frac': round(sum(v['ms'] for v in prof.values()) / ms_val, 4),
    }
    try:
        line['hbm_kernels'] = hbm_kernel_rates(peaks.get('hbm_gbs') or 6650.0)
    except Exception as exc:
        line['hbm_kernels'] = {'error': repr(exc)}
    line['impl'] = 'raleigh_b200'
    if ctx is not None:
        line['collectives'] = {'allreduce_calls': ctx.allreduce_calls, 'allreduce_MB': round(ctx.allreduce_bytes / 1e6, 1)}
    host_limit.restore_original_limits()
    if world == 1 and not args.no_cpu_baseline:
        try:
            load_reference_cpu()
            np.random.seed(1)
            t, desc, _ = cpu_pca_seconds(a_host, args.npc, budget_s=45.0)
            line['cpu_baseline'] = {'value': t, 'unit': 's', 'cores': os.cpu_count(), 'kind': 'reference',
                                    'sample': desc,
                                    'solver_iterations': getattr(cpu_pca_seconds, 'last_iterations', None)}
        except Exception as exc:  # the baseline must never cost us the measured line
            line['cpu_baseline'] = {'value': None, 'unit': 's', 'cores': os.cpu_count(), 'kind': 'reference',
                                    'sample': 'failed: %r' % (exc,)}
    print(json.dumps(line))


def hbm_kernel_rates(peak_gbs):
    """Block-SpMM and Gram bandwidth (the second half of BASELINE.json's metric) on the
    per-GPU block of config 4 at block size 32: n = 2,097,152 rows (256^3 / 8), fp64,
    7-point Laplacian slab of 128^3.  CUDA events around the C-ABI calls, L2 flushed
    between repetitions.  A few milliseconds in total; outside every timed region."""
    import numpy as np
    import torch
    import raleigh_b200 as rb
    from raleigh_b200._lib import lib, check
    from raleigh_b200 import device as dev
    import scipy.sparse as sp
    from raleigh_b200 import dist as rdist
    saved, rdist._current = rdist._current, None     # single-GPU micro-measurement: no sharding, no collectives
    try:
        return _hbm_kernel_rates(peak_gbs, np, torch, rb, lib, check, dev, sp)
    finally:
        rdist._current = saved


def _hbm_kernel_rates(peak_gbs, np, torch, rb, lib, check, dev, sp):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def timed(fn, reps=5, write_flush=True):
        fn()
        best = []
        for _ in range(reps):
            if write_flush:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            best.append(a.elapsed_time(b))
        best.sort()
        return best[len(best) // 2]

    out = {}
    n, m = 2097152, 32
    X, Y = rb.Vectors(n, m), rb.Vectors(n, m)
    X.fill_random_device(1)
    Y.fill_random_device(2)
    wsb = lib.rl_gram_ws_bytes(1, m, m, n)
    ws, g = dev.Buffer(wsb), dev.Buffer(m * m * 8)
    ms = timed(lambda: check(lib.rl_gram(1, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.ptr, ws.ptr, wsb,
                                         dev.stream())))
    by = 2.0 * n * m * 8
    out['gram'] = {'shape': 'n=%d, m=k=%d, fp64' % (n, m), 'ms': ms, 'GBps': by / ms / 1e6,
                   'frac_of_measured_hbm_peak': by / ms / 1e6 / peak_gbs, 'TFLOPs': 2.0 * n * m * m / ms / 1e9}
    # second reading without the write flush: the inputs (1.07 GB) exceed L2 (126 MB) on their own
    # (measured r1e: 0.247 ms vs 0.236 ms after the flush -- the flush does not penalise the kernel)
    ms2 = timed(lambda: check(lib.rl_gram(1, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.ptr, ws.ptr, wsb,
                                          dev.stream())), write_flush=False)
    out['gram'].update({'ms_inputs_exceed_l2_no_flush': ms2, 'frac_no_flush': by / ms2 / 1e6 / peak_gbs})
    N = 128

    def lap1(k):
        return sp.diags([-np.ones(k - 1), 2 * np.ones(k), -np.ones(k - 1)], [-1, 0, 1], format='csr')
    I = sp.identity(N, format='csr')
    L = (sp.kron(I, sp.kron(I, lap1(N))) + sp.kron(I, sp.kron(lap1(N), I)) + sp.kron(lap1(N), sp.kron(I, I))).tocsr()
    A = rb.SparseSymmetricMatrix(L)
    ms = timed(lambda: A.apply(X, Y))
    nnz = A.nnz()
    by = nnz * 12.0 + (n + 1) * 8.0 + 2.0 * n * m * 8
    out['spmm'] = {'shape': '7-point Laplacian 128^3 (n=%d, nnz=%d), m=%d, fp64, %s' % (n, nnz, m, A.layout()),
                   'ms': ms, 'GBps': by / ms / 1e6, 'frac_of_measured_hbm_peak': by / ms / 1e6 / peak_gbs}
    ms2 = timed(lambda: A.apply(X, Y), write_flush=False)
    out['spmm'].update({'ms_inputs_exceed_l2_no_flush': ms2, 'frac_no_flush': by / ms2 / 1e6 / peak_gbs})
    return out


def roofline_from(prof, peaks, steps):
    """Dominant kernel of the step by device time; tensor-bound dense operator
    application for this workload."""
    if not prof:
        return None
    name = max(prof, key=lambda k: prof[k]['ms'])
    rec = prof[name]
    if name.startswith('dense_apply'):
        peak = peaks.get('bf16_tflops_sustained') or 1400.0
        src = 'measured sustained bf16 (MEASURED_PEAKS.json)' if 'bf16_tflops_sustained' in peaks else 'fallback'
        return {'kernel': name, 'bound': 'tensor', 'achieved': rec['TFLOPs'], 'peak': peak, 'unit': 'TFLOP/s',
                'frac': rec['TFLOPs'] / peak, 'traffic': None, 'launches': rec['count'],
                'avg_launch_ms': rec['ms'] / rec['count'], 'share_of_device_time': rec['ms'] / sum(
                    v['ms'] for v in prof.values()),
                'peak_source': src,
                'note': 'fp32 result via tensor cores needs a 3xTF32 split: the attainable ceiling is '
                        'tf32 peak / 3 = bf16 peak / 6; HBM floor of this GEMM: %.2f ms per launch'
                        % (rec['bytes'] / rec['count'] / (peaks.get('hbm_gbs', 6650.0) * 1e6))}
    peak = peaks.get('hbm_gbs') or 6650.0
    return {'kernel': name, 'bound': 'hbm', 'achieved': rec['GBps'], 'peak': peak, 'unit': 'GB/s',
            'frac': rec['GBps'] / peak, 'traffic': None, 'launches': rec['count'],
            'avg_launch_ms': rec['ms'] / rec['count'],
            'peak_source': 'measured copy bandwidth (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback'}


def main():
    args = parse()
    wd = int(os.environ.get('RL_BENCH_WATCHDOG', '0'))
    if wd > 0:      # debugging aid: dump every thread's stack if the run is still going after `wd` seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
