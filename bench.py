#!/usr/bin/env python
"""bench.py -- time to k eigenpairs on BASELINE.json's quoted configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload = "C2"): PCA / truncated SVD of the synthetic 12,000 x 39,375 fp32 low-rank +
noise matrix (LFW 175x225 shape), 1,000 principal components -- BASELINE.json configs[1], the configuration
the reference's README quotes (27 s CPU / 12 s GPU).  At N = 1 the matrix is produced by the reference's
OWN generator (examples/pca/generate_matrix.py:74-77 `generate(12000, 39375, 2000, alpha=0.75, pca=True)`
plus the --ptb noise, :122-125, numpy.random.seed(1)) -- config 2 as written, identical in both arms.
One "step" = one complete solve: the reference's unmodified LowerRankApproximation.compute -> PartialSVD
on the raleigh_b200 backend, with the block Jacobi-CG iteration and the partial-SVD post-processing running
device-resident (raleigh_b200/jcg.py, psvd.py).

  value     seconds per solve with the data matrix already resident in HBM
  e2e       seconds per `pca(A_host, npc=1000, arch='gpu!')` call: host matrix in pinned memory -> H2D inside the
            timed region, components / scores D2H
  roofline  the dominant kernel of the step (dense operator application), measured live with CUDA events inside
            the C library over the timed region
  cpu_baseline  the reference's own CPU path (dense_numpy on NumPy/OpenBLAS; MKL is not installable here) on the
            box's host cores, in a child process that never imports the product; its singular values are
            compared with the GPU arm's (max_rel_sv_diff)
  c4        bounded BASELINE config 4 (256^3 Laplacian row-sharded over the N GPUs, fixed iteration count,
            block 32 and 120): seconds per iteration, SpMM / Gram GB/s per rank, halo and all-reduce traffic
  hbm_kernels  block SpMM and Gram bandwidth on the per-GPU block of config 4 (second half of BASELINE's metric)

--impl reference times only the reference's CPU path (no product import, no GPU) and prints the same line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_ROWS, N_COLS, RANK_GEN, NPC = 12000, 39375, 2000, 1000
METRIC = 'time to k eigenpairs (s)'
BASELINE_PUBLISHED_S = 12.0      # README.md:33 "raleigh GPU" column, 1000 components (GPU model not stated)
REFERENCE_BUDGET_S = 200.0       # the whole --impl reference run must end within a few minutes


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
    ap.add_argument('--no-c4', action='store_true', help='skip the bounded config-4 leg')
    ap.add_argument('--c4-grid', type=int, default=256)
    ap.add_argument('--c4-iters', type=int, default=10)
    ap.add_argument('--sv-out', default=None, help='(reference arm) save the singular values of the last solve')
    ap.add_argument('--full', action='store_true', help='(reference arm) always run the full workload')
    return ap.parse_args()


# --------------------------------------------------------------------------- workload
def c2_cache_path(rows, cols):
    return os.path.join(tempfile.gettempdir(), 'raleigh_b200_c2_%dx%d_seed1.npy' % (rows, cols))


def c2_host_matrix(rows, cols, rank=RANK_GEN):
    """Config 2 as written: the reference's own generator and noise recipe on the host.
    Cached in the temporary directory so that the two arms of one round (and the cpu_baseline child) share
    the ~10 s generation."""
    import numpy as np
    path = c2_cache_path(rows, cols)
    if os.path.exists(path):
        try:
            a = np.load(path)
            if a.shape == (rows, cols) and a.dtype == np.float32:
                return a
        except Exception:
            pass
    from tools import refenv
    if refenv.load_reference() is None:
        raise RuntimeError('reference package (baseline/_ref) not on this box')
    from raleigh.examples.pca.generate_matrix import generate
    np.random.seed(1)
    a, sigma, u, v = generate(rows, cols, min(rank, rows, cols), dtype=np.float32, alpha=0.75, pca=True)
    del u, v
    noise = 2 * np.random.rand(rows, cols).astype(np.float32) - 1            # generate_matrix.py:122-125
    s = 10 * np.sqrt(np.einsum('ij,ij->i', noise, noise))
    a += np.reshape(sigma[-1] / s, (rows, 1)) * noise
    del noise
    a = np.ascontiguousarray(a, dtype=np.float32)
    try:
        tmp = path + '.%d.tmp' % os.getpid()
        with open(tmp, 'wb') as fh:
            np.save(fh, a)
        os.replace(tmp, path)
    except OSError:
        pass
    return a


def generate_shard(rows, cols, rank, device, seed=1, world=1, shard=0):
    """N > 1 (weak scaling: `rows` samples per GPU): the reference's generator cannot produce one shard of a
    (world*rows) x cols matrix without forming all of it, so each rank draws its rows of a matrix of the same
    recipe (examples/pca/generate_matrix.py:55-77, 122-125) with torch: A = U diag(k^-0.75) V^T + noise,
    U[:, 0] = const, V common to all shards, U per shard scaled to stay (nearly) orthonormal globally."""
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


def workload_config(args, world):
    gen = ('reference generator examples/pca/generate_matrix.py generate(%d, %d, %d, alpha=0.75, pca=True) + --ptb noise, '
           'numpy.random.seed(1)' % (args.rows, args.cols, RANK_GEN)) if world == 1 else \
        'same recipe drawn per shard with torch (the reference generator cannot produce one shard of the global matrix)'
    return {'workload': 'C2: PCA of synthetic %dx%d fp32 low-rank+noise (LFW 175x225 shape), %d components, '
                        'block 128, svtol 1e-3%s' % (args.rows * world, args.cols, args.npc,
                                                     '' if world == 1 else ' (%d rows per GPU)' % args.rows),
            'generator': gen,
            'l2_policy': 'inputs_exceed_l2 (data matrix %.2f GB streamed every operator application)'
                         % (args.rows * args.cols * 4 / 1e9),
            'parallelism': 'single GPU' if world == 1 else
            'sample-partitioned data matrix over %d GPUs: row-sharded block vectors, NCCL all-reduce of Gram '
            'matrices and of the k x n_features partial products' % world,
            'solver': 'reference lra/partial_svd + Solver.solve unmodified; main loop and partial-SVD post-processing '
                      'device-resident (raleigh_b200/jcg.py, psvd.py: no host LAPACK)'}


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
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML (two cheap queries every
    100 ms from a thread started before the warm-up); `nvidia-smi -lms` as a fallback.  r2ai: the earlier version
    launched the nvidia-smi process right before the timed region -- its start-up (NVML initialisation over all the
    GPUs of the box) and its seven-field queries held driver locks during the measurement and cost the resident arm
    15-50 ms on some boxes (device_busy_frac 0.62 with `value` above the end-to-end number)."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')
    PERIOD = 0.1

    def __init__(self, index=0):
        self.rows = []            # (time, sm_mhz, reasons set)
        self.smax = None
        self.proc = None
        self.mode = None
        self._stop = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self.mode = 'nvml'
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                 '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = 'nvidia-smi'
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            try:
                return int(vis.split(',')[index])
            except (ValueError, IndexError):
                pass
        return index

    def _poll_nvml(self):
        nv = self._nv
        names = (('hw_slowdown', 'nvmlClocksEventReasonHwSlowdown', 0x8), ('hw_thermal_slowdown', 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
                 ('sw_thermal_slowdown', 'nvmlClocksEventReasonSwThermalSlowdown', 0x20), ('sw_power_cap', 'nvmlClocksEventReasonSwPowerCap', 0x4))
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons', None)
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = int(get_reasons(self._h)) if get_reasons else 0
                self.rows.append((time.time(), sm, {n for n, _, bit in names if mask & bit}))
            except Exception:
                pass
            time.sleep(self.PERIOD)

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.strip().split(',')]
            try:
                sm = float(parts[0])
                self.smax = float(parts[1])
            except (ValueError, IndexError):
                continue
            reasons = {name for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[3:7])
                       if val.lower().startswith('active')}
            self.rows.append((time.time(), sm, reasons))

    def stop(self, t0, t1):
        if self.mode is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml / nvidia-smi unavailable']}
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        sm, reasons = [], set()
        for ts, mhz, why in list(self.rows):
            if ts < t0 or ts > t1:
                continue
            sm.append(mhz)
            reasons |= why
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.smax, 'samples': len(sm),
                'reasons': sorted(reasons), 'source': self.mode}


# --------------------------------------------------------------------------- reference CPU path
def host_threads():
    """All the host threads the process may use (torchrun pins OMP_NUM_THREADS=1 in its children)."""
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def cpu_pca_seconds(a_host, r, k, total_rows, npc):
    """One solve with the reference's CPU pca on rows [0, r) with k components; returns (seconds extrapolated
    to the full workload of total_rows x npc, description, components, singular values, iterations)."""
    import numpy as np
    from raleigh.interfaces.lra import LowerRankApproximation
    from raleigh.algebra.dense_matrix import AMatrix
    from raleigh.core.solver import Options
    rows, cols = a_host.shape
    sample = a_host if r == rows else np.ascontiguousarray(a_host[:r])
    np.random.seed(1)
    t0 = time.perf_counter()
    # exactly what pca(sample, npc=k, arch='cpu') does (pca.py:142-153), spelled out so that the solver's
    # iteration count can be reported next to the GPU run's
    lra = LowerRankApproximation()
    lra.ortho = 1e-3 if sample.shape[0] < sample.shape[1] else 0
    lra.compute(AMatrix(sample, arch='cpu'), opt=Options(), rank=k, tol=0, norm='f', max_rank=-1, svtol=1e-3,
                shift=True, verb=0)
    trans, comps = lra.left(), lra.right()
    t = time.perf_counter() - t0
    sv = np.sqrt(np.einsum('ij,ij->j', trans.astype(np.float64), trans.astype(np.float64)))
    lib = 'reference dense_numpy on NumPy/OpenBLAS (not MKL), %d threads' % host_threads()
    if r == total_rows and k == npc:
        desc = 'full workload: pca(%dx%d fp32, npc=%d), %s' % (rows, cols, npc, lib)
        return t, desc, comps.shape[0], sv, int(lra.iterations)
    scale = (total_rows * npc) / float(r * k)
    desc = ('rows 0..%d of the %dx%d matrix, npc=%d (%.2f s measured), scaled x%.2f by rows*components to the full '
            'workload; %s' % (r, total_rows, cols, k, t, scale, lib))
    return t * scale, desc, comps.shape[0], sv, int(lra.iterations)


def run_reference(args, rank, world):
    """The reference arm: the reference's own CPU implementation of the path on the host cores.  Imports
    NumPy / SciPy / the reference only -- no raleigh_b200, no CUDA."""
    if rank != 0:
        return
    threads = host_threads()
    for var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[var] = str(threads)            # before NumPy loads its BLAS
    import numpy as np
    from tools import refenv
    if refenv.load_reference() is None:
        print(json.dumps({'impl': 'reference', 'unavailable': 'reference package (baseline/_ref) not on this box'}))
        return
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=threads)
    except Exception:
        pass
    t_gen = time.perf_counter()
    if world == 1:
        a_host = c2_host_matrix(args.rows, args.cols)
    else:
        import torch
        torch.set_num_threads(threads)
        a_host = generate_shard(args.rows, args.cols, RANK_GEN, 'cpu', seed=1, world=world, shard=0).numpy()
    t_gen = time.perf_counter() - t_gen
    total_rows = args.rows * world
    # size of one step: the full workload when K + W of them fit the budget, else a row/component-proportional
    # sample (cost ~ rows * components), extrapolated and labelled so
    x = np.random.rand(128, args.cols).astype(np.float32)
    t0 = time.perf_counter()
    y = x @ a_host.T
    _ = y @ a_host
    t_pair = time.perf_counter() - t0
    est_shard = 22 * t_pair * (args.npc / 1000.0) + 3.0            # one shard's rows, all components
    nsteps = args.steps + min(args.warmup, 1)
    per_step = max(REFERENCE_BUDGET_S - t_gen, 30.0) / max(nsteps, 1)
    r, k = args.rows, args.npc                                      # at most one shard's rows are ever held
    if not args.full and est_shard > per_step:
        k = max(64, int(args.npc * per_step / est_shard))
        if est_shard * k / args.npc > per_step:
            r = max(512, int(args.rows * per_step / (est_shard * k / args.npc)))
    times, desc, ncomp, sv, its = [], '', 0, None, None
    for i in range(nsteps):
        t, desc, ncomp, sv, its = cpu_pca_seconds(a_host, r, k, total_rows, args.npc)
        if i >= min(args.warmup, 1):
            times.append(t)
    val = sum(times) / len(times)
    if args.sv_out:
        np.save(args.sv_out, sv)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 's', 'n_gpus': args.gpus,
        'steps': args.steps, 'steps_executed': len(times), 'warmup': args.warmup,
        'warmup_executed': min(args.warmup, 1), 'ms_per_step': val * 1e3,
        'higher_is_better': False, 'scaling': 'weak', 'vs_baseline': val / 27.0, 'dtype': 'f32',
        'data': 'synthetic',
        'config': dict(workload_config(args, world), parallelism='host CPU, %d threads' % threads,
                       solver='reference core solver + lra/partial_svd on the reference\'s own dense_numpy algebra '
                              '(NumPy/OpenBLAS; MKL not installable)'),
        'cpu_baseline': {'value': val, 'unit': 's', 'cores': threads, 'kind': 'reference', 'sample': desc},
        'e2e': {'value': val, 'unit': 's', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'components': int(ncomp), 'solver_iterations': its, 'sample_rows': r, 'sample_components': k,
        'generate_s': round(t_gen, 2),
    }
    print(json.dumps(line))


def cpu_baseline_child(args):
    """cpu_baseline leg of the GPU arm: the reference arm in a child process (fresh thread settings, no
    product in the address space), one step; returns (record, singular values or None)."""
    svf = os.path.join(tempfile.gettempdir(), 'raleigh_b200_cpu_sv_%d.npy' % os.getpid())
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
           '--rows', str(args.rows), '--cols', str(args.cols), '--npc', str(args.npc), '--sv-out', svf, '--full']
    env = {k: v for k, v in os.environ.items() if k not in ('OMP_NUM_THREADS', 'RANK', 'WORLD_SIZE', 'LOCAL_RANK')}
    t0 = time.perf_counter()
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
    rec = None
    for ln in out.stdout.splitlines():
        if ln.startswith('{'):
            rec = json.loads(ln)
    if rec is None or 'value' not in rec:
        return {'value': None, 'unit': 's', 'cores': host_threads(), 'kind': 'reference',
                'sample': 'failed: %s' % (out.stderr[-300:],)}, None
    sv = None
    try:
        import numpy as np
        sv = np.load(svf)
        os.remove(svf)
    except Exception:
        pass
    base = dict(rec['cpu_baseline'])
    base['solver_iterations'] = rec.get('solver_iterations')
    base['components'] = rec.get('components')
    base['wall_s_including_generation'] = round(time.perf_counter() - t0, 1)
    return base, sv


# --------------------------------------------------------------------------- bounded config 4
def lap3d_rows(N, row0, nloc):
    """Rows [row0, row0+nloc) of the 7-point Laplacian on an N^3 grid (h = 1/(N+1), x fastest; the operator of
    examples/laplace.py:23-27) as CSR arrays built directly -- sorted columns, no COO pass, no global matrix."""
    import numpy as np
    import scipy.sparse as sp
    h2 = float(N + 1) ** 2
    r = np.arange(row0, row0 + nloc, dtype=np.int64)
    x, y, z = r % N, (r // N) % N, r // (N * N)
    offs = np.array([-N * N, -N, -1, 0, 1, N, N * N], dtype=np.int64)
    valid = np.stack([z > 0, y > 0, x > 0, np.ones(nloc, dtype=bool), x < N - 1, y < N - 1, z < N - 1], axis=1)
    del x, y, z
    cols = (r[:, None] + offs[None, :])[valid].astype(np.int32 if N ** 3 < 2 ** 31 else np.int64)
    vals = np.broadcast_to(np.array([-h2, -h2, -h2, 6.0 * h2, -h2, -h2, -h2]), (nloc, 7))[valid]
    indptr = np.zeros(nloc + 1, dtype=np.int64)
    np.cumsum(valid.sum(axis=1), out=indptr[1:])
    A = sp.csr_matrix((vals, cols, indptr), shape=(nloc, N ** 3))
    A.has_sorted_indices = True
    return A


class _StopAfter:
    def __init__(self, iters):
        self.iters = iters

    def satisfied(self, solver):
        return solver.iteration + 1 >= self.iters


def c4_leg(args, rank, world, ctx):
    """Bounded BASELINE config 4: 3D Laplacian grid^3 row-sharded over the ranks, a FIXED number of iterations of
    the device-resident block-CG driver (a full solve needs > 1000 iterations at 256^3) at
    block 32 and at the block the solver picks for 100 wanted pairs (120).  Per-iteration wall and device time,
    SpMM / Gram / update GB/s per rank from the library's CUDA-event spans, halo and all-reduce traffic."""
    import numpy as np
    import torch
    import raleigh_b200 as rb
    from raleigh_b200 import dist as rdist, profile
    import raleigh.core.solver as rs
    import torch.distributed as tdist
    N = args.c4_grid
    n = N ** 3
    t0 = time.perf_counter()
    if world > 1:
        row0, nloc = rdist.partition(n, world, rank)
    else:
        row0, nloc = 0, n
    op = rb.SparseSymmetricMatrix(lap3d_rows(N, row0, nloc), local_rows=(row0, n))
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    out = {'grid': '%d^3' % N, 'rows': n, 'rows_per_gpu': nloc, 'nnz_per_gpu': op.nnz(), 'dtype': 'f64',
           'operator_setup_s': round(setup, 2), 'iterations_run': args.c4_iters,
           'note': 'bounded: fixed iteration count, no convergence claimed'}

    def run(block, nev, iters):
        np.random.seed(1)
        opt = rs.Options()
        opt.block_size = block
        opt.max_iter = 10 ** 6
        opt.verbosity = -1
        opt.convergence_criteria = rs.DefaultConvergenceCriteria()
        opt.convergence_criteria.set_error_tolerance('k eigenvector error', 1e-6)
        opt.stopping_criteria = _StopAfter(iters)
        v = rb.Vectors(n, data_type=np.float64)
        solver = rs.Solver(rs.Problem(v, op))
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        t = time.perf_counter()
        solver.solve(v, opt, which=(nev, 0))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        return dt, solver

    free, _ = torch.cuda.mem_get_info()
    for block, nev in ((32, 20), (120, 100)):
        need = 9.5 * nloc * block * 8 + (2 << 30)
        key = 'block%d' % block
        if need > free:
            out[key] = {'skipped': 'needs %.0f GB per GPU, %.0f GB free' % (need / 1e9, free / 1e9)}
            continue
        try:
            run(block, nev, 2)                                   # warm-up: allocator, NCCL channels, plans
            profile.reset()
            profile.enable(True)
            if ctx is not None:
                ctx.allreduce_calls = ctx.allreduce_bytes = 0
            op.halo_bytes = 0
            dt, solver = run(block, nev, args.c4_iters)
            profile.enable(False)
            prof = profile.report()
            its = max(int(solver.iteration) + 1, 1)
            if world > 1:
                t = torch.tensor([dt], device='cuda')
                tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
                dt = t.item()
            rec = {'seconds_per_iteration': round(dt / its, 5), 'iterations': its,
                   'device_ms_per_iteration': round(sum(v['ms'] for v in prof.values()) / its, 3),
                   'kernels': {k: {'count': v['count'], 'ms': round(v['ms'], 2), 'GBps': round(v['GBps'], 1)}
                               for k, v in prof.items() if k in ('spmm', 'gram', 'update', 'dots', 'rr_solve', 'piv_chol')}}
            if ctx is not None:
                rec['nccl'] = {'allreduce_calls_per_iteration': round(ctx.allreduce_calls / its, 1),
                               'allreduce_KB_per_iteration': round(ctx.allreduce_bytes / its / 1e3, 1),
                               'halo_MB_per_iteration': round(getattr(op, 'halo_bytes', 0) / its / 1e6, 2)}
            out[key] = rec
        except Exception as exc:                                 # the leg must never cost the headline line
            out[key] = {'error': repr(exc)[:300]}
        torch.cuda.empty_cache()
    if ctx is not None:
        g = torch.zeros(240 * 240, dtype=torch.float64, device='cuda')
        for _ in range(3):
            tdist.all_reduce(g)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            tdist.all_reduce(g)
        b.record()
        torch.cuda.synchronize()
        out['allreduce_240x240_f64_ms'] = round(a.elapsed_time(b) / 20, 4)
    return out


# --------------------------------------------------------------------------- GPU arm
def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    from threadpoolctl import threadpool_limits
    host_limit = threadpool_limits(limits=1)         # the host keeps only length-m bookkeeping
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

    a_pinned = torch.empty((args.rows, args.cols), dtype=torch.float32, pin_memory=True)
    if world == 1:
        a_pinned.copy_(torch.from_numpy(c2_host_matrix(args.rows, args.cols)))
        a_dev = a_pinned.cuda()
    else:
        a_dev = generate_shard(args.rows, args.cols, RANK_GEN, 'cuda', seed=1, world=world, shard=rank)
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
    sampler = ClockSampler(local_rank) if rank == 0 else None       # started BEFORE the warm-up: see its docstring
    for _ in range(args.warmup):
        lra = solve_resident(matrix)
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
    del matrix, lra

    # ---- end-to-end arm
    for _ in range(min(args.warmup, 3)):
        res = solve_e2e()
    ms_e2e, res, _, _ = timed(solve_e2e, args.steps)
    mean, trans, comps = res
    sv_gpu = np.sqrt(np.einsum('ij,ij->j', trans.astype(np.float64), trans.astype(np.float64)))
    if world > 1:      # `trans` is gathered over ranks: check this rank's rows against its slab
        trans = trans[rank * args.rows:(rank + 1) * args.rows]
    em, ef = pca_error_gpu(a_dev, mean, trans, comps)
    h2d = a_host.nbytes
    d2h = res[1].nbytes + comps.nbytes + mean.nbytes
    del a_dev
    torch.cuda.empty_cache()

    c4 = None
    if not args.no_c4:
        try:
            c4 = c4_leg(args, rank, world, ctx)
        except Exception as exc:
            c4 = {'error': repr(exc)[:300]}
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
    if c4 is not None:
        line['c4'] = c4
    line['impl'] = 'raleigh_b200'
    if ctx is not None:
        line['collectives'] = {'allreduce_calls': ctx.allreduce_calls, 'allreduce_MB': round(ctx.allreduce_bytes / 1e6, 1)}
    host_limit.restore_original_limits()
    if world == 1 and not args.no_cpu_baseline:
        try:
            base, sv_cpu = cpu_baseline_child(args)
            line['cpu_baseline'] = base
            if sv_cpu is not None and len(sv_cpu) == len(sv_gpu):
                rel = np.abs(sv_gpu - sv_cpu) / sv_cpu
                line['max_rel_sv_diff'] = {'leading_100': float(rel[:100].max()), 'all': float(rel.max()),
                                           'note': 'singular values (column norms of the scores) of the GPU arm vs the '
                                                   'cpu_baseline run on the same matrix; both solve to svtol = 1e-3, so '
                                                   'the trailing values differ at that level between any two runs'}
        except Exception as exc:  # the baseline must never cost us the measured line
            line['cpu_baseline'] = {'value': None, 'unit': 's', 'cores': host_threads(), 'kind': 'reference',
                                    'sample': 'failed: %r' % (exc,)}
    print(json.dumps(line))


def hbm_kernel_rates(peak_gbs):
    """Block-SpMM and Gram bandwidth (the second half of BASELINE.json's metric) on the
    per-GPU block of config 4: n = 2,097,152 rows (256^3 / 8), fp64, 7-point Laplacian slab of 128^3, block 32
    (and the Gram product at block 120).  CUDA events around the C-ABI calls, L2 flushed
    between repetitions.  A few milliseconds in total; outside every timed region."""
    import numpy as np
    import torch
    import raleigh_b200 as rb
    from raleigh_b200._lib import lib, check
    from raleigh_b200 import device as dev
    from raleigh_b200 import dist as rdist
    saved, rdist._current = rdist._current, None     # single-GPU micro-measurement: no sharding, no collectives
    try:
        return _hbm_kernel_rates(peak_gbs, np, torch, rb, lib, check, dev)
    finally:
        rdist._current = saved


def _hbm_kernel_rates(peak_gbs, np, torch, rb, lib, check, dev):
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
    n = 2097152
    for m in (32, 120):
        X, Y = rb.Vectors(n, m), rb.Vectors(n, m)
        X.fill_random_device(1)
        Y.fill_random_device(2)
        wsb = lib.rl_gram_ws_bytes(1, m, m, n)
        ws, g = dev.Buffer(wsb), dev.Buffer(m * m * 8)
        ms = timed(lambda: check(lib.rl_gram(1, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.ptr, ws.ptr, wsb,
                                             dev.stream())))
        by = 2.0 * n * m * 8
        key = 'gram' if m == 32 else 'gram_block120'
        out[key] = {'shape': 'n=%d, m=k=%d, fp64' % (n, m), 'ms': ms, 'GBps': by / ms / 1e6,
                    'frac_of_measured_hbm_peak': by / ms / 1e6 / peak_gbs, 'TFLOPs': 2.0 * n * m * m / ms / 1e9}
        if m == 32:
            # second reading without the write flush: the inputs (1.07 GB) exceed L2 (126 MB) on their own
            ms2 = timed(lambda: check(lib.rl_gram(1, X._wptr(), X._ld, m, Y._wptr(), Y._ld, m, n, g.ptr, ws.ptr, wsb,
                                                  dev.stream())), write_flush=False)
            out[key].update({'ms_inputs_exceed_l2_no_flush': ms2, 'frac_no_flush': by / ms2 / 1e6 / peak_gbs})
            A = rb.SparseSymmetricMatrix(lap3d_rows(128, 0, n), local_rows=(0, n))
            ms = timed(lambda: A.apply(X, Y))
            nnz = A.nnz()
            by = nnz * 12.0 + (n + 1) * 8.0 + 2.0 * n * m * 8
            out['spmm'] = {'shape': '7-point Laplacian 128^3 (n=%d, nnz=%d), m=%d, fp64, %s' % (n, nnz, m, A.layout()),
                           'ms': ms, 'GBps': by / ms / 1e6, 'frac_of_measured_hbm_peak': by / ms / 1e6 / peak_gbs}
            ms2 = timed(lambda: A.apply(X, Y), write_flush=False)
            out['spmm'].update({'ms_inputs_exceed_l2_no_flush': ms2, 'frac_no_flush': by / ms2 / 1e6 / peak_gbs})
            del A
        del X, Y
    return out


def roofline_from(prof, peaks, steps):
    """Roofline of the dominant HOT-PATH kernel of the step by device time: the dense operator application
    (tensor-bound) for this workload.  The Rayleigh-Ritz spans (rr_solve, piv_chol, syevj) are chains of
    latency-bound small kernels with no bandwidth / flop roofline of their own; their share of the device time
    is reported next to it."""
    if not prof:
        return None
    total = sum(v['ms'] for v in prof.values())
    sized = {k: v for k, v in prof.items() if k not in ('rr_solve', 'piv_chol', 'syevj', 'small_dense')}
    name = max(sized, key=lambda k: sized[k]['ms'])
    rec = prof[name]
    small = {k: round(prof[k]['ms'] / total, 3) for k in ('rr_solve', 'piv_chol', 'syevj', 'small_dense') if k in prof}
    if name.startswith('dense_apply'):
        peak = peaks.get('bf16_tflops_sustained') or 1400.0
        src = 'measured sustained bf16 (MEASURED_PEAKS.json)' if 'bf16_tflops_sustained' in peaks else 'fallback'
        return {'kernel': name, 'bound': 'tensor', 'achieved': rec['TFLOPs'], 'peak': peak, 'unit': 'TFLOP/s',
                'frac': rec['TFLOPs'] / peak, 'traffic': None, 'launches': rec['count'],
                'avg_launch_ms': rec['ms'] / rec['count'], 'share_of_device_time': rec['ms'] / total,
                'small_dense_algebra_share_of_device_time': small,
                'algorithmic_GBps': rec['GBps'],
                'peak_source': src,
                'note': 'fp32 result via tensor cores needs a 3xTF32 split: the attainable ceiling is '
                        'tf32 peak / 3 = bf16 peak / 6; HBM floor of this GEMM: %.2f ms per launch'
                        % (rec['bytes'] / rec['count'] / (peaks.get('hbm_gbs', 6650.0) * 1e6))}
    peak = peaks.get('hbm_gbs') or 6650.0
    return {'kernel': name, 'bound': 'hbm', 'achieved': rec['GBps'], 'peak': peak, 'unit': 'GB/s',
            'frac': rec['GBps'] / peak, 'traffic': None, 'launches': rec['count'],
            'avg_launch_ms': rec['ms'] / rec['count'], 'share_of_device_time': rec['ms'] / total,
            'small_dense_algebra_share_of_device_time': small,
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
