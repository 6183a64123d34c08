"""Make the UNMODIFIED reference package run on this backend.

The reference's interfaces reach their algebra through hard-coded module
imports (partial_hevp.py:14-17 -> dense_cblas / sparse_mkl; dense_matrix.py:17-18
-> cuda_wrap / dense_cublas).  `install()` registers this package's modules under
those names *before* the interfaces are imported, so

    import raleigh_b200; raleigh_b200.install()
    from raleigh.interfaces.pca import pca                 # arch='gpu!'
    from raleigh.interfaces.partial_hevp import partial_hevp

run verbatim on the B200 kernels.  It also applies the SciPy >= 1.14 shim the
reference needs on any backend: `scipy.linalg.eigh(..., turbo=False)`
(solver.py:578,822,899,1470) no longer accepts `turbo`.
"""
import importlib
import os
import sys

_ALIASES = {
    'dense_cublas': 'raleigh_b200.vectors',
    'dense_cblas': 'raleigh_b200.vectors',
    'cuda_wrap': 'raleigh_b200.cuda',
    'sparse_mkl': 'raleigh_b200.sparse',
}


def find_reference():
    """Directory that contains the reference's `raleigh` package, or None.
    Order: already importable, $RALEIGH_REFERENCE, <repo>/baseline/_ref."""
    try:
        spec = importlib.util.find_spec('raleigh')
        if spec is not None and spec.submodule_search_locations:
            return os.path.dirname(list(spec.submodule_search_locations)[0])
    except (ImportError, ValueError):
        pass
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get('RALEIGH_REFERENCE'), os.path.join(here, 'baseline', '_ref')):
        if cand and os.path.isdir(os.path.join(cand, 'raleigh')):
            return cand
    return None


class _SlaProxy:
    """scipy.linalg with the removed `turbo` keyword of eigh() accepted and dropped."""

    def __init__(self, sla):
        self._sla = sla

    def __getattr__(self, name):
        return getattr(self._sla, name)

    def eigh(self, *args, turbo=None, **kwargs):
        return self._sla.eigh(*args, **kwargs)


class _NumpyProxy:
    """numpy with eye() accepting the float sizes the reference computes as
    `min(32, nsv/2)` (partial_svd.py:199-204); NumPy >= 2 rejects them."""

    def __init__(self, np):
        self._np = np

    def __getattr__(self, name):
        return getattr(self._np, name)

    def eye(self, n, *args, **kwargs):
        return self._np.eye(int(n), *args, **kwargs)


def _int_min(*args, **kwargs):
    import builtins
    r = builtins.min(*args, **kwargs)
    return int(r) if isinstance(r, float) else r


def shim_scipy():
    """Version shims the reference needs on ANY backend with current SciPy/NumPy;
    they touch no algebra."""
    import numpy
    import scipy.linalg as sla
    import raleigh.core.solver as rsolver
    if not isinstance(rsolver.sla, _SlaProxy):
        rsolver.sla = _SlaProxy(sla)
    try:
        import raleigh.interfaces.partial_svd as psvd
        if not isinstance(psvd.numpy, _NumpyProxy):
            psvd.numpy = _NumpyProxy(numpy)
        # `nv = min(32, nsv/2)` (partial_svd.py:201) is Python-2 integer division: under
        # Python 3 it yields a float that then breaks `select()` / slicing on the reference's
        # own NumPy backend.  A module-level `min` restores the integer result.
        psvd.min = _int_min
    except ImportError:
        pass


def _column_norms(a, axis):
    """2-norms along `axis` of a 2-D array in one vectorised pass (same value as the
    per-column `numpy.linalg.norm` loop up to the rounding of the summation order)."""
    import numpy
    if a.dtype.kind == 'c':
        sq = a.real * a.real + a.imag * a.imag
        return numpy.sqrt(sq.sum(axis=axis))
    return numpy.sqrt(numpy.einsum('ij,ij->j' if axis == 0 else 'ij,ij->i', a, a))


class _BigLapackProxy:
    """scipy.linalg for interfaces/partial_svd.py: `_finalize_svd` (partial_svd.py:162-235) factorises
    nsv x nsv matrices (1000 x 1000 at config 2: eigh, svd, inv) on the host, while the solver's own
    2m x 2m problems want ONE BLAS thread (256 x 256 LAPACK is slower multi-threaded, which is why the
    GPU arm runs with the pool limited to 1).  Calls on matrices of order >= 512 get a few threads back
    for their duration; everything else passes through untouched."""
    BIG = 512

    def __init__(self, inner, threads):
        self._inner = inner
        self._threads = threads

    def __getattr__(self, name):
        return getattr(self._inner, name)

    def _run(self, fn, a, *args, **kwargs):
        if self._threads > 1 and getattr(a, 'ndim', 0) == 2 and a.shape[0] >= self.BIG:
            from threadpoolctl import threadpool_limits
            with threadpool_limits(limits=self._threads):
                return fn(a, *args, **kwargs)
        return fn(a, *args, **kwargs)

    def svd(self, a, *args, **kwargs):
        return self._run(self._inner.svd, a, *args, **kwargs)

    def eigh(self, a, *args, **kwargs):
        return self._run(self._inner.eigh, a, *args, **kwargs)

    def inv(self, a, *args, **kwargs):
        return self._run(self._inner.inv, a, *args, **kwargs)


class _DenseMatrixNumpyProxy:
    """numpy for raleigh/algebra/dense_matrix.py: AMatrix.__init__ (:32-34) runs numpy.amin and
    numpy.amax over the host data matrix right after `Matrix(a)` has uploaded it -- two passes
    over 1.9 GB (~0.2 s) at config 2.  When `a` is the array that was just uploaded, both numbers
    come from one pass over the device copy (rl_minmax_h); anything else goes to NumPy."""

    def __init__(self, np):
        self._np = np
        self._cache = None            # (id(a), lo, hi)

    def __getattr__(self, name):
        return getattr(self._np, name)

    def _device_minmax(self, a):
        from . import vectors
        ent = vectors._recent_uploads.get(id(a))
        if ent is None or not isinstance(a, self._np.ndarray) or ent[1] != a.shape or ent[2] != a.dtype:
            return None
        if self._cache is not None and self._cache[0] == id(a):
            return self._cache[1:]
        mat = ent[0]()
        if mat is None:
            return None
        lo, hi = mat.minmax()
        self._cache = (id(a), lo, hi)
        return lo, hi

    def amin(self, a, *args, **kwargs):
        r = None if (args or kwargs) else self._device_minmax(a)
        return self._np.amin(a, *args, **kwargs) if r is None else r[0]

    def amax(self, a, *args, **kwargs):
        r = None if (args or kwargs) else self._device_minmax(a)
        if r is None:
            return self._np.amax(a, *args, **kwargs)
        from . import vectors
        vectors._recent_uploads.pop(id(a), None)      # AMatrix asks for amin, then amax: entry consumed
        self._cache = None
        return r[1]


def _big_lapack_threads():
    """Threads a big LAPACK call may use: at most 4, and never more than this process's share of the
    cores it may run on -- one process per GPU means up to 8 of them factorise at the same moment, and
    an oversubscribed OpenBLAS is catastrophically slow (measured: 16 threads on 8 cores, 0.3 s -> 10 s)."""
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    try:
        procs = int(os.environ.get('LOCAL_WORLD_SIZE') or os.environ.get('WORLD_SIZE') or 1)
    except ValueError:
        procs = 1
    return max(1, min(4, cores // max(1, procs)))


def shim_host_hotspots():
    """`solver._norm` (solver.py:1745-1746) is `numpy.apply_along_axis(numpy.linalg.norm, ...)`:
    one Python-level call per column, called for every pivot of `_piv_chol` (solver.py:1765-1768)
    -- about 150 000 interpreter round trips per config-2 solve, 35-45 % of the wall time once
    the algebra runs on the GPU.  Same quantity in a single vectorised pass; the solver source
    stays untouched.  RALEIGH_B200_FAST_HOST=0 keeps the reference's helper."""
    if os.environ.get('RALEIGH_B200_FAST_HOST', '1') == '0':
        return False
    import raleigh.core.solver as rsolver
    if rsolver._norm is not _column_norms:
        rsolver._reference_norm = rsolver._norm
        rsolver._norm = _column_norms
    try:
        import numpy
        import raleigh.algebra.dense_matrix as dmat
        if not isinstance(dmat.numpy, _DenseMatrixNumpyProxy):
            dmat.numpy = _DenseMatrixNumpyProxy(numpy)
    except ImportError:
        pass
    try:
        import threadpoolctl  # noqa: F401
        import raleigh.interfaces.partial_svd as psvd
        threads = int(os.environ.get('RALEIGH_B200_BIG_LAPACK_THREADS', _big_lapack_threads()))
        if not isinstance(psvd.sla, _BigLapackProxy):
            psvd._reference_sla = psvd.sla
            psvd.sla = _BigLapackProxy(psvd.sla, threads)
    except ImportError:
        pass
    return True


def unshim_host_hotspots():
    """Put the reference's own `_norm` back (the CPU reference arm of bench.py runs unmodified)."""
    try:
        import raleigh.core.solver as rsolver
    except ImportError:
        return
    if getattr(rsolver, '_reference_norm', None) is not None:
        rsolver._norm = rsolver._reference_norm
    try:
        import raleigh.interfaces.partial_svd as psvd
        if getattr(psvd, '_reference_sla', None) is not None:
            psvd.sla = psvd._reference_sla
    except ImportError:
        pass
    try:
        import raleigh.algebra.dense_matrix as dmat
        if isinstance(dmat.numpy, _DenseMatrixNumpyProxy):
            dmat.numpy = dmat.numpy._np
    except ImportError:
        pass


DEVICE_SOLVER = os.environ.get('RALEIGH_B200_DEVICE_SOLVER', '1') != '0'


def use_device_solver(on=True):
    """Switch between the device-resident block-CG driver (jcg.py, default) and the reference's own
    main loop running verbatim on this backend (kept for parity checks)."""
    global DEVICE_SOLVER
    DEVICE_SOLVER = bool(on)


def hook_device_solver():
    """Route `Solver._solve` (solver.py:587) to the device-resident driver whenever the problem is one
    it supports (standard problem on raleigh_b200 vectors); everything else, and everything when
    DEVICE_SOLVER is off, goes to the reference's own `_solve`.  `Solver.solve` -- argument checks,
    block size choice, the final dense Rayleigh-Ritz fallback -- stays the reference's."""
    import raleigh.core.solver as rsolver
    if getattr(rsolver.Solver, '_reference_solve', None) is not None:
        return
    rsolver.Solver._reference_solve = rsolver.Solver._solve

    def _solve(self, eigenvectors, options, which, extra, init):
        if DEVICE_SOLVER:
            from . import jcg
            if jcg.supported(self, eigenvectors):
                from .engine import DeviceEngine
                return jcg.solve(self, eigenvectors, options, which, extra, init, DeviceEngine())
        return self._reference_solve(eigenvectors, options, which, extra, init)

    rsolver.Solver._solve = _solve


def hook_device_finalize_svd():
    """PartialSVD._finalize_svd (partial_svd.py:163-235) -> psvd.finalize_svd whenever the vectors are
    raleigh_b200 blocks and the device solver is on; the reference's host LAPACK version otherwise."""
    try:
        import raleigh.interfaces.partial_svd as rpsvd
    except ImportError:
        return
    cls = rpsvd.PartialSVD
    if getattr(cls, '_reference_finalize_svd', None) is not None:
        return
    cls._reference_finalize_svd = staticmethod(cls._finalize_svd)

    def _finalize(v, Av, eps):
        if DEVICE_SOLVER and hasattr(v, '_rl_device_block'):
            from . import psvd
            return psvd.finalize_svd(v, Av, eps)
        return cls._reference_finalize_svd(v, Av, eps)

    cls._finalize_svd = staticmethod(_finalize)


class _LraNumpyProxy:
    """numpy for raleigh/interfaces/lra.py: `concatenate` of two device-data handles (see vectors.DeviceData)
    along axis 1 happens on the device; every other call, and every other argument, goes to NumPy."""

    def __init__(self, np):
        self._np = np

    def __getattr__(self, name):
        return getattr(self._np, name)

    def concatenate(self, arrays, axis=0, *args, **kwargs):
        from .vectors import DeviceData
        arrays = tuple(arrays)
        if axis == 1 and len(arrays) == 2 and not args and not kwargs and all(isinstance(a, DeviceData) for a in arrays):
            return arrays[0].concatenate(arrays[1])
        arrays = tuple(self._np.asarray(a) if isinstance(a, DeviceData) else a for a in arrays)
        return self._np.concatenate(arrays, axis, *args, **kwargs)


LAZY_UPDATE_BYTES = 16 << 20


def hook_device_lra_update():
    """LowerRankApproximation.update (lra.py:158-379) runs verbatim, but while it runs the two big `data()` calls
    of its left-factor growth (lra.py:287-290) stay on the device: see vectors.DeviceData.  Off with the device
    solver switch (use_device_solver(False)) so that parity runs can compare both routes."""
    try:
        import numpy
        import raleigh.interfaces.lra as rlra
    except ImportError:
        return
    cls = rlra.LowerRankApproximation
    if getattr(cls, '_reference_update', None) is not None:
        return
    cls._reference_update = cls.update
    if not isinstance(rlra.numpy, _LraNumpyProxy):
        rlra.numpy = _LraNumpyProxy(rlra.numpy)

    def update(self, matrix, *args, **kwargs):
        from . import vectors
        from . import dist
        ctx = dist.current()
        sharded = ctx is not None and ctx.world > 1
        lazy = (DEVICE_SOLVER or sharded) and hasattr(matrix.as_vectors(), '_rl_device_block')
        saved = vectors.LAZY_DATA_MIN_BYTES
        # sample-partitioned run: the left factor is row-sharded and its pieces must be joined process by process
        # (Vectors.append(axis=1)), which only the device route does -- every data() goes through the stand-in
        vectors.LAZY_DATA_MIN_BYTES = (0 if sharded else LAZY_UPDATE_BYTES) if lazy else None
        try:
            return cls._reference_update(self, matrix, *args, **kwargs)
        finally:
            vectors.LAZY_DATA_MIN_BYTES = saved

    cls.update = update

    if getattr(cls, '_reference_icompute', None) is None:
        cls._reference_icompute = cls.icompute

        def icompute(self, matrix, batch_size, *args, **kwargs):
            # chunked run: the chunks are row slices of `matrix` taken in order (lra.py:403-421) -- let the Matrix
            # constructor upload the next one in the background while this one is processed (vectors._ChunkPrefetch)
            from . import vectors
            arch = kwargs.get('arch', 'cpu')
            saved = vectors.CHUNK_PREFETCH
            vectors.CHUNK_PREFETCH = bool(DEVICE_SOLVER and isinstance(arch, str) and arch[:3] == 'gpu')
            try:
                return cls._reference_icompute(self, matrix, batch_size, *args, **kwargs)
            finally:
                vectors.CHUNK_PREFETCH = saved
                vectors._chunk_prefetch.drop()

        cls.icompute = icompute


def hook_device_operator_svd():
    """_OperatorSVD.apply (partial_svd.py:258-291) -> opsvd.apply: the rank-one mean-shift corrections keep their
    coefficients on the device, no pipeline drain per operator application."""
    try:
        import raleigh.interfaces.partial_svd as rpsvd
    except ImportError:
        return
    cls = rpsvd._OperatorSVD
    if getattr(cls, '_reference_apply', None) is not None:
        return
    cls._reference_apply = cls.apply

    def apply(self, x, y):
        if DEVICE_SOLVER:
            from . import opsvd
            if opsvd.supported(self, x, y):
                return opsvd.apply(self, x, y)
        return cls._reference_apply(self, x, y)

    cls.apply = apply


def install(reference_path=None, sparse=True, dense=True):
    """Alias the backend into `raleigh.algebra`; returns the `raleigh` package.
    Raises ImportError if the reference package cannot be found."""
    path = reference_path or find_reference()
    if path is None:
        raise ImportError('reference package `raleigh` not found (looked at sys.path, '
                          '$RALEIGH_REFERENCE, baseline/_ref)')
    if path not in sys.path:
        sys.path.insert(0, path)
    import raleigh
    import raleigh.algebra as algebra
    # Import the reference's CPU selector FIRST so that arch='cpu' keeps its own
    # algebra (dense_cpu.py:10-17 picks MKL cblas or NumPy) before dense_cblas is aliased.
    try:
        importlib.import_module('raleigh.algebra.dense_cpu')
    except Exception:
        pass
    for name, target in _ALIASES.items():
        if name == 'sparse_mkl' and not sparse:
            continue
        if name != 'sparse_mkl' and not dense:
            continue
        mod = importlib.import_module(target)
        sys.modules['raleigh.algebra.' + name] = mod
        setattr(algebra, name, mod)
    shim_scipy()
    shim_host_hotspots()
    hook_device_solver()
    hook_device_finalize_svd()
    hook_device_lra_update()
    hook_device_operator_svd()
    return raleigh
