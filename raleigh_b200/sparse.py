"""Sparse symmetric operator and preconditioner wrappers on the device.

Drop-in for the parts of raleigh/algebra/sparse_mkl.py that lie on the
preconditioned branch of partial_hevp (partial_hevp.py:202-224):
`SparseSymmetricMatrix` (sparse_mkl.py:16-48) and `Operator`
(sparse_mkl.py:143-154).  The reference has no GPU sparse path at all
(README.md:49); MKL's `mkl_?csrmm` with descr 'SUNF' (mkl_wrap.py:264-276) is
what `apply` computes.  `SparseSymmetricSolver` (PARDISO) and `IncompleteLU`
are out of scope and raise.
"""
import numpy
import scipy.sparse as scs

from . import _lib
from ._lib import lib, check
from . import device as dev
from .vectors import Vectors


class SparseSymmetricMatrix:
    """Y = A_sym X with A_sym the symmetric matrix whose upper triangle is
    stored (the reference keeps only `triu(A)`; sparse_mkl.py:18-31).  The
    device holds the FULL matrix as 0-based CSR (int64 indptr, int32 indices) so
    that one gather-only SpMM kernel serves it; building it is one-off host
    set-up, like the reference's own `triu` + `sort_indices`."""

    def __init__(self, matrix, local_rows=None):
        """`matrix`: SciPy sparse matrix (or another SparseSymmetricMatrix), as in the
        reference.  Under an active ShardContext (dist.enable()) the operator is
        row-partitioned: every rank keeps the rows of its slab; pass
        `local_rows=(row0, n_global)` together with a (nloc, n_global) matrix holding
        just those rows of the FULL symmetric operator to avoid ever forming the
        global matrix (config 4 generates its slabs this way)."""
        from . import dist
        ctx = dist.current()
        self.__plan = None
        if local_rows is not None:
            row0, n_global = int(local_rows[0]), int(local_rows[1])
            slab = matrix.tocsr()
            if not slab.has_sorted_indices:
                slab.sort_indices()
            if ctx is None or ctx.world == 1 or not ctx.shard_matrices:
                # single process: `matrix` must be ALL rows of the full symmetric operator (no triu / mirror pass:
                # config 4's 117 M entries are generated directly in this form)
                if row0 != 0 or slab.shape[0] != n_global:
                    raise ValueError('local_rows without an active ShardContext needs the whole operator')
                full = slab
            csr = None
            self.__n = n_global
        else:
            try:
                csr = matrix.csr()
            except Exception:
                csr = scs.triu(matrix, format='csr')
                csr.sort_indices()
            strict = scs.triu(csr, k=1, format='csr')
            full = (csr + strict.T).tocsr()
            full.sort_indices()
            self.__n = csr.shape[0]
            slab = full
            if ctx is not None and ctx.shard_matrices and ctx.world > 1:
                row0, nloc = dist.partition(self.__n, ctx.world, ctx.rank)
                slab = full[row0:row0 + nloc].tocsr()
                slab.sort_indices()
        self.__csr = csr
        dtype = slab.data.dtype.type
        self.__dtype = slab.data.dtype
        self.__code = _lib.dtype_code(dtype)          # raises ValueError for complex
        self.__nnz = int(slab.nnz)
        self.__nrows = slab.shape[0]
        if ctx is not None and ctx.shard_matrices and ctx.world > 1:
            nloc = slab.shape[0]
            if local_rows is None:
                row0 = dist.partition(self.__n, ctx.world, ctx.rank)[0]
            ctx.register(self.__n, row0, nloc)
            self.__plan = dist.HaloPlan(ctx, slab.indptr, slab.indices, row0, nloc, self.__n)
            self.__send_idx = _to_device(self.__plan.send_idx)
            self.__full_diag = numpy.asarray(slab[:, row0:row0 + nloc].diagonal())
            full = scs.csr_matrix((slab.data, self.__plan.local_indices, slab.indptr),
                                  shape=(nloc, nloc + self.__plan.nhalo))
        indptr = numpy.ascontiguousarray(full.indptr, dtype=numpy.int64)
        indices = numpy.ascontiguousarray(full.indices, dtype=numpy.int32)
        values = numpy.ascontiguousarray(full.data, dtype=dtype)
        self.__indptr = _to_device(indptr)
        self.__indices = _to_device(indices)
        self.__values = _to_device(values)
        if self.__plan is None:
            self.__full_diag = full.diagonal()
        self.__sell = _build_sell32(indptr, indices, values) if SPMM_LAYOUT == 'sell' else None
        self.__order, self.__warps, self.footprint_ratio = None, 0, None
        if self.__sell is None and SPMM_CLUSTER_WARPS > 0 and self.__nrows >= 32 * SPMM_CLUSTER_WARPS:
            order, self.footprint_ratio = cluster_runs(indptr, indices, self.__nrows, SPMM_CLUSTER_WARPS)
            self.__order, self.__warps = _to_device(order), SPMM_CLUSTER_WARPS

    def size(self):
        return self.__n

    def data_type(self):
        return self.__dtype

    def csr(self):
        return self.__csr

    def nnz(self):
        """Stored entries of the full (mirrored) device matrix."""
        return self.__nnz

    def diagonal(self):
        return self.__full_diag

    def apply(self, x, y):
        if not isinstance(x, Vectors) or not isinstance(y, Vectors):
            raise ValueError('SparseSymmetricMatrix.apply needs device Vectors')
        m = x.nvec()
        if m != y.nvec():
            raise ValueError('Numbers of input and output vectors differ')
        if x.dimension() != self.__n or y.dimension() != self.__n:
            raise ValueError('Matrix and vectors dimensions incompatible')
        if x.data_type() != self.__dtype.type or y.data_type() != self.__dtype.type:
            raise ValueError('Matrix and vectors data types differ')
        if m < 1:
            return
        y._touch()
        ncl, halo = 0, 0
        if self.__plan is not None:
            if not (x.is_sharded() and y.is_sharded() and x.local_dimension() == self.__nrows):
                raise ValueError('vectors are not laid out like the row-sharded operator')
            ncl, halo = self.__nrows, self._exchange_halo(x, m)
        if self.__sell is not None:
            sp, sc, sv, nsl = self.__sell
            check(lib.rl_sell_spmm_halo(self.__code, self.__nrows, self.__nnz, nsl, sp.ptr, sc.ptr, sv.ptr,
                                        x._wptr(), x._ld, y._wptr(), y._ld, m, ncl, halo, dev.stream()))
            return
        check(lib.rl_csr_spmm_ex(self.__code, self.__nrows, self.__nnz, self.__indptr.ptr, self.__indices.ptr,
                                 self.__values.ptr, x._wptr(), x._ld, y._wptr(), y._ld, m, ncl, halo,
                                 self.__order.ptr if self.__order is not None else None, self.__warps,
                                 dev.stream()))

    def _exchange_halo(self, x, m):
        """Pack the boundary rows every peer needs (kernel), all-to-all over NVLink
        (NCCL), return the device address of the received row-interleaved halo."""
        import torch
        plan = self.__plan
        tdt = torch.float32 if x.data_type() is numpy.float32 else torch.float64
        nsend = int(plan.send_idx.shape[0])
        send = torch.empty(max(nsend * m, 1), dtype=tdt, device='cuda')
        recv = torch.empty(max(plan.nhalo * m, 1), dtype=tdt, device='cuda')
        if nsend:
            check(lib.rl_pack_rows(x._code, x._wptr(), x._ld, m, self.__send_idx.ptr, nsend, send.data_ptr(),
                                   dev.stream()))
        plan.exchange(send[:nsend * m], recv[:plan.nhalo * m], m)
        self.__halo_keepalive = (send, recv)
        self.halo_bytes = getattr(self, 'halo_bytes', 0) + (nsend + plan.nhalo) * m * x.data_size()
        return recv.data_ptr()

    def layout(self):
        if self.__sell is not None:
            return 'sell32'
        return 'csr' if self.__order is None else 'csr+clustered%d' % self.__warps


import os as _os
# 32-row runs per CTA of the staged-CSR kernel, grouped at set-up by shared column footprint
# (rl_spmm_cluster_runs) so that stencil neighbours in y and z are L1 hits; 0 = consecutive runs
SPMM_CLUSTER_WARPS = int(_os.environ.get('RALEIGH_B200_SPMM_CLUSTER', '0'))
# 'csr' (default: staged-CSR kernel, r1e: faster than SELL-32 on every matrix measured once its gather
# addressing was fixed) or 'sell' (SELL-32 when the padding rule below allows it)
SPMM_LAYOUT = _os.environ.get('RALEIGH_B200_SPMM_LAYOUT', 'csr')


def cluster_runs(indptr, indices, nrows, group):
    """Footprint-clustered order of the 32-row runs (host set-up, C++ in the library)."""
    import ctypes
    nruns = (nrows + 31) // 32
    order = numpy.empty(nruns, dtype=numpy.int32)
    ratio = ctypes.c_double()
    check(lib.rl_spmm_cluster_runs(nrows, dev.host_ptr(indptr), dev.host_ptr(indices), group, dev.host_ptr(order),
                                   ctypes.byref(ratio)))
    return order, ratio.value


SELL_MAX_PADDING = 1.5     # use SELL-32 only if it stores at most this many times nnz entries
SELL_MIN_ROW_NNZ = 24      # ... and rows are long.  History (profiles/r1c_kernel_tuning.md, r1e_gram_spmm.md):
                           # SELL-32 used to win at 55 nnz/row (0.67 vs 0.38 TB/s); with hoisted gather
                           # addressing both kernels got faster and staged CSR with 4-entry batches leads
                           # (1.03 vs 0.90 TB/s), so SELL-32 is opt-in (RALEIGH_B200_SPMM_LAYOUT=sell)


def _build_sell32(indptr, indices, values):
    """CSR -> SELL-32 on the host (one-off set-up).  Returns device buffers
    (slice_ptr int64, cols int32, vals) and the slice count, or None when the
    row lengths are too ragged for the padded layout to pay off."""
    n = indptr.shape[0] - 1
    nnz = int(indptr[-1])
    if n == 0 or nnz == 0:
        return None
    lens = numpy.diff(indptr)
    nsl = (n + 31) // 32
    padded = numpy.zeros(nsl * 32, dtype=numpy.int64)
    padded[:n] = lens
    width = padded.reshape(nsl, 32).max(axis=1)
    total = int(width.sum()) * 32
    if total > SELL_MAX_PADDING * nnz + 1024 or nnz < SELL_MIN_ROW_NNZ * n:
        return None
    slice_ptr = numpy.zeros(nsl + 1, dtype=numpy.int64)
    numpy.cumsum(width * 32, out=slice_ptr[1:])
    rows = numpy.repeat(numpy.arange(n, dtype=numpy.int64), lens)
    q = numpy.arange(nnz, dtype=numpy.int64) - indptr[rows]
    dest = slice_ptr[rows // 32] + q * 32 + (rows % 32)
    cols = numpy.zeros(total, dtype=numpy.int32)
    # padding entries: val = 0, col = 0 (always a valid column)
    vals = numpy.zeros(total, dtype=values.dtype)
    cols[dest] = indices
    vals[dest] = values
    return _to_device(slice_ptr), _to_device(cols), _to_device(vals), nsl


def _to_device(a):
    buf = dev.Buffer(max(a.nbytes, 16))
    if a.nbytes:
        check(lib.rl_h2d(buf.ptr, dev.host_ptr(a), a.nbytes, dev.stream()))
        check(lib.rl_sync_stream(dev.stream()))
    return buf


class DiagonalPreconditioner:
    """Jacobi preconditioner y = x / diag(A) as a device operator: pass it as
    `T` to partial_hevp (or wrap it in Operator).  The user-object contract of
    the reference (`T.apply(x, y)` on 2-D ndarrays, partial_hevp.py:64-73) is
    also honoured for host arrays so the same object works on the CPU oracle."""

    def __init__(self, A):
        if isinstance(A, SparseSymmetricMatrix):
            d = A.diagonal()
        elif scs.issparse(A):
            d = A.diagonal()
        else:
            d = numpy.asarray(A).reshape(-1)
        self.__inv = 1.0 / numpy.asarray(d)
        self.__dev = {}

    def _inv_on_device(self, dtype, x=None):
        key = (numpy.dtype(dtype).type, 0, self.__inv.shape[0])
        inv = self.__inv
        if x is not None and x.is_sharded() and inv.shape[0] == x.dimension() != x.local_dimension():
            row0 = x._shard[1]                      # diagonal given globally: keep this rank's rows
            key = (key[0], row0, x.local_dimension())
            inv = inv[row0:row0 + x.local_dimension()]
        if key not in self.__dev:
            self.__dev[key] = _to_device(numpy.ascontiguousarray(inv, dtype=key[0]))
        return self.__dev[key]

    def apply(self, x, y):
        if isinstance(x, Vectors):
            m = x.nvec()
            if m < 1:
                return
            d = self._inv_on_device(x.data_type(), x)
            y._touch()
            check(lib.rl_diag_mul(x._code, y._wptr(), y._ld, x._wptr(), x._ld, m, x.local_dimension(), d.ptr,
                                  dev.stream()))
        else:
            y[...] = x * self.__inv[None, :].astype(x.dtype)


class Operator:
    """sparse_mkl.py:143-154.  The reference hands `x.data()`, `y.data()` (host
    ndarrays that alias the vectors) to the user's `op.apply`; on the device
    `data()` is a copy, so device-aware operators (anything that accepts
    Vectors, e.g. DiagonalPreconditioner) get the Vectors themselves, and a
    host-only user operator is served by a round trip through host memory."""

    def __init__(self, op):
        self.__op = op
        self.__device_aware = isinstance(op, (DiagonalPreconditioner, SparseSymmetricMatrix)) or \
            getattr(op, 'accepts_device_vectors', False)

    def apply(self, x, y):
        if self.__device_aware:
            self.__op.apply(x, y)
            return
        xh = x.data()
        yh = numpy.empty_like(xh)
        self.__op.apply(xh, yh)
        y.fill(yh)


class SparseSymmetricSolver:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('shift-and-invert (PARDISO, sparse_mkl.py:51-119) is out of scope of raleigh_b200')


class IncompleteLU:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('ILUT preconditioning (sparse_mkl.py:122-140) is out of scope of raleigh_b200')
